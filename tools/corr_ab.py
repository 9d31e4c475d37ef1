"""A/B of the two correlation kernels (the `kernel` argument of sie_corr_tau): R bitwise, tau, time.  Stored-R sweep shape + 25 km tau-only."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from seaiceextentforecasting_b200 import synthetic as syn
from seaiceextentforecasting_b200 import _lib
from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest

KERNELS = (("tiles", _lib.SIE_CORR_TILES), ("rows", _lib.SIE_CORR_ROWS))
STORED = KERNELS + (("mirror", _lib.SIE_CORR_ROWS_MIRROR), ("tma", _lib.SIE_CORR_TMA))


def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def stored(X, Y, Ts, latlon=False):
    B = len(Ts); T = max(Ts); C = X * Y
    data, _ = syn.make_field(X, Y, T, 11)
    n_upper = int((~np.isnan(data).all(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, B, latlon=latlon, n_upper=n_upper, keep_R=True, max_areas=8)
    fields = h2d(data.reshape(1, C, T))
    jf = torch.zeros(B, dtype=torch.int32, device="cuda"); jT = torch.tensor(Ts, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(t, 0.01) for t in Ts]))
    eng.detrend_zscore(fields, jf, jT, True); torch.cuda.synchronize()
    out = {}
    for kern, kid in STORED:
        eng.R.fill_(-7.0)
        eng.corr_tau(rc, store_R=True, kernel=kid); torch.cuda.synchronize()
        N = eng.n_nodes.cpu().numpy()
        out[kern] = ([eng.R[b, :N[b], :N[b]].cpu().numpy().copy() for b in range(min(B, 6))], eng.tau.cpu().numpy().copy(),
                     eng.tau_cnt.cpu().numpy().copy())
        ms = timed(lambda: eng.corr_tau(rc, store_R=True, kernel=kid))
        byts = float(((8.0 if kern == "mirror" else 4.0) * N.astype(np.float64) ** 2).sum())
        flop = float((N.astype(np.float64) * (N + 1.0) * np.asarray(Ts, dtype=np.float64)).sum())
        print(f"{X}x{Y} B={B} {kern}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s stored (4N^2; mirror 8N^2), {flop/ms/1e9:.1f} TFLOP/s  N={N[0]} ldn={eng.ldn}")
    for b, (a, c) in enumerate(zip(out["tiles"][0], out["rows"][0])):
        iu = np.triu_indices(a.shape[0], 1)
        same = np.array_equal(a[iu], c[iu])            # the tile kernel stores the upper triangle only
        f = out["mirror"][0][b]
        sym = np.array_equal(f, f.T, equal_nan=True) and np.array_equal(f[iu], a[iu])     # the mirror-writing variant
        t = out["tma"][0][b]
        assert np.array_equal(t[iu], a[iu]) and np.isnan(np.diag(t)).all(), "tma kernel differs"
        print(f"  job {b}: R bitwise equal {same}, symmetric {sym}, diag NaN {np.isnan(np.diag(c)).all()}, untouched {(c == -7.0).sum()}")
        assert same and sym
    print("  tau rel diff", np.abs(out["tiles"][1] - out["rows"][1]).max() / np.abs(out["tiles"][1]).max(), "cnt equal", np.array_equal(out["tiles"][2], out["rows"][2]))
    assert np.array_equal(out["tiles"][2], out["rows"][2])
    del eng


def tau_only():
    X, Y, T = 448, 304, 42
    data, _ = syn.make_field(X, Y, T, 7); C = X * Y
    n_upper = int((~np.isnan(data).any(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=8)
    fields = h2d(data.reshape(1, C, T))
    jf = torch.zeros(1, dtype=torch.int32, device="cuda"); jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
    eng.detrend_zscore(fields, jf, jT, True); torch.cuda.synchronize()
    N = int(eng.n_nodes.item())
    res = {}
    for kern, kid in KERNELS:
        for shards in (1, 8):
            ms = timed(lambda: eng.corr_tau(rc, store_R=False, shard_rank=0, shard_count=shards, kernel=kid), reps=3)
            print(f"25km {kern} shards={shards}: {ms:.2f} ms {N*(N+1.0)*T/shards/ms/1e9:.2f} TFLOP/s")
        eng.corr_tau(rc, store_R=False, kernel=kid); torch.cuda.synchronize()
        res[kern] = (eng.tau_sum.item(), eng.tau_cnt.item())
        ss, cc = 0.0, 0
        for r in range(4):
            eng.corr_tau(rc, store_R=False, shard_rank=r, shard_count=4, kernel=kid); torch.cuda.synchronize(); ss += eng.tau_sum.item(); cc += eng.tau_cnt.item()
        print(f"  {kern}: unsharded {res[kern]}  4 shards {ss, cc}")
        assert cc == res[kern][1]
    assert res["tiles"][1] == res["rows"][1]
    print("  tau_sum rel diff tiles/rows", abs(res["tiles"][0] - res["rows"][0]) / abs(res["tiles"][0]))


def tma_only():
    """time the stored-R default alone on the 4-member sweep shape (chunk / ring depth tuning in a SIE_TUNE build)"""
    X = Y = 57
    Ts = sorted([7 + i % 36 for i in range(576)], reverse=True)
    if len(sys.argv) > 2:
        Ts = [int(sys.argv[2])] * 288                 # uniform window: where the kernel sits against each roof
    B = len(Ts); T = max(Ts); C = X * Y
    data, _ = syn.make_field(X, Y, T, 11)
    n_upper = int((~np.isnan(data).all(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, B, latlon=False, n_upper=n_upper, keep_R=True, max_areas=8)
    fields = h2d(data.reshape(1, C, T))
    jf = torch.zeros(B, dtype=torch.int32, device="cuda"); jT = torch.tensor(Ts, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(t, 0.01) for t in Ts]))
    eng.detrend_zscore(fields, jf, jT, True); torch.cuda.synchronize()
    ms = timed(lambda: eng.corr_tau(rc, store_R=True, kernel=_lib.SIE_CORR_TMA))
    N = eng.n_nodes.cpu().numpy().astype(np.float64)
    flop = float((N * (N + 1.0) * np.asarray(Ts, dtype=np.float64)).sum()); byts = float((4.0 * N * N).sum())
    print(f"tma B={B} T={min(Ts)}..{max(Ts)}: {ms:.3f} ms  {flop / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.0f} GB/s stored")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "tma":
        tma_only(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tau":
        tau_only(); sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stored":
        stored(57, 57, [7 + (i * 35) // 23 for i in range(24)])
        stored(57, 57, sorted([7 + i % 36 for i in range(576)], reverse=True)); sys.exit(0)     # 4 members x 4 inits of the sweep
    stored(20, 22, [7, 12, 30, 42])
    stored(57, 57, [7 + (i * 35) // 23 for i in range(24)])
    stored(26, 90, [9, 42], latlon=True)
    tau_only()

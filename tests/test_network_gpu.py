"""GPU parity: drop-in Network (K1-K6 through the C ABI) vs the oracle on identical seeded inputs.
Bar (BASELINE.json north_star): nodes / area membership / order bit-exact; floats <= 1e-9 relative."""
import warnings

import numpy as np
import pytest

from seaiceextentforecasting_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

CASES = [  # X, Y, T, latlon, seed
    (14, 14, 20, False, 1), (16, 12, 9, False, 2), (10, 24, 30, True, 3), (12, 30, 12, True, 4),
    (20, 20, 42, False, 5), (24, 24, 7, False, 6), (26, 90, 42, True, 7), (33, 31, 17, False, 8),
]


def _oracle(dt, latlon, kw):
    from oracle.network import Network as ON
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = ON(data=dt)
        ON.tau(o, 0.01, keep_corrs=False)
        ON.area_level(o, latlon_grid=latlon)
        ON.intra_links(o, **kw)
    return o


def _product(dt, latlon, kw):
    from seaiceextentforecasting_b200.ComplexNetworks import Network
    n = Network(data=dt)
    Network.tau(n, 0.01)
    Network.area_level(n, latlon_grid=latlon)
    Network.intra_links(n, **kw)
    return n


def compare_networks(n, o, dt):
    assert n.nodes.dtype == np.int64 and n.nodes.shape == o.nodes.shape
    assert np.array_equal(n.nodes, o.nodes)                                     # bit-exact node mask
    R = n.correlation_matrix()
    assert np.array_equal(R, R.T, equal_nan=True)                               # bitwise symmetric (H1)
    assert np.isnan(np.diag(R)).all()
    m = ~np.isnan(o.R)
    assert np.max(np.abs(R[m] - o.R[m])) <= 1e-9                                # |R| <= 1: absolute == relative bar
    assert abs(n.tau - o.tau) <= 1e-9 * abs(o.tau)
    assert list(n.V.keys()) == list(o.V.keys())                                 # dict order
    for k in o.V:
        assert n.V[k] == o.V[k], f"area {k} differs"                            # membership AND list order
    assert n.V is n.A
    for k in o.V:
        np.testing.assert_allclose(n.anomaly[k], o.anomaly[k], rtol=1e-9, atol=1e-9 * np.abs(o.anomaly[k]).max())
        sc = max(1e-300, np.abs(np.asarray(o.links[k])).max())
        np.testing.assert_allclose(np.asarray(n.links[k], dtype=float), np.asarray(o.links[k], dtype=float),
                                   rtol=1e-9, atol=1e-9 * sc)
        assert abs(n.strength[k] - o.strength[k]) <= 1e-9 * abs(o.strength[k])
    assert np.array_equal(np.isnan(n.strengthmap), np.isnan(o.strengthmap))
    ok = ~np.isnan(o.strengthmap)
    np.testing.assert_allclose(n.strengthmap[ok], o.strengthmap[ok], rtol=1e-9)


MORE_CASES = [  # wider seeded sweep of shapes / window lengths (T = 7 gives many two-cell areas, long T few big ones)
    (30, 28, 7, False, 101), (30, 28, 8, False, 102), (41, 37, 12, False, 103), (41, 37, 42, False, 104),
    (18, 60, 7, True, 105), (18, 60, 25, True, 106), (26, 90, 9, True, 107), (57, 57, 7, False, 108),
    (57, 57, 19, False, 109), (48, 52, 33, False, 110), (22, 44, 42, True, 111), (64, 40, 15, False, 112),
]


@pytest.mark.parametrize("X,Y,T,latlon,seed", CASES + MORE_CASES)
def test_network_matches_oracle(lib_built, X, Y, T, latlon, seed):
    from oracle.gp import detrend as odetrend
    from seaiceextentforecasting_b200.forecast import detrend
    data, _ = syn.make_field(X, Y, T, seed, latlon=latlon)
    dt, trend = detrend(data)
    odt, otrend = odetrend(data)
    assert np.array_equal(np.isnan(dt), np.isnan(odt))
    ok = ~np.isnan(odt)
    assert np.max(np.abs(dt[ok] - odt[ok])) <= 1e-12                            # residuals of O(1) series
    okt = ~np.isnan(otrend)
    np.testing.assert_allclose(trend[okt], otrend[okt], rtol=1e-9, atol=1e-12)
    kw = {"lat": syn.make_lat_grid(X, Y)} if latlon else {"area": syn.make_psar(X, Y)}
    # both sides start from the SAME detrended field (the reference's Network takes `dt` as input)
    n = _product(odt, latlon, kw)
    o = _oracle(odt, latlon, kw)
    compare_networks(n, o, odt)


def test_network_full_north_grid(lib_built):
    """57x57x42 (config 1 size): the oracle finishes in seconds, the reference takes ~45 s."""
    from oracle.gp import detrend as odetrend
    data, _ = syn.make_field(57, 57, 42, 11)
    odt, _ = odetrend(data)
    kw = {"area": syn.make_psar(57, 57)}
    n = _product(odt, False, kw)
    o = _oracle(odt, False, kw)
    compare_networks(n, o, odt)
    assert len(n.V) >= 2


@pytest.mark.parametrize("X,Y,T,latlon,seed", [(96, 96, 20, False, 201), (64, 144, 12, True, 202)])
def test_network_above_8192_cells_32bit_evicted_variant(lib_built, X, Y, T, latlon, seed):
    """Grids of >= 8192 cells take the <512 threads, 32-bit indices> domain-growth variant with per-cell arrays evicted to
    global scratch (csrc/area.cu plan_area: the `place` mask) -- the variant the 25 km builds use, here at a size the
    oracle finishes in ~15 s, on a polar and on a lat-lon (wrapping) grid: bit-exact domains like every other grid."""
    from oracle.gp import detrend as odetrend
    assert X * Y >= 8192
    data, _ = syn.make_field(X, Y, T, seed, latlon=latlon)
    odt, _ = odetrend(data)
    kw = {"lat": syn.make_lat_grid(X, Y)} if latlon else {"area": syn.make_psar(X, Y)}
    n = _product(odt, latlon, kw)
    o = _oracle(odt, latlon, kw)
    compare_networks(n, o, odt)
    assert len(n.V) >= 50


def test_network_south_grid(lib_built):
    """81x81 (config 3, south/February1st.py:79): 6561 cells do not fit the all-on-chip layout of k_area_level, so
    this exercises the placement that evicts integer arrays to global scratch."""
    from oracle.gp import detrend as odetrend
    data, _ = syn.make_field(81, 81, 20, 5)
    odt, _ = odetrend(data)
    kw = {"area": syn.make_psar(81, 81)}
    n = _product(odt, False, kw)
    o = _oracle(odt, False, kw)
    compare_networks(n, o, odt)
    assert len(n.V) >= 2


def test_network_giant_area(lib_built):
    """A coherent field whose largest area has ~2400 cells: areas beyond one pairwise leaf (re-summing growth path,
    multi-level pairwise trees) and beyond the 1024-row dense block of the merge step (gather path)."""
    from oracle.gp import detrend as odetrend
    data, _ = syn.make_field(64, 64, 24, 5, n_modes=3, noise=0.08, blob=(14., 24.))
    odt, _ = odetrend(data)
    kw = {"area": syn.make_psar(64, 64)}
    n = _product(odt, False, kw)
    o = _oracle(odt, False, kw)
    assert max(len(v) for v in o.V.values()) > 1024
    compare_networks(n, o, odt)


def test_errors_like_reference(lib_built):
    from seaiceextentforecasting_b200.ComplexNetworks import Network
    rng = np.random.default_rng(0)
    data = rng.standard_normal((6, 6, 12))          # no NaN cell -> IndexError (ComplexNetworks.py:50-51)
    n = Network(data=data)
    Network.tau(n, 0.01)
    with pytest.raises(IndexError):
        Network.area_level(n)
    data[0, 0, :] = np.nan                           # white noise: no area reaches tau -> ValueError (:212)
    n = Network(data=data)
    Network.tau(n, 0.01)
    with pytest.raises(ValueError):
        Network.area_level(n)


def test_corrs_lazy_view(lib_built):
    from seaiceextentforecasting_b200.ComplexNetworks import Network
    data, _ = syn.make_field(10, 10, 16, 21)
    from oracle.gp import detrend as odetrend
    odt, _ = odetrend(data)
    n = Network(data=odt)
    Network.tau(n, 0.01)
    from oracle.network import Network as ON
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        o = ON(data=odt)
        ON.tau(o, 0.01)
    c = np.asarray(n.corrs)
    assert c.shape == o.corrs.shape
    assert np.array_equal(np.isnan(c), np.isnan(o.corrs))
    m = ~np.isnan(o.corrs)
    assert np.max(np.abs(c[m] - o.corrs[m])) <= 1e-9
    assert n.corrs[3, 4, 5] == c[3, 4, 5] or np.isnan(c[3, 4, 5])


def test_row_sharded_tau_matches_unsharded_and_numpy(lib_built):
    """configs[3] / SURVEY.md 8(e): `sie_corr_tau(shard_rank, shard_count)` with R not stored -- the shards' partial
    (sum, count) add up to the unsharded result, and tau equals the reference formula evaluated with numpy."""
    import torch
    from scipy import stats
    from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
    from seaiceextentforecasting_b200.parallel import tau_from_shards
    X, Y, T = 40, 36, 25
    data, _ = syn.make_field(X, Y, T, 9)
    n_upper = int((~np.isnan(data).any(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=8)
    f = h2d(data.reshape(1, X * Y, T))
    jf = torch.zeros(1, dtype=torch.int32, device="cuda")
    jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(T, 0.01)]))
    eng.detrend_zscore(f, jf, jT, True)
    eng.corr_tau(rc, store_R=False)
    torch.cuda.synchronize()
    s_all, c_all, tau_all = eng.tau_sum.item(), eng.tau_cnt.item(), eng.tau.item()
    for world in (2, 3, 8):
        ss, cc = 0.0, 0
        for r in range(world):
            eng.corr_tau(rc, store_R=False, shard_rank=r, shard_count=world)
            torch.cuda.synchronize()
            ss += eng.tau_sum.item()
            cc += eng.tau_cnt.item()
        assert cc == c_all
        assert abs(ss - s_all) <= 1e-12 * abs(s_all)
    assert abs(tau_from_shards(eng.tau_sum * 0 + s_all, eng.tau_cnt * 0 + c_all).item() - tau_all) <= 1e-15
    # the reference's formula on the detrended field (ComplexNetworks.py:32-47)
    from oracle.gp import detrend as odetrend
    dt, _ = odetrend(data)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ID = np.where(np.abs(np.nanmax(dt, 2)) > 0)
        R = np.corrcoef(dt[ID])
        np.fill_diagonal(R, np.nan)
        df = T - 2
        P = stats.t.sf(R * np.sqrt(df / (1 - R ** 2)), df)
        ref = np.mean(R[(R >= 0) & (P < 0.01)])
    assert abs(tau_all - ref) <= 1e-9 * abs(ref)


@pytest.mark.parametrize("X,Y,Ts,latlon", [(20, 22, [7, 12, 30, 42], False), (57, 57, [7, 25, 42], False),
                                           (26, 90, [9, 42], True), (81, 81, [33], False)])
def test_both_correlation_kernels_agree(lib_built, X, Y, Ts, latlon):
    """`sie_corr_tau` has three kernels (csrc/corr.cu): the TMA-store kernel (default when R is stored), the tile kernel
    (its fallback for very long windows) and the row-resident, warp-specialised one (default for the tau-only pass).  Either can serve either mode (the `kernel` argument): the stored
    upper triangle of R (what every reader addresses, R[min][max]) must be bitwise identical with a NaN diagonal and no
    element left unwritten, the count identical and tau within 1e-12; SIE_CORR_ROWS_MIRROR also writes the mirror image, which
    must equal it bitwise; R matches numpy's corrcoef of the detrended nodes to 1e-9 (ComplexNetworks.py:34-35)."""
    import torch
    from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
    B, T, C = len(Ts), max(Ts), X * Y
    data, _ = syn.make_field(X, Y, T, 21)
    n_upper = int((~np.isnan(data).all(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, B, latlon=latlon, n_upper=n_upper, keep_R=True, max_areas=8)
    fields = h2d(data.reshape(1, C, T))
    jf = torch.zeros(B, dtype=torch.int32, device="cuda")
    jT = torch.tensor(Ts, dtype=torch.int32, device="cuda")
    rc = h2d(np.array([r_crit_ttest(t, 0.01) for t in Ts]))
    eng.detrend_zscore(fields, jf, jT, True)
    torch.cuda.synchronize()
    N = eng.n_nodes.cpu().numpy()
    out = {}
    from seaiceextentforecasting_b200 import _lib
    for kern, kid in (("tiles", _lib.SIE_CORR_TILES), ("rows", _lib.SIE_CORR_ROWS), ("mirror", _lib.SIE_CORR_ROWS_MIRROR),
                      ("tma", _lib.SIE_CORR_TMA)):
        eng.R.fill_(-7.0)
        eng.corr_tau(rc, store_R=True, kernel=kid)
        torch.cuda.synchronize()
        Rs = [eng.R[b, :N[b], :N[b]].cpu().numpy().copy() for b in range(B)]
        stored = (eng.tau.cpu().numpy().copy(), eng.tau_cnt.cpu().numpy().copy())
        eng.corr_tau(rc, store_R=False, kernel=_lib.SIE_CORR_ROWS if kern == "tma" else kid)   # tma: stored mode only
        torch.cuda.synchronize()
        out[kern] = (Rs, stored, (eng.tau.cpu().numpy().copy(), eng.tau_cnt.cpu().numpy().copy()))
    for b in range(B):
        a, c, f, t = out["tiles"][0][b], out["rows"][0][b], out["mirror"][0][b], out["tma"][0][b]
        iu = np.triu_indices(a.shape[0], 1)
        assert np.array_equal(a[iu], c[iu]) and np.array_equal(a[iu], f[iu]) and np.array_equal(a[iu], t[iu])
        assert np.isnan(np.diag(t)).all()
        assert not (a[iu] == -7.0).any() and np.isfinite(a[iu]).all()
        assert np.isnan(np.diag(a)).all() and np.isnan(np.diag(c)).all() and np.isnan(np.diag(f)).all()
        assert np.array_equal(f, f.T, equal_nan=True) and not (f == -7.0).any()       # the mirror-writing variant
    dtb = eng.dt[0, :, :Ts[0]].cpu().numpy()
    nodes = eng.node_cell[0, :N[0]].cpu().numpy()
    ref = np.corrcoef(dtb[nodes])
    np.fill_diagonal(ref, np.nan)
    m = ~np.isnan(ref)
    assert np.max(np.abs(out["mirror"][0][0][m] - ref[m])) <= 1e-9
    for mode in (1, 2):
        assert np.array_equal(out["tiles"][mode][1], out["rows"][mode][1])
        np.testing.assert_allclose(out["tiles"][mode][0], out["rows"][mode][0], rtol=1e-12)
    assert np.array_equal(out["tma"][1][1], out["rows"][1][1])            # same pairs counted; its own summation order
    np.testing.assert_allclose(out["tma"][1][0], out["rows"][1][0], rtol=1e-12)
    assert np.array_equal(out["rows"][1][1], out["rows"][2][1])
    np.testing.assert_allclose(out["rows"][1][0], out["rows"][2][0], rtol=1e-12)


def test_full_size_25km_correlation_properties(lib_built):
    """BASELINE.json configs[3] at its full size (448x304 grid, T = 42, ~63 k nodes; the oracle cannot run it: R would be
    32 GB) through size-independent properties of `sie_corr_tau` with R not stored: the (sum, count) partials of 8 row
    shards add up to the unsharded result (count exactly), both kernels count the same pairs and agree on tau, and tau is
    the mean of correlations that all exceed the critical value."""
    import torch
    from seaiceextentforecasting_b200.engine import NetworkBatch, h2d, r_crit_ttest
    X, Y, T = 448, 304, 42
    data, _ = syn.make_field(X, Y, T, 7)
    n_upper = int((~np.isnan(data).any(axis=2)).sum())
    eng = NetworkBatch(X, Y, T, 1, latlon=False, n_upper=n_upper, keep_R=False, max_areas=8)
    f = h2d(data.reshape(1, X * Y, T))
    jf = torch.zeros(1, dtype=torch.int32, device="cuda")
    jT = torch.full((1,), T, dtype=torch.int32, device="cuda")
    rcv = r_crit_ttest(T, 0.01)
    rc = h2d(np.array([rcv]))
    eng.detrend_zscore(f, jf, jT, True)
    res = {}
    from seaiceextentforecasting_b200 import _lib
    for kern, kid in (("rows", _lib.SIE_CORR_ROWS), ("tiles", _lib.SIE_CORR_TILES)):
        eng.corr_tau(rc, store_R=False, kernel=kid)
        torch.cuda.synchronize()
        res[kern] = (eng.tau_sum.item(), eng.tau_cnt.item(), eng.tau.item())
    N = int(eng.n_nodes.item())
    assert N > 60000
    s_all, c_all, tau_all = res["rows"]
    assert res["tiles"][1] == c_all and abs(res["tiles"][0] - s_all) <= 1e-12 * s_all
    assert 0 < c_all < N * (N - 1) and c_all % 2 == 0                      # both triangles
    assert rcv < tau_all <= 1.0 and abs(tau_all - s_all / c_all) <= 1e-15
    ss, cc = 0.0, 0
    for r in range(8):
        eng.corr_tau(rc, store_R=False, shard_rank=r, shard_count=8)
        torch.cuda.synchronize()
        ss += eng.tau_sum.item()
        cc += eng.tau_cnt.item()
    assert cc == c_all and abs(ss - s_all) <= 1e-12 * s_all
    # a 512-node sample against numpy: the count of significant pairs inside the sample, from z rows of the device
    idx = np.random.default_rng(0).choice(N, 512, replace=False)
    z = eng.z[0, :N, :T].cpu().numpy()[np.sort(idx)]
    Rs = z @ z.T
    dtc = eng.dt[0].cpu().numpy()[eng.node_cell[0, :N].cpu().numpy()[np.sort(idx)], :T]
    ref = np.corrcoef(dtc)
    assert np.max(np.abs(Rs - ref)) <= 1e-9                                 # unit-norm rows reproduce np.corrcoef


@pytest.mark.parametrize("X,Y,T,latlon,seed", [(20, 20, 42, False, 5), (26, 90, 42, True, 7), (57, 57, 19, False, 109),
                                               (57, 57, 42, False, 9), (18, 60, 7, True, 105), (81, 81, 12, False, 31)])
def test_network_without_stored_matrix(lib_built, X, Y, T, latlon, seed):
    """The no-R path (large grids: `tau` runs the tau-only correlation pass, `area_level` recomputes every correlation it
    consumes from the unit-norm rows, `corrs[n]` recomputes rows on demand) against the stored-R path and the oracle:
    nodes, V (keys, membership, order) bit-exact; tau <= 1e-12; node series equal; correlation rows <= 2 ulp."""
    from seaiceextentforecasting_b200.ComplexNetworks import Network
    from seaiceextentforecasting_b200.forecast import detrend

    class NoMatrix(Network):
        max_matrix_bytes = 0

    data, _ = syn.make_field(X, Y, T, seed, latlon=latlon)
    dt, _ = detrend(data)
    kw = {"lat": syn.make_lat_grid(X, Y)} if latlon else {"area": syn.make_psar(X, Y)}
    a = _product(dt, latlon, kw)
    b = NoMatrix(data=dt)
    NoMatrix.tau(b, 0.01)
    assert b._eng.R is None
    NoMatrix.area_level(b, latlon_grid=latlon)
    NoMatrix.intra_links(b, **kw)
    o = _oracle(dt, latlon, kw)
    assert np.array_equal(a.nodes, b.nodes) and np.array_equal(b.nodes, o.nodes)
    assert abs(a.tau - b.tau) <= 1e-12 * abs(a.tau) and abs(b.tau - o.tau) <= 1e-9 * abs(o.tau)
    assert list(b.V.keys()) == list(o.V.keys()) and all(b.V[k] == o.V[k] for k in o.V)      # bit-exact domains
    assert list(a.V.keys()) == list(b.V.keys()) and all(a.V[k] == b.V[k] for k in a.V)
    for k in a.V:
        assert np.array_equal(a.anomaly[k], b.anomaly[k])
    rows = [0, 1, b.nodes.shape[1] // 2, b.nodes.shape[1] - 1]
    Ra, Rb = a.correlation_rows(rows), b.correlation_rows(rows)
    assert np.array_equal(np.isnan(Ra), np.isnan(Rb))
    m = ~np.isnan(Ra)
    assert np.abs(Ra[m] - Rb[m]).max() <= 4.5e-16          # tensor-core accumulation vs the sequential FMA chain
    print(f"{X}x{Y}x{T}: recomputed rows bitwise equal to the stored (DMMA) matrix: {np.array_equal(Ra[m], Rb[m])}")
    ca, cb = np.asarray(a.corrs[rows[2]]), np.asarray(b.corrs[rows[2]])
    assert ca.shape == cb.shape == (X, Y) and np.array_equal(np.isnan(ca), np.isnan(cb))

"""GPU parity of the ingest kernels against the reference's own numpy / scipy call
sequence (oracle/ingest.py) on synthetic NSIDC-format files.  Decode, monthly mean and polar-hole fill are bit-exact
(same operation order); the regrid is scipy's LinearNDInterpolator evaluated from host-built barycentric weights and is
checked to 1e-12 absolute (values in [0, 1])."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def synth_files(rng, n_files, dimX, dimY):
    base = rng.integers(0, 251, size=(dimX, dimY))
    files = []
    for _ in range(n_files):
        img = np.clip(base + rng.integers(-20, 21, size=(dimX, dimY)), 0, 250).astype(np.uint8)
        img[rng.uniform(size=(dimX, dimY)) < 0.02] = 254          # land / coast flags (> 250 -> > 1 -> NaN)
        img[:3, :] = 253
        files.append(bytes(rng.integers(0, 256, size=300, dtype=np.uint8)) + img.tobytes())
    return files


def source_grid(dimX, dimY):
    """A 25 km lattice in the reference's projection plane, centred on the pole, with its lat/lon."""
    from seaiceextentforecasting_b200.ingest import polar_stereo
    xs = (np.arange(dimY) - dimY / 2.0 + 0.37) * 25e3
    ys = (np.arange(dimX) - dimX / 2.0 + 0.21) * 25e3
    X, Y = np.meshgrid(xs, ys)
    lon, lat = polar_stereo(X, Y, 360.0, inverse=True)
    return X, Y, lon, lat


@pytest.mark.parametrize("n_files", [1, 5, 30, 31])
def test_decode_monthly_and_hole_fill_bit_exact(lib_built, n_files):
    import torch
    from oracle import ingest as oi
    from seaiceextentforecasting_b200.engine import h2d
    from seaiceextentforecasting_b200.ingest import nsidc_monthly, polar_hole_fill
    dimX, dimY = 120, 96
    rng = np.random.default_rng(n_files)
    files = synth_files(rng, n_files, dimX, dimY)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = oi.decode_monthly(files, dimX, dimY)
    got = nsidc_monthly(files, C_cells=dimX * dimY)
    g = got.cpu().numpy().reshape(dimX, dimY)
    assert np.array_equal(np.isnan(g), np.isnan(ref))
    assert np.array_equal(g[~np.isnan(ref)], ref[~np.isnan(ref)])          # bit-exact
    _, _, _, lat = source_grid(dimX, dimY)
    for hole in (84.5, 87.2, 89.2):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rf, rp = oi.hole_fill(ref.copy(), lat, hole)
        filled, phole = polar_hole_fill(got, h2d(lat.reshape(-1)), hole)
        f = filled.cpu().numpy().reshape(dimX, dimY)
        p = float(phole.item())
        assert (np.isnan(p) and np.isnan(rp)) or p == rp
        assert np.array_equal(np.isnan(f), np.isnan(rf))
        assert np.array_equal(f[~np.isnan(rf)], rf[~np.isnan(rf)])


def test_regrid_matches_griddata(lib_built):
    import torch
    from oracle import ingest as oi
    from seaiceextentforecasting_b200.engine import h2d
    from seaiceextentforecasting_b200.ingest import Regridder, make_npstere_grid, polar_stereo
    dimX, dimY = 448, 304
    X, Y, lon, lat = source_grid(dimX, dimY)
    lonr, latr, xr, yr, (x0, y0) = make_npstere_grid(65, 360, 1e5)
    assert xr.shape == (57, 57)                                            # the grid every north script relies on
    x, y = polar_stereo(lon, lat, 360.0, x0, y0)
    rng = np.random.default_rng(0)
    fields = rng.uniform(0, 1, size=(3, dimX, dimY))
    fields[1][rng.uniform(size=(dimX, dimY)) < 0.01] = np.nan              # NaN source cells propagate
    rg = Regridder(x, y, xr, yr)
    out = rg(h2d(fields.reshape(3, -1))).cpu().numpy()
    for f in range(3):
        ref = oi.regrid(x, y, fields[f], xr, yr)
        assert np.array_equal(np.isnan(out[f]), np.isnan(ref))
        ok = ~np.isnan(ref)
        assert ok.sum() > 2000
        assert np.max(np.abs(out[f][ok] - ref[ok])) <= 1e-12


def test_projection_round_trip_and_grid_size():
    from seaiceextentforecasting_b200.ingest import make_npstere_grid, polar_stereo
    rng = np.random.default_rng(1)
    lon = rng.uniform(-180, 180, 1000)
    lat = rng.uniform(40, 89.9, 1000)
    x, y = polar_stereo(lon, lat, 360.0, 1.5e6, -2e5)
    lo, la = polar_stereo(x, y, 360.0, 1.5e6, -2e5, inverse=True)
    assert np.max(np.abs(la - lat)) < 1e-9
    assert np.max(np.abs(((lo - lon + 180) % 360) - 180)) < 1e-9
    lonr, latr, xr, yr, _ = make_npstere_grid(65, 360, 1e5)
    assert xr.shape == (57, 57) and xr.dtype == np.float64
    assert abs(latr[28, 0] - 65.0) < 0.7 and latr.max() > 89.0               # edge mid-points touch the 65N circle

"""Ingest: the steps of `readNSIDC` (north/September1st.py:72-139) that feed the hot path -- NSIDC binary decode,
daily -> monthly mean, polar-hole fill and the linear regrid of the 25 km field onto the analysis grid -- on the GPU
(csrc/ingest.cu), plus the host-side geometry the reference gets from pyproj / scipy.

* `polar_stereo` / `make_npstere_grid`: the spherical polar-stereographic projection the reference configures
  (`+proj=stere +R=6370997 +lat_ts=90 +lat_0=90`, north/September1st.py:19-41) in closed form.  pyproj is not
  installed in this image, so the projection itself is UNPINNED against the reference (it is checked against its own
  inverse and the grid size the scripts rely on: `make_npstere_grid(65, 360, 1e5)` -> 57 x 57).
* `Regridder`: `scipy.interpolate.griddata(points, values, targets, 'linear')` = Delaunay triangulation + barycentric
  interpolation.  The triangulation (scipy.spatial.Delaunay, once per grid pair) and the weights are host work; the
  interpolation of every field is the device SpMV `sie_regrid_linear`.
There is no CPU fallback: the device calls raise without a GPU / the shared library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .engine import _ptr, _stream, h2d, require_cuda

R_SPHERE = 6370997.0
NSIDC_HEADER = 300          # bytes (north/September1st.py:100-101: struct.unpack_from(..., offset=300))
NSIDC_SHAPE = (448, 304)    # north/September1st.py:73-74


def polar_stereo(lon, lat, lon_0, x_0=0.0, y_0=0.0, inverse=False):
    """North-polar stereographic on a sphere, true scale at the pole (`+lat_ts=90 +lat_0=90 +R=6370997`).
    forward: (lon, lat) degrees -> (x, y) metres; inverse: (x, y) -> (lon, lat)."""
    if not inverse:
        lam = np.radians(np.asarray(lon, dtype=np.float64) - lon_0)
        rho = 2.0 * R_SPHERE * np.tan(np.pi / 4.0 - np.radians(np.asarray(lat, dtype=np.float64)) / 2.0)
        return rho * np.sin(lam) + x_0, -rho * np.cos(lam) + y_0
    x = np.asarray(lon, dtype=np.float64) - x_0
    y = np.asarray(lat, dtype=np.float64) - y_0
    rho = np.hypot(x, y)
    la = 90.0 - 2.0 * np.degrees(np.arctan(rho / (2.0 * R_SPHERE)))
    lo = lon_0 + np.degrees(np.arctan2(x, -y))
    return lo, la


def make_npstere_grid(boundinglat, lon_0, grid_res=25e3):
    """`make_npstere_grid` of the scripts (north/September1st.py:19-41): square grid inscribed in the `boundinglat`
    circle.  Returns lon, lat, x, y, (x_0, y_0) with x, y the float32-index products the reference builds."""
    y_ = polar_stereo(lon_0, boundinglat, lon_0)[1]
    llcrnrlat = polar_stereo(np.sqrt(2.0) * y_, 0.0, lon_0, inverse=True)[1]
    llx, lly = polar_stereo(lon_0 - 45.0, llcrnrlat, lon_0)
    x_0, y_0 = -llx, -lly
    urx, ury = polar_stereo(lon_0 + 135.0, llcrnrlat, lon_0, x_0, y_0)
    nx = int(urx / grid_res) + 1
    ny = int(ury / grid_res) + 1
    dx = urx / (nx - 1)
    dy = ury / (ny - 1)
    x = dx * np.indices((ny, nx), np.float32)[1, :, :]
    y = dy * np.indices((ny, nx), np.float32)[0, :, :]
    lon, lat = polar_stereo(x, y, lon_0, x_0, y_0, inverse=True)
    return lon, lat, x, y, (x_0, y_0)


class Regridder:
    """griddata(src_points, values, dst_points, 'linear') with the triangulation done once."""

    def __init__(self, src_x, src_y, dst_x, dst_y):
        from scipy.spatial import Delaunay
        require_cuda()
        self.lib = _lib.load()
        pts = np.column_stack([np.asarray(src_x, dtype=np.float64).ravel(), np.asarray(src_y, dtype=np.float64).ravel()])
        self.C = pts.shape[0]
        self.dst_shape = np.asarray(dst_x).shape
        xi = np.column_stack([np.asarray(dst_x, dtype=np.float64).ravel(), np.asarray(dst_y, dtype=np.float64).ravel()])
        self.Ct = xi.shape[0]
        tri = Delaunay(pts)
        simplex = tri.find_simplex(xi)
        inside = simplex >= 0
        sidx = np.where(inside, simplex, 0)
        T = tri.transform[sidx]                                  # (Ct, 3, 2): inverse affine map + offset row
        d = xi - T[:, 2, :]
        c0 = T[:, 0, 0] * d[:, 0] + T[:, 0, 1] * d[:, 1]         # scipy _barycentric_coordinates, same order
        c1 = T[:, 1, 0] * d[:, 0] + T[:, 1, 1] * d[:, 1]
        c2 = 1.0 - c0 - c1                                       # (1 - c0) - c1, as scipy accumulates it
        self.bary_host = np.ascontiguousarray(np.stack([c0, c1, c2], axis=1))
        vert = tri.simplices[sidx].astype(np.int32)
        vert[~inside] = -1
        self.vert_host = np.ascontiguousarray(vert)
        self.vert = h2d(self.vert_host)
        self.bary = h2d(self.bary_host)

    def __call__(self, fields_dev):
        """fields_dev: device float64 [F, C] (or [C]) -> device [F, *dst_shape]."""
        one = fields_dev.dim() == 1
        src = fields_dev.reshape(1, -1) if one else fields_dev.reshape(fields_dev.shape[0], -1)
        assert src.shape[1] == self.C and src.is_contiguous()
        F = src.shape[0]
        dst = torch.empty((F, self.Ct), dtype=torch.float64, device="cuda")
        rc = self.lib.sie_regrid_linear(_ptr(src), F, self.C, _ptr(self.vert), _ptr(self.bary), self.Ct, _ptr(dst),
                                        _stream())
        _lib.check(rc, "sie_regrid_linear")
        out = dst.reshape((F,) + tuple(self.dst_shape))
        return out[0] if one else out


def nsidc_monthly(files_bytes, C_cells=NSIDC_SHAPE[0] * NSIDC_SHAPE[1], header=NSIDC_HEADER):
    """Raw NSIDC `.bin` images of one month (one monthly file, or the daily near-real-time files,
    north/September1st.py:86-127) -> device [C] monthly concentration, flag values (> 1) -> NaN."""
    require_cuda()
    lib = _lib.load()
    bufs = [np.frombuffer(b, dtype=np.uint8) for b in files_bytes]
    stride = header + C_cells
    host = np.empty((len(bufs), stride), dtype=np.uint8)
    for i, b in enumerate(bufs):
        assert b.size >= stride, "file shorter than header + grid"
        host[i] = b[:stride]
    dev = h2d(host)
    monthly = torch.empty(C_cells, dtype=torch.float64, device="cuda")
    rc = lib.sie_nsidc_monthly(_ptr(dev), len(bufs), stride, header, C_cells, _ptr(monthly), _stream())
    _lib.check(rc, "sie_nsidc_monthly")
    return monthly


def polar_hole_fill(monthly_dev, lat_dev, hole):
    """`phole = nanmean(monthly[(lat > hole-0.5) & (lat < hole)])`, `filled = where(lat >= hole-0.5, phole, monthly)`
    (north/September1st.py:129-136).  Returns (filled device [C], phole device scalar)."""
    lib = _lib.load()
    Cn = monthly_dev.numel()
    filled = torch.empty(Cn, dtype=torch.float64, device="cuda")
    phole = torch.empty(1, dtype=torch.float64, device="cuda")
    scratch = torch.empty(Cn, dtype=torch.float64, device="cuda")
    rc = lib.sie_polar_hole_fill(_ptr(monthly_dev), _ptr(lat_dev), C.c_double(float(hole)), Cn, _ptr(filled),
                                 _ptr(phole), _ptr(scratch), _stream())
    _lib.check(rc, "sie_polar_hole_fill")
    return filled, phole


def polar_hole_latitude(year):
    """north/September1st.py:129-134."""
    if year <= 1987:
        return 84.5
    if year < 2008:
        return 87.2
    return 89.2

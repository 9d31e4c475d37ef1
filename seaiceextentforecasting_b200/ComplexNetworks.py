"""Drop-in for the reference's `ComplexNetworks.py` (class `Network`, ComplexNetworks.py:11-326).

Same constructor, same three methods invoked the way every reference script invokes them
(`CN.Network.tau(network, 0.01)`, `CN.Network.area_level(network, latlon_grid=...)`,
`CN.Network.intra_links(network, area=... | lat=...)`), same result attributes with the same Python
shapes (`tau` float shadowing the method, `nodes` int64 (1,N), `corrs` (N,dimX,dimY) with NaN scatter,
`V`/`A` the same dict of `[i, j]` lists in insertion order, `anomaly`/`links`/`strength` dicts keyed like
`V`, `strengthmap` (dimX,dimY)).  The arithmetic runs in libsie_b200's CUDA kernels; nothing is computed
on the CPU and a missing library or GPU raises.

The reference raises `IndexError` when `data` holds no NaN cell (ComplexNetworks.py:50-51) and `ValueError`
when fewer than two areas survive (:212/:278, with `V` already populated); so does this class.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import NetworkBatch, h2d, r_crit_ttest, require_cuda


class Network:
    # Largest correlation matrix (bytes) the class keeps on the device; above it `tau` runs the tau-only correlation
    # pass and `area_level` recomputes the correlations it needs from the unit-norm rows (a 25 km grid: 32-148 GB; the
    # stored matrix makes domain growth ~10x faster).  None: half of the device memory that is free when the engine is
    # created (a 63 574-node 25 km network, 32 GB, is stored on a 180-GB B200).
    max_matrix_bytes = None

    def __init__(self, data, V={}, A={}, corrs=[], tau=0, nodes=[], unavail=[], anomaly={}, links={},
                 strength={}, strengthmap=[]):
        """`data`: de-trended (zero-mean) series, (x, y, t) or (lat, lon, t) -- ComplexNetworks.py:12-29."""
        self.data = data
        self.dimX, self.dimY, self.dimT = self.data.shape
        self.V = V
        self.A = A
        self.corrs = corrs
        self.tau = tau
        self.nodes = nodes
        self.unavail = unavail
        self.anomaly = anomaly
        self.links = links
        self.strength = strength
        self.strengthmap = strengthmap
        self._eng = None

    # -- device state ------------------------------------------------------------------------------
    def _engine(self):
        if self._eng is None:
            require_cuda()
            data = np.ascontiguousarray(self.data, dtype=np.float64)
            n_upper = int((~np.isnan(data).all(axis=2)).sum())
            ldn = max(128, (n_upper + 127) // 128 * 128)
            limit = self.max_matrix_bytes
            if limit is None:
                limit = torch.cuda.mem_get_info()[0] // 2
            self._eng = NetworkBatch(self.dimX, self.dimY, self.dimT, 1, latlon=False, n_upper=n_upper,
                                     keep_R=8 * ldn * ldn <= limit)
            self._fields = h2d(data.reshape(1, self.dimX * self.dimY, self.dimT))
            self._job_field = torch.zeros(1, dtype=torch.int32, device="cuda")
            self._job_T = torch.full((1,), self.dimT, dtype=torch.int32, device="cuda")
        return self._eng

    # -- ComplexNetworks.py:31-47 -------------------------------------------------------------------
    def tau(self, significance=0.01):
        eng = self._engine()
        eng.detrend_zscore(self._fields, self._job_field, self._job_T, do_detrend=False)
        rc = torch.tensor([r_crit_ttest(self.dimT, float(significance))], dtype=torch.float64, device="cuda")
        eng.corr_tau(rc, store_R=eng.R is not None)
        N = int(eng.n_nodes.cpu()[0])
        if int(eng.status.cpu()[0]) == _lib.SIE_JOB_CAPACITY:
            raise _lib.SieError("node capacity exceeded")
        self._N = N
        self.nodes = np.atleast_2d(eng.node_cell[0, :N].cpu().numpy().astype(np.int64))
        self.corrs = _LazyCorrs(self, N)
        self.tau = float(eng.tau.cpu()[0])

    def correlation_matrix(self):
        """Dense N x N correlation matrix with NaN diagonal (the reference's `R` before the scatter)."""
        if self._eng.R is None:
            return self.correlation_rows(np.arange(self._N))
        U = torch.triu(self._eng.R[0, :self._N, :self._N], diagonal=1)      # only the upper triangle is stored
        R = (U + U.T).cpu().numpy()
        np.fill_diagonal(R, np.nan)
        return R

    def correlation_rows(self, rows):
        """Rows `rows` of the correlation matrix, (len(rows), N): read from the stored matrix, or recomputed from the
        unit-norm rows on the device (sie_corr_rows) when the matrix is not kept."""
        rows = np.atleast_1d(np.asarray(rows, dtype=np.int64))
        eng, N = self._eng, self._N
        if eng.R is not None:                  # element (r, c) lives at R[min(r, c)][max(r, c)]
            r = torch.from_numpy(rows).cuda()[:, None]
            c = torch.arange(N, device="cuda")[None, :]
            out = eng.R[0][torch.minimum(r, c), torch.maximum(r, c)]
            out[r == c] = float("nan")
            return out.cpu().numpy()
        out = torch.empty((len(rows), N), dtype=torch.float64, device="cuda")
        for r0 in range(0, len(rows), 32768):
            part = torch.from_numpy(rows[r0:r0 + 32768].astype(np.int32)).cuda()
            rc = eng.lib.sie_corr_rows(eng.z[0].data_ptr(), part.data_ptr(), len(part), N, self.dimT, eng.Tp,
                                       out[r0:].data_ptr(), N, torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "sie_corr_rows")
        return out.cpu().numpy()

    # -- ComplexNetworks.py:49-278 ------------------------------------------------------------------
    def area_level(self, latlon_grid=False):
        eng = self._engine()
        eng.latlon = bool(latlon_grid)
        eng.area_level()
        status, V = eng.areas_to_host()[0]
        if status == _lib.SIE_JOB_NO_NAN_CELL:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # :51
        if status == _lib.SIE_JOB_CAPACITY:
            raise _lib.SieError("area capacity exceeded")
        self.V = V
        self.A = self.V                                                               # `V is A`
        self.unavail = [cell for key in V for cell in V[key]]
        if status == _lib.SIE_JOB_FEW_AREAS:
            raise ValueError("max() arg is an empty sequence")                        # :212 / :278

    # -- ComplexNetworks.py:283-326 -----------------------------------------------------------------
    def intra_links(self, area=None, lat=None):
        eng = self._engine()
        if lat is not None:
            scale = np.sqrt(np.cos(np.radians(lat)))
        elif area is not None:
            scale = np.sqrt(area)
        else:
            scale = np.ones((self.dimX, self.dimY))
        scale = np.broadcast_to(np.asarray(scale, dtype=np.float64), (self.dimX, self.dimY))
        eng.intra_links(h2d(scale.reshape(-1)))
        nA = int(eng.n_areas.cpu()[0])
        keys = [int(k) for k in eng.area_key[0, :nA].cpu().numpy()]
        anom = eng.anomaly[0, :nA, :self.dimT].cpu().numpy()
        links = eng.links[0, :nA, :nA].cpu().numpy()
        strength = eng.strength[0, :nA].cpu().numpy()
        self.anomaly = {k: anom[a].copy() for a, k in enumerate(keys)}
        self.links = {k: [float(x) if a2 != a else 0 for a2, x in enumerate(links[a])] for a, k in enumerate(keys)}
        self.strength = {k: strength[a] for a, k in enumerate(keys)}
        self.strengthmap = eng.strengthmap[0].cpu().numpy().reshape(self.dimX, self.dimY)


class _LazyCorrs:
    """`corrs` is an (N, dimX, dimY) NaN-scattered copy of R in the reference (ComplexNetworks.py:36-39):
    58 MB at 57x57 and impossible at 25 km, so it is materialised from the device matrix on first use."""

    def __init__(self, net, N):
        self._net, self._N, self._arr = net, N, None
        self.shape = (N, net.dimX, net.dimY)

    def _materialise(self):
        if self._arr is None:
            net = self._net
            R = net.correlation_matrix()
            full = np.full((self._N, net.dimX * net.dimY), np.nan)
            full[:, net.nodes[0]] = R
            self._arr = full.reshape(self.shape)
        return self._arr

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        if self._arr is None and self._net._eng.R is None:
            # no stored matrix (large grid): serve the requested node rows without materialising N x dimX x dimY
            first, rest = (idx[0], idx[1:]) if isinstance(idx, tuple) else (idx, ())
            rows = np.arange(self._N)[first]
            net = self._net
            R = net.correlation_rows(rows)
            full = np.full((R.shape[0], net.dimX * net.dimY), np.nan)
            full[:, net.nodes[0]] = R
            full = full.reshape((R.shape[0], net.dimX, net.dimY))
            if np.ndim(rows) == 0:
                full = full[0]
            return full[rest] if rest else full
        return self._materialise()[idx]

    def __len__(self):
        return self._N


CN = None  # `from ComplexNetworks import CN` (north/June1st.py:197) resolves to this module, set below


def _self_module():
    import sys
    return sys.modules[__name__]


CN = _self_module()

"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of `ComplexNetworks.Network`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package.  The product path (`seaiceextentforecasting_b200`) never does.

It restates /root/reference/ComplexNetworks.py (class `Network`, lines 11-326) and calls the *same*
third-party routines the reference calls for its arithmetic (`np.corrcoef`, `scipy.stats.t.sf`,
`np.nanmean`/`np.sum` pairwise summation, `scipy.stats.pearsonr`, `np.std`), so floating-point results
are those of the reference under the numpy/scipy installed in this image (numpy 2.3.5 / scipy 1.18.1;
the reference pins no versions).  What is restated is the control flow: the reference's Python-list
membership scans (`[i,j] in self.unavail`, O(n) each) become bitmap lookups, which changes nothing
observable.

Pinning: the reference has no tests and no golden vectors.  This restatement is
pinned against the reference itself: `tests/golden/make_golden.py` imports
/root/reference/ComplexNetworks.py in the authoring container, runs it on seeded synthetic grids and
commits its outputs; `tests/test_oracle_golden.py` checks this module against those fixtures
bit-for-bit (V keys, cell lists and order, nodes, tau, anomaly) on every run.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy import stats

_DIRS = ((-1, 0), (1, 0), (0, -1), (0, 1))  # up, down, left, right: ComplexNetworks.py:54-77


def _nanmean_1d(values):
    """`np.nanmean(list)` exactly as the reference calls it (ComplexNetworks.py:113, :250, :252)."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        return np.nanmean(values)


class Network:
    """Same constructor/attributes as ComplexNetworks.py:12-29 (mutable defaults kept on purpose)."""

    def __init__(self, data, V={}, A={}, corrs=[], tau=0, nodes=[], unavail=[], anomaly={}, links={},
                 strength={}, strengthmap=[]):
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

    # ------------------------------------------------------------------ tau: ComplexNetworks.py:31-47
    def tau(self, significance=0.01, keep_corrs=True):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            ID = np.where(np.abs(np.nanmax(self.data, 2)) > 0)          # :32
        N = np.shape(ID)[1]
        R = np.corrcoef(self.data[ID])                                   # :34
        np.fill_diagonal(R, np.nan)                                      # :35
        self.nodes = np.atleast_2d(ID[0] * self.dimY + ID[1])           # :37
        self.R = R                                                       # dense N x N (oracle-only handle)
        if keep_corrs:
            corrs = np.full((N, self.dimX * self.dimY), np.nan)         # :36, :38-39 (scatter)
            corrs[:, self.nodes[0]] = R
            self.corrs = corrs.reshape(N, self.dimX, self.dimY)
        df = self.dimT - 2                                               # :41
        Rp = R[R >= 0]                                                   # :42
        with np.errstate(divide="ignore", invalid="ignore"):
            Tt = Rp * np.sqrt(df / (1 - Rp ** 2))                        # :43
        P = stats.t.sf(Tt, df)                                           # :44
        Rp = Rp[P < significance]                                        # :45
        self.tau = np.mean(Rp)                                           # :47

    # ------------------------------------------------------------- area_level: ComplexNetworks.py:49-278
    def area_level(self, latlon_grid=False):
        X, Y = self.dimX, self.dimY
        ids = np.where(np.isnan(self.data))                              # :50
        i_nan = ids[0][0]                                                # :51 (IndexError if no NaN cell)
        j_nan = ids[1][0]
        R = self.R
        tau = self.tau
        node_of = np.full(X * Y, -1, dtype=np.int64)
        node_of[self.nodes[0]] = np.arange(self.nodes.shape[1])
        taken = np.zeros(X * Y, dtype=bool)

        def cell_neighbours(i, j):
            """gen_cell_neighbours :53-78 -> flat cell id or -1 (the NaN sentinel cell)."""
            out = []
            for d, (di, dj) in enumerate(_DIRS):
                a, b = i + di, j + dj
                if 0 <= a < X and 0 <= b < Y:
                    out.append(-1 if taken[a * Y + b] else a * Y + b)
                elif latlon_grid and d == 2 and 0 <= a < X:
                    out.append(i * Y + (Y - 1))                          # :64 wrap, `unavail` not consulted
                elif latlon_grid and d == 3 and 0 <= a < X:
                    out.append(i * Y + 0)                                # :72
                else:
                    out.append(-1)
            return out

        # ---- step 1 (:154-196)
        V = {}
        self.A = {}
        k = 0
        for i in range(X):
            for j in range(Y):
                c0 = i * Y + j
                ID = node_of[c0]
                if ID < 0 or taken[c0]:
                    continue
                nei = cell_neighbours(i, j)
                cs = [R[ID, node_of[c]] if (c >= 0 and node_of[c] >= 0) else np.nan for c in nei]
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    nei_max = np.nanmax(cs)                              # :172
                if not (nei_max > tau):                                  # :174
                    continue
                s = nei[int(np.where(np.asarray(cs) == nei_max)[0][0])]  # :175-181 first max
                if taken[s]:                                             # :182 (lat-lon wrap only)
                    continue
                area = [c0, s]
                taken[c0] = True
                taken[s] = True
                # expand :120-152
                while True:
                    cand = []
                    arr = np.asarray(area)
                    ai, aj = arr // Y, arr % Y
                    for di, dj in _DIRS:                                 # gen_area_neighbours :80-94
                        a, b = ai + di, aj + dj
                        ok = (a >= 0) & (a < X) & (b >= 0) & (b < Y)
                        cc = (a * Y + b)[ok]
                        cc = cc[~taken[cc]]
                        cand.append(cc[node_of[cc] >= 0])                # :100 node filter
                    cand = np.concatenate(cand)                          # direction-major, duplicates kept
                    if cand.size == 0:
                        break
                    G = R[np.ix_(node_of[cand], node_of[arr])]           # :109-112 (area order)
                    R_mean = [_nanmean_1d(G[q]) for q in range(G.shape[0])]   # :113
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", RuntimeWarning)
                        Rmax = np.nanmax(R_mean)                         # :116
                    if not (Rmax > tau):                                 # :134
                        break
                    m = int(cand[int(np.where(np.asarray(R_mean) == Rmax)[0][0])])   # :135-142
                    area.append(m)
                    taken[m] = True
                V[k] = area
                k += 1

        # ---- step 2 (:200-265)
        taken[:] = False
        label = np.full(X * Y, -1, dtype=np.int64)
        for key, cells in V.items():
            label[cells] = key
        final = set()
        while True:
            best, best_n = None, -1
            for key, cells in V.items():                                 # :207-212, first max in dict order
                n = 0 if key in final else len(cells)
                if n > best_n:
                    best, best_n = key, n
            if best is None:
                raise ValueError("max() arg is an empty sequence")      # :212 with no areas at all
            if best_n == 0:
                break
            order = []
            seen = set()
            keypos = {key: p for p, key in enumerate(V)}
            for c in V[best]:                                            # :217
                nei = cell_neighbours(c // Y, c % Y)
                hits = []
                for d, cnb in enumerate(nei):
                    if cnb >= 0 and label[cnb] >= 0 and label[cnb] != best:
                        hits.append((keypos[label[cnb]], d, int(label[cnb])))
                for _, _, key in sorted(hits):                           # `for k in self.V: for nei in nei_list`
                    if key not in seen:
                        seen.add(key)
                        order.append(key)
            stats_ = []
            for key in order:                                            # :224-253
                hyp = node_of[np.asarray(V[best] + V[key])]
                n = hyp.size
                G = R[np.ix_(hyp, hyp)]
                r = [_nanmean_1d(G[p, p + 1:]) for p in range(n)]        # last one = nanmean([]) = nan
                stats_.append(_nanmean_1d(r))                            # :253
            chosen = None
            if order:
                cur = stats_[0]                                          # max(dict.items(), key=itemgetter(1))
                chosen = 0
                for q in range(1, len(order)):
                    if stats_[q] > cur:
                        cur = stats_[q]
                        chosen = q
                if not (cur > tau):                                      # :257
                    chosen = None
            if chosen is not None:
                key = order[chosen]
                cells = V.pop(key)                                       # :259
                V[best] = V[best] + cells                                # :260-261 (dict position of best kept)
                label[cells] = best
            else:
                final.add(best)                                          # :262-265
                taken[V[best]] = True
        self.V = {int(key): [[int(c // Y), int(c % Y)] for c in cells] for key, cells in V.items()}
        self.A = self.V                                                  # `V is A` in the reference
        self.unavail = [cell for key in self.V for cell in self.V[key]]
        if len(self.V) < 2:                                              # :269-278 tail
            raise ValueError("max() arg is an empty sequence")

    # ------------------------------------------------------------ intra_links: ComplexNetworks.py:283-326
    def intra_links(self, area=None, lat=None):
        self.anomaly = {}
        self.links = {}
        self.strength = {}
        self.strengthmap = np.zeros((self.dimX, self.dimY)) * np.nan
        if lat is not None:
            scale = np.sqrt(np.cos(np.radians(lat)))
        elif area is not None:
            scale = np.sqrt(area)
        else:
            scale = np.ones((self.dimX, self.dimY))
        for A in self.V:                                                 # :303-307
            temp_array = np.zeros(self.data.shape) * np.nan
            for cell in self.V[A]:
                temp_array[cell[0], cell[1], :] = np.multiply(self.data[cell[0], cell[1], :],
                                                              scale[cell[0], cell[1]])
            self.anomaly[A] = np.nansum(temp_array, axis=(0, 1))
        for A in self.anomaly:                                           # :309-316
            sdA = np.std(self.anomaly[A])
            for A2 in self.anomaly:
                sdA2 = np.std(self.anomaly[A2])
                if A2 != A:
                    self.links.setdefault(A, []).append(
                        stats.pearsonr(self.anomaly[A], self.anomaly[A2])[0] * (sdA * sdA2))
                else:
                    self.links.setdefault(A, []).append(0)
        for A in self.links:                                             # :318-326
            self.strength[A] = np.nansum([abs(link) for link in self.links[A]])
            for cell in self.V[A]:
                self.strengthmap[cell[0], cell[1]] = self.strength[A]

"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the ingest steps of `readNSIDC`
(north/September1st.py:72-139) with the same numpy / scipy calls the reference makes: `struct.unpack_from` + `/250`
(:100-104, :121-127), `np.nanmean(daily, 2)` (:105), `monthly[monthly>1] = nan` (:128), the polar-hole mean and
`np.ma.where` fill (:129-136) and `scipy.interpolate.griddata(..., 'linear')` (:137-138).  Pinned on the reference by
construction (it IS the reference's call sequence on the same arrays); the projection (pyproj) is not restated here.

Only `tests/` may import this module."""
from __future__ import annotations

import struct

import numpy as np
from scipy.interpolate import griddata


def decode_monthly(files_bytes, dimX, dimY):
    s = "%dB" % (int(dimX * dimY),)
    if len(files_bytes) == 1:
        z = struct.unpack_from(s, files_bytes[0], offset=300)
        monthly = (np.array(z).reshape((dimX, dimY))) / 250
    else:
        daily = np.zeros((dimX, dimY, len(files_bytes))) * np.nan
        for f, contents in enumerate(files_bytes):
            z = struct.unpack_from(s, contents, offset=300)
            daily[:, :, f] = (np.array(z).reshape((dimX, dimY))) / 250
        monthly = np.nanmean(daily, 2)
    monthly[monthly > 1] = np.nan
    return monthly


def hole_fill(monthly, lat, hole):
    phole = np.nanmean(monthly[(lat > hole - 0.5) & (lat < hole)])
    filled = np.ma.where((lat >= hole - 0.5), phole, monthly)
    return np.asarray(filled), phole


def regrid(x, y, filled, xr, yr):
    return griddata((x.ravel(), y.ravel()), filled.ravel(), (xr, yr), 'linear')

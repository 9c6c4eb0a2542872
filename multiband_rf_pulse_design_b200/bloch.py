"""Host-side mirror of the reference Bloch-simulator MEX interface.

    [mx,my,mz] = blochC(b1, gr, tp, t1, t2, df, dp, mode, mx, my, mz)     (bloch.m:4)

Same positional arguments, defaults and output shapes as the reference gateway
(bloch_simulation/blochC.c:514-927; blochH.c differs only in GAMMA).  All argument
handling is done inside libmbrf (``mbrf_bloch``), exactly as a MEX gateway would use it;
this file only converts numpy arrays to the split real/imag, column-major planes the
MEX API hands over.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import c_double_p, check, lib

GAMMA_C13 = 6726.1   # blochC.c:6
GAMMA_H1 = 26754.0   # blochH.c:6


def _plane(a):
    """Column-major float64 plane of `a`, as mxGetPr would return it."""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.ravel(order="F"))


def _p(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _mn(a):
    """(M, N) of a MATLAB value: scalars are 1x1, 1-D numpy arrays are treated as row vectors."""
    a = np.asarray(a)
    if a.ndim == 0:
        return 1, 1
    if a.ndim == 1:
        return 1, a.shape[0]
    if a.ndim == 2:
        return a.shape
    raise ValueError("at most 2-D inputs")


def bloch(b1, gr, tp, t1, t2, df, dp, mode=0, mx=None, my=None, mz=None, *, gamma=GAMMA_C13, out=None):
    """[mx,my,mz] = bloch(b1,gr,tp,t1,t2,df,dp,mode,mx,my,mz) on the GPU.

    `out`, if given, is a tuple of three preallocated float64 arrays (e.g. pinned) of
    ntout*npos*nf elements that receive the result.
    Returns three arrays shaped like the reference's outputs (blochC.c:880-904).
    """
    b1 = np.asarray(b1)
    ntime = b1.size                                        # blochC.c:571
    b1r = _plane(b1.real)
    b1i = _plane(b1.imag) if np.iscomplexobj(b1) else None  # blochC.c:576-587
    grp = _plane(gr)
    tpp = _plane(tp)
    dfp = _plane(df)
    dpp = _plane(dp)
    pm, pn = _mn(dp)
    nf = dfp.size
    npos = pm if pn in (2, 3) else pm * pn                 # blochC.c:701-758
    mode = int(mode)
    ntout = ntime if (mode & 2) else 1
    total = ntout * npos * nf
    m0 = [None, None, None]
    n_m0 = 0
    if mx is not None and my is not None and mz is not None:   # nrhs > 10, blochC.c:820
        m0 = [_plane(mx), _plane(my), _plane(mz)]
        n_m0 = m0[0].size if (m0[0].size == m0[1].size == m0[2].size) else -1
    if out is None:
        out = tuple(np.empty(max(total, 1), dtype=np.float64) for _ in range(3))
    else:
        for o in out:
            if o.dtype != np.float64 or o.size < total or not o.flags.c_contiguous:
                raise ValueError("out arrays must be contiguous float64 with ntout*npos*nf elements")
    dims = (C.c_int * 4)()
    check(lib().mbrf_bloch(_p(b1r), _p(b1i), ntime, _p(grp), grp.size, _p(tpp), tpp.size,
                           float(np.asarray(t1).ravel()[0]), float(np.asarray(t2).ravel()[0]),
                           _p(dfp), nf, _p(dpp), pm, pn, mode,
                           _p(m0[0]), _p(m0[1]), _p(m0[2]), n_m0,
                           _p(out[0]), _p(out[1]), _p(out[2]), dims, float(gamma)))
    shape = tuple(dims[i] for i in range(dims[3]))
    return tuple(o.ravel()[:total].reshape(shape, order="F") for o in out)


def blochC(b1, gr, tp, t1, t2, df, dp, mode=0, mx=None, my=None, mz=None, **kw):
    """13C Bloch simulator (bloch_simulation/blochC.c)."""
    return bloch(b1, gr, tp, t1, t2, df, dp, mode, mx, my, mz, gamma=GAMMA_C13, **kw)


def blochH(b1, gr, tp, t1, t2, df, dp, mode=0, mx=None, my=None, mz=None, **kw):
    """1H Bloch simulator (bloch_simulation/blochH.c)."""
    return bloch(b1, gr, tp, t1, t2, df, dp, mode, mx, my, mz, gamma=GAMMA_H1, **kw)


def blochsimfz(b1real, b1imag, xgrad, ygrad, zgrad, tsteps, t1, t2, dfreq, dxpos, dypos, dzpos,
               mx, my, mz, mode, gamma=GAMMA_C13):
    """The reference's inner C entry (blochC.c:422-426): mx/my/mz are updated IN PLACE."""
    arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64)
            for a in (b1real, b1imag, xgrad, ygrad, zgrad, tsteps, dfreq, dxpos, dypos, dzpos)]
    b1r, b1i, gx, gy, gz, dt, df, dx, dy, dz = arrs
    for o in (mx, my, mz):
        if not (isinstance(o, np.ndarray) and o.dtype == np.float64 and o.flags.c_contiguous):
            raise ValueError("mx,my,mz must be contiguous float64 arrays (updated in place)")
    check(lib().mbrf_blochsimfz(_p(b1r), _p(b1i), _p(gx), _p(gy), _p(gz), _p(dt), b1r.size, float(t1), float(t2),
                                _p(df), df.size, _p(dx), _p(dy), _p(dz), dx.size, _p(mx), _p(my), _p(mz),
                                int(mode), float(gamma)))
    return mx, my, mz

"""Host-side mirror of the reference forward-SLR interface.

    [alpha, beta] = abrx(rf, g, x [, y])     rf_tools/mex5/abrx.c:35-79 (MEX)
    [a b]         = abrm(rf, [g,] x [, y])   rf_tools/abrm.m
    [a, b]        = abr(rf, [g,] x [, y])    rf_tools/abr.m (abrx, then b = -conj(b))
"""
from __future__ import annotations

import numpy as np

from ._lib import c_double_p, check, lib

ABRX, ABRM, ABR = 0, 1, 2


def _p(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _f(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _run(rf, g, x, y, convention, use_gy):
    rf = np.asarray(rf).ravel()
    g = np.asarray(g).ravel()
    ns = rf.size                                           # abrx.c:43: max(M,N) of a vector
    if g.size != ns:                                       # abrx.c:44-45
        raise ValueError("rf and gradient vectors are of different lengths")
    rfr = _f(rf.real)
    rfi = _f(rf.imag) if np.iscomplexobj(rf) else None     # abrx.c:51,92 (mxGetPi NULL for real)
    gx = _f(g.real)
    gy = _f(g.imag) if (use_gy and np.iscomplexobj(g)) else None   # abrx.c:50
    xv = _f(x)
    yv = _f(y) if y is not None else None
    nx, ny = xv.size, (yv.size if yv is not None else 1)
    outs = [np.empty(max(nx * ny, 1)) for _ in range(4)]
    check(lib().mbrf_abr(_p(rfr), _p(rfi), _p(gx), _p(gy), ns, _p(xv), nx, _p(yv), ny, convention,
                         *[_p(o) for o in outs]))
    a = (outs[0][:nx * ny] + 1j * outs[1][:nx * ny]).reshape((nx, ny), order="F")
    b = (outs[2][:nx * ny] + 1j * outs[3][:nx * ny]).reshape((nx, ny), order="F")
    return a, b


def abrx(rf, g, x, y=None):
    """[alpha, beta] = abrx(rf, g, x {, y}); outputs are nx-by-ny (abrx.c:59-62)."""
    if y is None:
        # 3-argument call: ny = 1, y = 0 and the y gradient is not read (abrx.c:50,54-57,68)
        return _run(rf, g, x, None, ABRX, use_gy=False)
    return _run(rf, g, x, y, ABRX, use_gy=True)


def abr(rf, g, x=None, y=None):
    """[a, b] = abr(rf, <g,> x <, y>) — Le Roux's convention on beta (abr.m:19-37)."""
    rf = np.asarray(rf)
    if x is None:                                          # nargin == 2, abr.m:23-26
        x = g
        g = np.ones(rf.size) * 2 * np.pi / rf.size
    if y is None:
        return _run(rf, g, x, None, ABR, use_gy=False)
    return _run(rf, g, x, y, ABR, use_gy=True)


def abrm(rf, g, x=None, y=None):
    """[a b] = abrm(rf, [g,] x [, y]) — the .m twin's convention (abrm.m:26-64)."""
    rf = np.asarray(rf)
    if x is None:                                          # nargin == 2, abrm.m:26-29
        x = g
        g = np.ones(rf.size) * 2 * np.pi / rf.size
    if y is None:
        y = np.zeros(1)                                    # abrm.m:30-32
    return _run(rf, g, x, y, ABRM, use_gy=True)


# --------------------------------------------------------------------------------------------
# inverse SLR (the step after the FIR design, dzrf_mb.m:239-240), batched on the GPU — csrc/islr.cu
# --------------------------------------------------------------------------------------------
def _planes(z):
    z = np.atleast_2d(np.asarray(z))
    re = np.ascontiguousarray(z.real, dtype=np.float64)
    im = np.ascontiguousarray(z.imag, dtype=np.float64) if np.iscomplexobj(z) else None
    return z.shape, re, im


def b2a(bc):
    """aca = b2a(bc) — rf_tools/b2a.m:13-28: minimum-phase, minimum-power alpha polynomial of a beta polynomial.
    bc: [n] or a batch [B, n]; n = 2^k <= 1024 (radix-2 transform of length 8n in shared memory) or any other n <= 512 (Bluestein)."""
    shape, br, bi = _planes(bc)
    B, n = shape
    ar = np.empty((B, n)); ai = np.empty((B, n))
    check(lib().mbrf_b2a_batch(_p(br), _p(bi), n, B, _p(ar), _p(ai)))
    out = ar + 1j * ai
    return out[0] if np.ndim(bc) == 1 else out


def ab2rf(ac, bc):
    """rf = ab2rf(ac, bc) — rf_tools/ab2rf.m:12-26 (complex RF version): the hard-pulse RF that produces alpha, beta.
    ac, bc: [n] or batches [B, n], n <= 2048."""
    shape, ar, ai = _planes(ac)
    shape_b, br, bi = _planes(bc)
    if shape != shape_b:
        raise ValueError("ac and bc must have the same shape")
    B, n = shape
    rr = np.empty((B, n)); ri = np.empty((B, n))
    check(lib().mbrf_ab2rf_batch(_p(ar), _p(ai), _p(br), _p(bi), n, B, _p(rr), _p(ri)))
    out = rr + 1j * ri
    return out[0] if np.ndim(ac) == 1 else out


def b2rf(bc):
    """rf = ab2rf(b2a(bc), bc) — dzrf_mb.m:239-240, for one beta polynomial or a batch."""
    return ab2rf(b2a(bc), bc)

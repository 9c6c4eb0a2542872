"""TEST INFRASTRUCTURE ONLY — ctypes access to the CPU oracles.

Two families live here:

* ``oracle/liboracle.so`` — our C restatement (``bloch_oracle.c``, ``slr_oracle.c``).
* ``oracle/_ref/*.so``    — the UNMODIFIED reference C
  (``/root/reference/bloch_simulation/blochC.c``, ``blochH.c``,
  ``rf_tools/mex5/abrx.c``, ``b2rf.c``) compiled against ``oracle/mex_stub/mex.h``
  by ``oracle/Makefile``.  Its ``mexFunction`` is driven through stub mxArrays,
  so gateway behaviour (input normalisation, output shapes, error messages) is
  checked against the real thing, not only the inner loops.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

GAMMA_C13 = 6726.1   # blochC.c:6
GAMMA_H1 = 26754.0   # blochH.c:6

_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always possible) and oracle/_ref (only where the reference tree exists)."""
    need = force or not os.path.exists(os.path.join(HERE, "liboracle.so"))
    ref_present = os.path.isdir("/root/reference/bloch_simulation")
    if ref_present and not os.path.exists(os.path.join(REF_DIR, "libblochC.so")):
        need = True
    if need:
        subprocess.check_call(["make", "-s", "-C", HERE] + (["-B"] if force else []))


def _as(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


_cache: dict = {}


def _lib(path):
    if path not in _cache:
        if not os.path.exists(path):
            build()
        _cache[path] = C.CDLL(path)
    return _cache[path]


def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in
               ("libblochC.so", "libblochH.so", "libabrx.so", "libb2rf.so"))


# ----------------------------------------------------------------------------
# inner ABI: blochsimfz (blochC.c:422-426)
# ----------------------------------------------------------------------------
_BLOCHSIMFZ_ARGS = [_dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double, _dp, C.c_int,
                    _dp, _dp, _dp, C.c_int, _dp, _dp, _dp, C.c_int]


def _prep_sim(b1, gx, gy, gz, dt, df, dx, dy, dz, mode, m0):
    b1 = np.asarray(b1)
    nt = b1.size
    b1r = _as(b1.real).ravel()
    b1i = _as(b1.imag).ravel() if np.iscomplexobj(b1) else np.zeros(nt)
    z = np.zeros(nt)
    gx = _as(gx).ravel() if gx is not None else z
    gy = _as(gy).ravel() if gy is not None else z
    gz = _as(gz).ravel() if gz is not None else z
    dt = _as(dt).ravel()
    if dt.size == 1:
        dt = np.full(nt, dt[0])
    df = _as(df).ravel()
    dx = _as(dx).ravel()
    zp = np.zeros(dx.size)
    dy = _as(dy).ravel() if dy is not None else zp
    dz = _as(dz).ravel() if dz is not None else zp
    ntout = nt if (mode & 2) else 1
    ns = df.size * dx.size
    out = []
    for k in range(3):
        o = np.zeros(ns * ntout)
        if m0 is None:
            o[::ntout] = 1.0 if k == 2 else 0.0
        else:
            o[::ntout] = np.asarray(m0[k], dtype=np.float64).ravel()
        out.append(o)
    return b1r, b1i, gx, gy, gz, dt, nt, df, dx, dy, dz, ntout, out


def blochsimfz_oracle(b1, gx, gy, gz, dt, t1, t2, df, dx, dy=None, dz=None, mode=0, m0=None,
                      gamma=GAMMA_C13):
    """Restatement (liboracle.so).  Returns (mx,my,mz) flat arrays of ntout*npos*nf, index t+ntout*(p+npos*f)."""
    lib = _lib(os.path.join(HERE, "liboracle.so"))
    fn = lib.oracle_blochsimfz
    fn.argtypes = _BLOCHSIMFZ_ARGS + [C.c_double]
    fn.restype = C.c_int
    b1r, b1i, gx, gy, gz, dt, nt, df, dx, dy, dz, ntout, out = _prep_sim(b1, gx, gy, gz, dt, df, dx, dy, dz, mode, m0)
    rc = fn(_ptr(b1r), _ptr(b1i), _ptr(gx), _ptr(gy), _ptr(gz), _ptr(dt), nt, t1, t2, _ptr(df), df.size,
            _ptr(dx), _ptr(dy), _ptr(dz), dx.size, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), mode, gamma)
    assert rc == 0
    return out


def blochsimfz_ref(b1, gx, gy, gz, dt, t1, t2, df, dx, dy=None, dz=None, mode=0, m0=None,
                   nucleus="C-13"):
    """The reference's own blochsimfz (oracle/_ref).  Modes 0 and 2 only (modes 1/3 hit UB at blochC.c:132)."""
    if mode not in (0, 2):
        raise ValueError("reference modes 1/3 are undefined behaviour (blochC.c:132); use blochsimfz_oracle")
    lib = _lib(os.path.join(REF_DIR, "libblochC.so" if nucleus == "C-13" else "libblochH.so"))
    fn = lib.blochsimfz
    fn.argtypes = _BLOCHSIMFZ_ARGS
    fn.restype = C.c_int
    b1r, b1i, gx, gy, gz, dt, nt, df, dx, dy, dz, ntout, out = _prep_sim(b1, gx, gy, gz, dt, df, dx, dy, dz, mode, m0)
    # the reference prints a progress line per 10 % above 40 000 spins (blochC.c:502-503); silence it
    with _quiet_stdout():
        fn(_ptr(b1r), _ptr(b1i), _ptr(gx), _ptr(gy), _ptr(gz), _ptr(dt), nt, t1, t2, _ptr(df), df.size,
           _ptr(dx), _ptr(dy), _ptr(dz), dx.size, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), mode)
    return out


class _quiet_stdout:
    """Redirect C-level stdout to /dev/null (the reference MEX prints banners, blochC.c:565-569)."""

    def __enter__(self):
        import sys
        sys.stdout.flush()
        self._saved = os.dup(1)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 1)

    def __exit__(self, *exc):
        C.CDLL(None).fflush(None)
        os.dup2(self._saved, 1)
        os.close(self._saved)
        os.close(self._null)


# ----------------------------------------------------------------------------
# MEX level: run the reference's mexFunction on stub mxArrays
# ----------------------------------------------------------------------------
class MxArray(C.Structure):
    _fields_ = [("m", C.c_size_t), ("n", C.c_size_t), ("pr", _dp), ("pi", _dp),
                ("ndim", C.c_int), ("dims", C.c_int * 3), ("is_char", C.c_int)]


def _to_mx(a, keep):
    """numpy (<=2-D, MATLAB column-major semantics) -> stub mxArray.  Scalars become 1x1."""
    a = np.asarray(a)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)          # MATLAB row vector
    m, n = a.shape
    re = np.asfortranarray(a.real, dtype=np.float64).ravel(order="F").copy()
    mx = MxArray()
    mx.m, mx.n, mx.ndim = m, n, 2
    mx.dims[0], mx.dims[1], mx.dims[2] = m, n, 1
    mx.pr = _ptr(re)
    keep.append(re)
    if np.iscomplexobj(a):
        im = np.asfortranarray(a.imag, dtype=np.float64).ravel(order="F").copy()
        mx.pi = _ptr(im)
        keep.append(im)
    else:
        mx.pi = None
    return mx


def _from_mx(p):
    mx = p.contents
    cnt = mx.m * mx.n
    dims = [mx.dims[i] for i in range(mx.ndim)]
    re = np.ctypeslib.as_array(mx.pr, shape=(max(cnt, 1),))[:cnt].copy()
    if mx.pi:
        im = np.ctypeslib.as_array(mx.pi, shape=(max(cnt, 1),))[:cnt].copy()
        re = re + 1j * im
    return re.reshape(dims, order="F")


def mex_call(libname, nlhs, *args):
    """Call `mexFunction` of oracle/_ref/<libname> with numpy inputs; returns (outputs, error_message|None)."""
    lib = _lib(libname if os.path.isabs(libname) else os.path.join(REF_DIR, libname))
    lib.mex_stub_call.argtypes = [C.c_int, C.POINTER(C.POINTER(MxArray)), C.c_int, C.POINTER(C.POINTER(MxArray))]
    lib.mex_stub_call.restype = C.c_int
    lib.mex_stub_last_error.restype = C.c_char_p
    lib.mex_stub_destroy.argtypes = [C.POINTER(MxArray)]
    keep: list = []
    ins = [_to_mx(a, keep) for a in args]
    prhs = (C.POINTER(MxArray) * max(len(ins), 1))(*[C.pointer(x) for x in ins])
    plhs = (C.POINTER(MxArray) * max(nlhs, 1))()
    with _quiet_stdout():
        rc = lib.mex_stub_call(nlhs, plhs, len(ins), prhs)
    if rc:
        return None, lib.mex_stub_last_error().decode()
    outs = [_from_mx(plhs[i]) for i in range(nlhs)]
    for i in range(nlhs):
        lib.mex_stub_destroy(plhs[i])
    return outs, None


def bloch_mex_ref(nucleus, *args):
    """[mx,my,mz] = blochC/blochH(b1,gr,tp,t1,t2,df,dp[,mode[,mx,my,mz]]) through the reference gateway."""
    outs, err = mex_call("libblochC.so" if nucleus == "C-13" else "libblochH.so", 3, *args)
    assert err is None, err
    return outs


def abrx_mex_ref(rf, g, x, y=None, nlhs=2):
    """[alpha,beta] = abrx(rf,g,x[,y]) through the reference gateway (abrx.c:35-79).

    The reference keeps `gy` in a file-scope global that is only assigned on 4-argument
    calls (abrx.c:30,50): a 3-argument call after a 4-argument one would read a stale
    pointer.  The library is therefore re-opened for every call so each starts clean.
    """
    path = os.path.join(REF_DIR, "libabrx.so")
    if not os.path.exists(path):
        build()
    import shutil
    import tempfile
    args = (rf, g, x) if y is None else (rf, g, x, y)
    with tempfile.TemporaryDirectory() as td:
        tmp = os.path.join(td, "libabrx_fresh.so")
        shutil.copy(path, tmp)
        try:
            return mex_call(tmp, nlhs, *args)
        finally:
            _cache.pop(tmp, None)


def b2rf_ref(b):
    """rf = b2rf(b) (rf_tools/mex5/b2rf.c): inverse SLR, used only to GENERATE workload pulses."""
    outs, err = mex_call("libb2rf.so", 1, np.asarray(b))
    assert err is None, err
    return outs[0].ravel()


# ----------------------------------------------------------------------------
# forward SLR restatements
# ----------------------------------------------------------------------------
def abrx_oracle(rf, g, x, y=None):
    """Restatement of abrx.c (liboracle.so).  Returns (alpha, beta) complex, shape (nx, ny)."""
    lib = _lib(os.path.join(HERE, "liboracle.so"))
    fn = lib.oracle_abrx
    fn.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp]
    fn.restype = None
    rf = np.asarray(rf).ravel()
    g = np.asarray(g).ravel()
    rfi = _as(rf.real)
    rfq = _as(rf.imag) if np.iscomplexobj(rf) else None
    gx = _as(g.real)
    # abrx.c:50 — mxGetPi(g) is NULL for a real g even on 4-argument calls
    gy = _as(g.imag) if (np.iscomplexobj(g) and y is not None) else None
    x = _as(x).ravel()
    yy = _as(y).ravel() if y is not None else np.zeros(1)
    nx, ny = x.size, (yy.size if y is not None else 1)
    o = [np.zeros(nx * ny) for _ in range(4)]
    fn(_ptr(rfi), _ptr(rfq), _ptr(gx), _ptr(gy), rf.size, _ptr(x), nx, _ptr(yy), ny,
       1 if (y is not None and gy is not None) else 0, *[_ptr(v) for v in o])
    sh = (nx, ny)
    return (o[0] + 1j * o[1]).reshape(sh, order="F"), (o[2] + 1j * o[3]).reshape(sh, order="F")


def abrm_oracle(rf, g, x, y=None):
    """Restatement of abrm.m:26-64 (3/4-argument form).  Returns (a, b) of shape (lx, ly)."""
    lib = _lib(os.path.join(HERE, "liboracle.so"))
    fn = lib.oracle_abrm
    fn.argtypes = [_dp, _dp, _dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _dp, _dp, _dp]
    fn.restype = None
    rf = np.asarray(rf).ravel()
    g = np.asarray(g).ravel()
    rfr = _as(rf.real)
    rfi = _as(rf.imag) if np.iscomplexobj(rf) else None
    gr = _as(g.real)
    gi = _as(g.imag) if np.iscomplexobj(g) else None
    x = _as(x).ravel()
    yy = _as(y).ravel() if y is not None else np.zeros(1)
    lx, ly = x.size, yy.size
    o = [np.zeros(lx * ly) for _ in range(4)]
    fn(_ptr(rfr), _ptr(rfi), _ptr(gr), _ptr(gi), rf.size, _ptr(x), lx, _ptr(yy), ly, *[_ptr(v) for v in o])
    sh = (lx, ly)
    return (o[0] + 1j * o[1]).reshape(sh, order="F"), (o[2] + 1j * o[3]).reshape(sh, order="F")


# ----------------------------------------------------------------------------
# workload pulses (reference recipe: dzrf.m:40,62,81 + msinc.m:12-15 + rfscaleg.m:10-12)
# ----------------------------------------------------------------------------
def msinc(n, m):
    """msinc.m:12-15 — Hamming-windowed sinc with m cycles."""
    x = np.arange(-n / 2, (n - 1) / 2 + 1e-9, 1.0) / (n / 2)
    snc = np.sin(m * 2 * np.pi * x + 0.00001) / (m * 2 * np.pi * x + 0.00001)
    return snc * (0.54 + 0.46 * np.cos(np.pi * x)) * 4 * m / n


def rfscaleg(rf, t_ms, gamma_khz_per_g):
    """rfscaleg.m:10-12 — radians -> Gauss."""
    return np.asarray(rf) / (2 * np.pi * gamma_khz_per_g * (t_ms / len(rf)))


def mag2mp_m(x):
    """mag2mp.m:25-33 — magnitude spectrum -> spectrum of the minimum-phase (analytic) signal."""
    n = len(x)
    xlf = np.fft.fft(np.log(x))
    xlfp = np.zeros(n, dtype=complex)
    xlfp[0] = xlf[0]
    xlfp[1:n // 2] = 2 * xlf[1:n // 2]
    xlfp[n // 2] = xlf[n // 2]
    return np.exp(np.fft.ifft(xlfp))


def b2a_m(bc):
    """b2a.m:16-28 — minimum-phase alpha polynomial for a beta polynomial (pad x8)."""
    bc = np.asarray(bc, dtype=complex)
    n = len(bc)
    blp = n * 8
    bf = np.fft.fft(np.concatenate([bc, np.zeros(blp - n)]))
    bfmax = np.abs(bf).max()
    if bfmax >= 1.0:
        bf = bf / (1e-8 + bfmax)
    afa = mag2mp_m(np.sqrt(1 - (bf * np.conj(bf)).real))
    aca = np.fft.fft(afa) / blp
    return aca[:n][::-1]


def ab2rf_m(ac, bc):
    """ab2rf.m:12-26 — inverse SLR peel-off recursion, complex rf."""
    ac = np.asarray(ac, dtype=complex).copy()
    bc = np.asarray(bc, dtype=complex).copy()
    n = len(ac)
    rf = np.zeros(n, dtype=complex)
    for i in range(n, 0, -1):
        c = np.sqrt(1 / (1 + abs(bc[i - 1] / ac[i - 1]) ** 2))
        s = np.conj(c * bc[i - 1] / ac[i - 1])
        theta = np.arctan2(abs(s), c)
        psi = np.angle(s)
        rf[i - 1] = 2 * (theta * np.cos(psi) + 1j * theta * np.sin(psi))
        acn = c * ac + s * bc
        bcn = -np.conj(s) * ac + c * bc
        ac = acn[1:i]
        bc = bcn[0:i - 1]
    return rf


def dzrf_ms_ex(n, tb):
    """dzrf(n, tb, 'ex', 'ms'): b = sqrt(1/2)*msinc(n, tb/4); rf = b2rf(b)  (dzrf.m:40,62,81).

    The reference's compiled b2rf overflows its static work arrays for n >= 512
    (b2a.code.c:16-17,33-38: nnc = 2*nextpow2(n)*16 > MAXN = 16384), so it is used for
    n < 512 and the toolbox's own .m pair b2a.m + ab2rf.m (what dzrf_mb.m:239-240 calls),
    restated above, for longer pulses.  Workload generation only.
    """
    b = np.sqrt(0.5) * msinc(n, tb / 4.0)
    if n < 512:
        return b2rf_ref(b)
    return ab2rf_m(b2a_m(b), b)

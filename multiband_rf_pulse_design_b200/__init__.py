"""mbrf-b200: B200 (sm_100a) engine behind the MATLAB/MEX signatures of the
multiband RF pulse design toolbox's data-parallel paths.

Host-side mirror of the reference interface (same names, argument meaning and error
behaviour) over the C ABI in ``include/mbrf.h`` / ``libmbrf.so``:

    blochC, blochH, bloch            <- bloch_simulation/blochC.c, blochH.c (mexFunction)
    abrx, abrm, abr                  <- rf_tools/mex5/abrx.c, rf_tools/abrm.m, rf_tools/abr.m
    b2a, ab2rf, b2rf                 <- rf_tools/b2a.m, rf_tools/ab2rf.m (batched inverse SLR, dzrf_mb.m:239-240)
    fmp2                             <- fir_ap_cvx.m:253-304 (batched spectral factorisation)
    fir_ap_cvx, fir_ap               <- fir_ap_cvx.m, fir_ap.m (solve = batched interior point / restarted PDHG on the GPU)
    fir_flip_zero                    <- fir_flip_zero.m (all flip patterns expanded in one launch)
    fir_qprog_phs, fir_min_order_qprog_phs <- ss/fir_qprog_phs.m, ss/fir_min_order_qprog_phs.m
    sim_rf_spectral                  <- sim_rf_spectral.m (the spectral profile of a designed pulse, without the figures)
    sim_rf_scale, bloch_scale_sweep  <- sim_rf_scale.m (all B1 scalings x off-resonances in ONE launch of the sweep kernel)
    dzrf_mb, rfscaleg                <- dzrf_mb.m (the design driver: orchestration of the stages above), rf_tools/rfscaleg.m
    rf_ripple_GFA, rf_Mrange_desired, rf_bandedge, dinf, spectrum_C13, multiband_spec <- the specification builders
                                        (rf_ripple_GFA.m, rf_Mrange_desired.m, rf_bandedge.m, dinf.m, spectrum_C13.m, dzrf_mb.m:92-157)

There is no CPU fallback: importing works anywhere, computing needs the built library
and a CUDA device.
"""
from ._lib import lib, MbrfError, library_path  # noqa: F401
from .bloch import bloch, blochC, blochH, blochsimfz, GAMMA_C13, GAMMA_H1  # noqa: F401
from .slr import ab2rf, abr, abrm, abrx, b2a, b2rf  # noqa: F401
from .fir import (fir_ap, fir_ap_cvx, fir_ap_cvx_batch, fir_linprog, fir_min_order, fir_qp,  # noqa: F401
                  fir_min_order_linprog, fir_qp_cvx, fmp2)
from .fir_post import fir_flip_zero, fir_min_order_qprog_phs, fir_qprog_phs  # noqa: F401
from .design import dzrf_mb, rfscaleg  # noqa: F401
from .sim import bloch_scale_sweep, sim_rf_scale, sim_rf_spectral  # noqa: F401
from .spec import dinf, multiband_spec, rf_bandedge, rf_Mrange_desired, rf_ripple_GFA, spectrum_C13  # noqa: F401

__all__ = ["bloch", "blochC", "blochH", "blochsimfz", "abr", "abrm", "abrx", "b2a", "ab2rf", "b2rf", "fir_ap", "fir_ap_cvx",
           "fir_ap_cvx_batch", "fir_qp", "fir_flip_zero", "fir_qprog_phs", "fir_min_order_qprog_phs", "fir_linprog", "fir_min_order", "fir_min_order_linprog", "fir_qp_cvx", "fmp2", "lib", "MbrfError",
           "library_path", "GAMMA_C13", "GAMMA_H1"]

function [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
%FIR_AP_CVX  Drop-in for the toolbox's fir_ap_cvx.m: same signature, defaults and status strings, with the CVX solve
%  (fir_ap_cvx.m:160-169) replaced by libmbrf's GPU solvers (MEX gateway fir_solve_mex: interior point by default, the
%  first-order solver with setenv('MBRF_FIR_METHOD','pdhg')) and the spectral factorisation (:185-202, :253-304) by the
%  fmp2 MEX gateway (matlab/islr_mex.c, -DMBRF_STAGE=3).  fir_ap_cvx_batch.m solves many designs in one call.
%
%     [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
%
%  The problem data (grid, per-band bounds, transition rows, stop rows) follows fir_ap_cvx.m:44-142 step by step.
narginchk(4, 7);
if nargin < 5 || isempty(obj),  obj = 0;  end
if nargin < 6 || isempty(Peak), Peak = 1e-3; end
if obj < 0, error('invalid input of obj'); end          % fir_ap_cvx.m:172-174

[hs, sts] = fir_ap_cvx_batch(n, {f}, a, d, obj, Peak);
h = hs{1};   status = sts{1};
end

function [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
%FIR_AP_CVX  Drop-in for the toolbox's fir_ap_cvx.m: same signature, defaults and status strings, with the CVX solve
%  (fir_ap_cvx.m:160-169) replaced by libmbrf's batched restarted PDHG on the GPU (MEX gateway fir_pdhg_mex) and the
%  spectral factorisation (:185-202, :253-304) by the fmp2 MEX gateway (matlab/islr_mex.c, -DMBRF_STAGE=3).
%
%     [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
%
%  The problem data (grid, per-band bounds, transition rows, stop rows) follows fir_ap_cvx.m:44-142 step by step.
narginchk(4, 7);
if nargin < 5 || isempty(obj),  obj = 0;  end
if nargin < 6 || isempty(Peak), Peak = 1e-3; end
if obj < 0, error('invalid input of obj'); end          % fir_ap_cvx.m:172-174

edges = reshape(f, 1, []) * pi;   amps = reshape(a, 1, []);   ripple = reshape(d, 1, []);
nbands = numel(edges) / 2;
w = sort([linspace(-pi, pi, 2 * n * 15), edges]);                         % :44-48, oversampling 15, band edges added
inband = false(size(w));   upper = [];   lower = [];   bandidx = [];
for k = 1:nbands
    e0 = edges(2*k-1);   e1 = edges(2*k);
    sel = find(w >= e0 & w <= e1);                                         % :54
    if e0 == e1
        target = repmat(amps(2*k-1), size(sel));                           % :57-58
    else
        target = amps(2*k-1) + (amps(2*k) - amps(2*k-1)) * (w(sel) - e0) / (e1 - e0);   % :60
    end
    bandidx = [bandidx, sel];              %#ok<AGROW>
    upper = [upper, target + ripple(k)];   %#ok<AGROW>
    lower = [lower, target - ripple(k)];   %#ok<AGROW>
    inband(sel) = true;
end
tranidx = find(~inband);                                                   % :67-82
w = [w(bandidx), w(tranidx)];                                              % :86-91, band rows first
m = numel(w);
U_b = [upper, repmat(max(upper), 1, numel(tranidx))] .^ 2;                 % :103-106
L_b = max([lower, repmat(min(0, min(lower)), 1, numel(tranidx))], 0) .^ 2; % :110-112
L_b = max(L_b, 1e-20);                                                     % :115-116
stoprows = find(sqrt(U_b) < min(sqrt(U_b)) + 1e-2);                        % :125
ns = numel(stoprows);

% canonical form of the solver: z = x (2n-1 autocorrelation coefficients); the stop rows are appended once more as the
% "simplex block", because obj*ripple_stop with A(idx_stop,:)*x <= ripple_stop is obj * max_i (A x)_i over those rows
nx = 2*n - 1;
w_row = [w, w(stoprows)];
col_type  = [0, ones(1, n-1), 2*ones(1, n-1)];        % 1, cos, sin columns of A = [1, 2cos(w k), 2sin(w k)], :100
col_kappa = [0, 1:n-1, 1:n-1];
col_amp   = [1, 2*ones(1, 2*n-2)];
c  = [1; zeros(nx-1, 1)];                             % minimise x(1) + ..., :163
lo = [L_b, -inf(1, ns)].';
hi = [U_b, zeros(1, ns)].';
bl = -inf(nx, 1);   bu = inf(nx, 1);
bl(1) = -n*Peak;    bu(1) = n*Peak;                   % |x(1)| <= n Peak, :166-168 (i = 1)
rho = ((n - (2:n) + 1) * Peak).';                     % ||(x_i, x_{n+i-1})|| <= (n-i+1) Peak
obj_upper = n*Peak + obj * max(U_b(stoprows));        % no feasible point has a larger objective
opts = [200000, 64, 8e-7, 1e-4, 5e-5];                % max iterations, check period, eps_pr, eps_dr, eps_gap
[z, info] = fir_pdhg_mex(w_row, [], col_type, col_kappa, col_amp, 0, 2:n, n+1:2*n-1, c, lo, hi, bl, bu, rho, ...
                         obj_upper, opts, [m+1, ns, obj]);
if info(1) ~= 1                                       % 2: infeasible, 3: iteration limit -> 'Failed', :176-182
    status = 'Failed';   h = [];
    return
end
status = 'Solved';
x = z(1:nx);
r = [x(1); x(2:n) + 1i * x(n+1:nx)];                  % :185
h = fmp2([conj(flipud(r(2:end))); r]);                % :186 and the minimum-phase factor, on the GPU
end

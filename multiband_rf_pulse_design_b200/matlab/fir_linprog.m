function [h, status, info] = fir_linprog(n, f, a, d, h0, dbg, a_min) %#ok<INUSL>
%FIR_LINPROG  Drop-in for the toolbox's ss/fir_linprog.m (same signature, status strings and tap layout): the LP
%     min fmin*x  s.t.  [A; -A] x <= [U, -L]                                   (ss/fir_linprog.m:221-252)
%  goes to libmbrf's interior-point solver on the GPU (fir_solve_mex) instead of MATLAB's linprog.  h0 is the reference's
%  starting point for linprog's medium-scale algorithm (:157); an interior-point method starts from its own centred point,
%  so h0 does not change the result and is not used.  info (third output) is the solver's 8-vector (status code first).
%  a_min (optional 7th argument, used by fir_min_order.m): lower bound of the response in the transition regions
%  (ss/fir_pm.m:42-43,102); default min(0, min(L)) as in ss/fir_linprog.m:165-170.
f = reshape(f, 1, []) * pi;   a = reshape(a, 1, []);   d = reshape(d, 1, []);      % :46
real_filter = ~(min(f) < 0);                                                       % :48-52
odd = mod(n, 2) == 1;                                                              % :56-60
h = [];   status = 'Failed';   info = [];
if ~odd && any(a(abs(f) == pi) == 1), return; end                                  % :63-75
nhalf = ceil(n / 2);                                                               % :79
if real_filter, w = linspace(0, pi, 15*n); else, w = linspace(-pi, pi, 30*n); end  % :92-101
w = sort([w, f]);                                                                  % :107
nbands = numel(f) / 2;
inband = false(size(w));   U = [];   L = [];   bandidx = [];
for k = 1:nbands
    e0 = f(2*k-1);   e1 = f(2*k);
    sel = find(w >= e0 & w <= e1);
    if e0 == e1, target = repmat(a(2*k-1), size(sel));
    else,        target = a(2*k-1) + (a(2*k) - a(2*k-1)) * (w(sel) - e0) / (e1 - e0); end
    bandidx = [bandidx, sel];  U = [U, target + d(k)];  L = [L, target - d(k)];  %#ok<AGROW>
    inband(sel) = true;
end
tranidx = find(~inband);                                                           % :163-171
w = [w(bandidx), w(tranidx)];                                                      % :175-180
hi = [U, repmat(max(U), 1, numel(tranidx))].';                                     % :221-226 (amplitude, not power)
if nargin < 7 || isempty(a_min), a_min = min(0, min(L)); end
lo = [L, repmat(a_min, 1, numel(tranidx))].';
if odd                                                                             % :195-217
    kc = 1:nhalf-1;
    ctype = [0, ones(1, nhalf-1)];   kappa = [0, kc];   amp = [1, 2*ones(1, nhalf-1)];
    if ~real_filter, ctype = [ctype, 2*ones(1, nhalf-1)];  kappa = [kappa, kc];  amp = [amp, 2*ones(1, nhalf-1)]; end
else
    kc = (0:nhalf-1) + 0.5;
    ctype = ones(1, nhalf);   kappa = kc;   amp = 2*ones(1, nhalf);
    if ~real_filter, ctype = [ctype, 2*ones(1, nhalf)];  kappa = [kappa, kc];  amp = [amp, 2*ones(1, nhalf)]; end
end
wt = w(numel(bandidx)+1:end);                                                      % fmin = sum(A(idx_tran,:), 1), :231
c = zeros(numel(ctype), 1);
for j = 1:numel(ctype)
    switch ctype(j)
        case 0, c(j) = amp(j) * numel(wt);
        case 1, c(j) = amp(j) * sum(cos(wt * kappa(j)));
        case 2, c(j) = amp(j) * sum(sin(wt * kappa(j)));
    end
end
[z, info] = fir_solve_mex(1, w, [], [], ctype, kappa, amp, [], [], [], c, lo, hi, [], [], [], [], [100, 1e-7, 2e-6, 1e-12], [], []);
if info(1) ~= 1, return; end                                                       % exitflag ~= 1 -> 'Failed', :265-271
status = 'Solved';
x = z(:).';
if real_filter                                                                     % fill_h, :274-296
    if odd, h = [x(end:-1:2), x]; else, h = [x(end:-1:1), x]; end
elseif odd
    hh = x(1:nhalf) + 1i * [0, x(nhalf+1:end)];
    h = [conj(hh(end:-1:2)), hh];
else
    hh = x(1:nhalf) + 1i * x(nhalf+1:end);
    h = [conj(hh(end:-1:1)), hh];
end
end

function [hs, sts, info, xs] = fir_ap_cvx_batch(n, fcell, a, d, objs, Peaks)
%FIR_AP_CVX_BATCH  B fir_ap_cvx designs of one order n in ONE solver call (the reference solves them one CVX call at a
%  time: the bisections of fir_ap.m:78-105, trade-off sweeps).  fcell{b} are the band edges of design b, objs(b) / Peaks(b)
%  its stop-band weight and peak bound (scalars are expanded); a, d are shared.  Returns cell arrays hs (taps, [] where
%  'Failed') and sts ('Solved' / 'Failed'), info (8-by-B: status code 1 solved / 2 infeasible / 3 iteration limit, iterations,
%  objective, dual objective, max violation, residual, bound, ripple_stop) and xs (the autocorrelation variables, (2n-1)-by-B).
%
%  The problem data of every design (grid, per-band bounds, transition rows, stop rows) follows fir_ap_cvx.m:44-142 step by
%  step; designs with different band edges share one matrix through the union of their grid rows (a row a design does not
%  own gets the bounds (-inf, +inf)), the stop rows are appended once more as the block of `obj * max_i (A x)_i`.
B = numel(fcell);
if isscalar(objs),  objs = repmat(objs, 1, B);   end
if isscalar(Peaks), Peaks = repmat(Peaks, 1, B); end
if ~strcmpi(getenv('MBRF_FIR_METHOD'), 'pdhg')
    % interior point (default): ONE call, specification in -> taps out.  Grid, band masks, bounds, stop rows and radii
    % (fir_ap_cvx.m:44-142) are assembled on the GPU, the batch is solved there and h = fmp2(r) (:185-202) is taken there too.
    F = zeros(numel(fcell{1}), B);
    for b = 1:B, F(:, b) = reshape(fcell{b}, [], 1); end
    [H, info, xs] = fir_ap_mex(n, F, reshape(a, [], 1), reshape(d, [], 1), objs, Peaks);
    hs = cell(1, B);   sts = cell(1, B);
    for b = 1:B
        if info(1, b) == 1, sts{b} = 'Solved';  hs{b} = H(:, b);          % 2: infeasible, 3: iteration limit -> 'Failed', :176-182
        else,               sts{b} = 'Failed';  hs{b} = [];  end
    end
    return
end
% first-order solver: the problem is assembled here and handed over as arrays
amps = reshape(a, 1, []);   ripple = reshape(d, 1, []);
W = cell(1, B);  LO = cell(1, B);  HI = cell(1, B);  STOP = cell(1, B);
for b = 1:B
    edges = reshape(fcell{b}, 1, []) * pi;                                    % fir_ap_cvx.m:44
    nbands = numel(edges) / 2;
    w = sort([linspace(-pi, pi, 2 * n * 15), edges]);                         % :45-48
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
    W{b} = [w(bandidx), w(tranidx)];                                           % :86-91
    U_b = [upper, repmat(max(upper), 1, numel(tranidx))] .^ 2;                 % :103-106
    L_b = max([lower, repmat(min(0, min(lower)), 1, numel(tranidx))], 0) .^ 2; % :110-112
    LO{b} = max(L_b, 1e-20);   HI{b} = U_b;                                    % :115-116
    STOP{b} = sqrt(U_b) < min(sqrt(U_b)) + 1e-2;                               % :125
end
allw = unique([W{:}]);                         % union of the designs' grid points
M1 = numel(allw);
stop_any = false(1, M1);
pos = cell(1, B);
for b = 1:B
    [~, pos{b}] = ismember(W{b}, allw);
    stop_any(pos{b}(STOP{b})) = true;
end
srows = find(stop_any);   ns = numel(srows);
srank = zeros(1, M1);   srank(srows) = 1:ns;
M = M1 + ns;   nx = 2*n - 1;
lo = -inf(M, B);   hi = inf(M, B);
for b = 1:B
    for k = 1:numel(pos{b})                    % a grid point may occur twice (band edge on a base sample): keep the tighter
        r = pos{b}(k);
        lo(r, b) = max(lo(r, b), LO{b}(k));
        hi(r, b) = min(hi(r, b), HI{b}(k));
    end
    hi(M1 + srank(pos{b}(STOP{b})), b) = 0;    % membership flag of the stop block (fir_ap_cvx.m:165)
end
w_row = [allw, allw(srows)];
col_type  = [0, ones(1, n-1), 2*ones(1, n-1)];        % A = [1, 2cos(w k), 2sin(w k)], :100
col_kappa = [0, 1:n-1, 1:n-1];
col_amp   = [1, 2*ones(1, 2*n-2)];
c  = [ones(1, B); zeros(nx-1, B)];                    % minimise x(1) + obj*ripple_stop, :163
bl = -inf(nx, B);   bu = inf(nx, B);
bl(1, :) = -n*Peaks;    bu(1, :) = n*Peaks;           % |x(1)| <= n Peak, :166-168 (i = 1)
rho = (n - (2:n) + 1).' * reshape(Peaks, 1, []);      % ||(x_i, x_{n+i-1})|| <= (n-i+1) Peak
upper_obj = zeros(1, B);
for b = 1:B, upper_obj(b) = n*Peaks(b) + objs(b) * max(HI{b}(STOP{b})); end
blocks = [M1+1, ns, 0, 0, 0, 0, 0, 0, 0];
block_w = [reshape(objs, 1, []); zeros(3, B)];
[z, info] = fir_solve_mex(0, w_row, [], [], col_type, col_kappa, col_amp, [], 2:n, n+1:2*n-1, c, lo, hi, bl, bu, rho, ...
                          upper_obj, [200000, 64, 8e-7, 1e-4, 5e-5], blocks, block_w);
xs = z;
hs = cell(1, B);   sts = cell(1, B);
ok = find(info(1, :) == 1);                           % 2: infeasible, 3: iteration limit -> 'Failed', :176-182
for b = 1:B, sts{b} = 'Failed';  hs{b} = [];  end
if ~isempty(ok)
    R = zeros(2*nx - 1 - (nx - 1), numel(ok));        % two-sided sequences of length 2n-1
    for k = 1:numel(ok)
        x = z(:, ok(k));
        r = [x(1); x(2:n) + 1i * x(n+1:nx)];          % :185
        R(:, k) = [conj(flipud(r(2:end))); r];        % :186
    end
    H = fmp2(R);                                      % minimum-phase factors of all solved designs in one GPU call
    for k = 1:numel(ok)
        sts{ok(k)} = 'Solved';   hs{ok(k)} = H(:, k);
    end
end
end

function [h, status, info] = fir_qprog_phs(n, f, ac, dc, x0, dbg) %#ok<INUSL>
%FIR_QPROG_PHS  Drop-in for ss/fir_qprog_phs.m (same signature and status strings): minimum-energy FIR filter whose response
%  stays inside a magnitude AND phase window per band.  The reference builds  min 1/2 x'x  s.t.  A x <= B  and calls quadprog
%  (ss/fir_qprog_phs.m:326-345; x0 is replaced by [] at :337 and is unused here as well).  Minimising ||x|| has the same
%  minimiser, and ||x|| is the norm term of libmbrf's first-order solver (fir_solve_mex, method 'pdhg'): every row of A is
%  sign * Re(exp(-1i*(w*q + phase)) * h), i.e. one Fourier matrix with a phase per row, K = [cos(w q + phase), sin(w q + phase)].
if nargin < 6, dbg = 0; end %#ok<NASGU>
f = reshape(f, 1, []);   ac = reshape(ac, 1, []);   dc = reshape(dc, 1, []);
nband = numel(f) / 2;                                                              % :51
if any(ac(1:2:end) ~= ac(2:2:end)), error('Does not support sloped bands'); end    % :55-59
a = abs(ac(1:2:end));   aphs = angle(a);                                           % :63-67 (angle AFTER abs, as the reference)
d = abs(dc);   dphs = angle(dc);                                                   % :68-69
straddle = (a + d) .* (a - d) < 0;
if any(straddle & (a ~= 0 | dphs ~= 0)), error('Bands straddling 0 must have a = 0, angle(d) = 0'); end   % :76-82
err_tol = 0.05;                                                                    % :87
for b = find(a ~= 0)                                                               % :88-100
    if (a(b) - d(b)) * (sec(dphs(b)) - 1) >= 2 * d(b)
        warning('Reducing phase ripple to that feasible');
        dphs(b) = 0.99 * acos((a(b) - d(b)) / (a(b) + d(b)));
    end
end
nseg = ceil(2 * pi / acos(1 - err_tol));                                           % :105
amax = max(a + d);
circle = (0:nseg) / nseg * 2 * pi;
tran_pts = circle;   band_pts = cell(1, nband);
for b = 1:nband                                                                    % :108-124
    if a(b) == 0
        band_pts{b} = circle;
    else
        np_ = ceil(2 * dphs(b) / acos(1 - err_tol * 2 * d(b)));
        band_pts{b} = ((0:np_) / np_ * 2 - 1) * dphs(b) + aphs(b);
        if a(b) + d(b) >= amax * (1 - err_tol), tran_pts = [tran_pts, aphs(b) - dphs(b), aphs(b) + dphs(b)]; end %#ok<AGROW>
    end
end
tran_pts = unique([mod(tran_pts, 2 * pi), 0, 2 * pi]);                             % :128-129
fw = f * pi;                                                                       % :178
odd = mod(n, 2) == 1;
h = [];   status = 'Failed';   info = [];
if ~odd && any(abs(ac(abs(fw) == pi)) ~= 0)                                        % :190-201
    warning('n odd and frequency spec non-zero at fs/2');
    return
end
nhalf = ceil(n / 2);                                                               % :205
w = sort([linspace(-pi, pi, 2 * 15 * n), fw]);                                     % :212-222
if odd, q = -(nhalf - 1):(nhalf - 1); else, q = (-nhalf:nhalf - 1) + 0.5; end      % :226-230
inband = false(size(w));
uw = [];  up = [];  ub = [];   lw = [];  lp = [];  lb = [];
for b = 1:nband                                                                    % :237-272
    sel = w(w >= fw(2*b-1) & w <= fw(2*b));
    inband(w >= fw(2*b-1) & w <= fw(2*b)) = true;
    pts = band_pts{b};
    step = angle(exp(1i * pts(2)) * exp(-1i * pts(1)));                            % :244-245
    for k = 1:numel(pts) - 1                                                       % polygon around the magnitude bound, :247-253
        uw = [uw, sel];  up = [up, repmat(pts(k) + step / 2, size(sel))];  ub = [ub, repmat((a(b) + d(b)) * cos(step / 2), size(sel))]; %#ok<AGROW>
    end
    if a(b) ~= 0
        lw = [lw, sel];  lp = [lp, repmat(aphs(b), size(sel))];          lb = [lb, repmat(a(b) - d(b), size(sel))]; %#ok<AGROW>  :257-260
        uw = [uw, sel];  up = [up, repmat(pts(end) + pi / 2, size(sel))];  ub = [ub, zeros(size(sel))];             %#ok<AGROW>  :264-265
        lw = [lw, sel];  lp = [lp, repmat(pts(1) + pi / 2, size(sel))];    lb = [lb, zeros(size(sel))];             %#ok<AGROW>  :269-271
    end
end
wt = w(~inband);                                                                   % :276-282
for k = 1:numel(tran_pts) - 1                                                      % :306-317
    step = tran_pts(k + 1) - tran_pts(k);
    uw = [uw, wt];  up = [up, repmat(tran_pts(k) + step / 2, size(wt))];  ub = [ub, repmat(amax * cos(step / 2), size(wt))]; %#ok<AGROW>
end
w_row = [uw, lw];   row_phase = [up, lp];   M = numel(w_row);   N = 2 * n;
if any(isnan(row_phase)), return; end                                              % zero phase ripple on a pass band: 0/0 at :113
hi = [ub, inf(size(lb))].';   lo = [-inf(size(ub)), lb].';
fin = [ub, lb];
big = 2 * sqrt(n) * max(1, max(abs(fin)));
blocks = zeros(1, 9);   blocks(7) = N;                                             % norm term over all of x
% every grid point lies in a polygon inscribed in |H| = amax and the base grid is a DFT grid of 30n - 1 >= n points, so
% ||x|| <= amax for every feasible point: a dual bound above it certifies infeasibility
[z, info] = fir_solve_mex(0, w_row, row_phase, [], [ones(1, n), 2 * ones(1, n)], [q, q], ones(1, N), [], [], [], zeros(N, 1), ...
                          lo, hi, -big * ones(N, 1), big * ones(N, 1), [], amax * (1 + 1e-9), [400000, 64, 8e-7, 1e-5, 2e-5], ...
                          blocks, [0; 0; 1; 0]);
if info(1) ~= 1, return; end                                                       % exitflag == 1, :388-394
status = 'Solved';
h = z(1:n) + 1i * z(n+1:2*n);                                                      % a column, like x(1:n) + i*x(n+1:end)
end

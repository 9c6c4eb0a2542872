function [h, status, info] = fir_qp_cvx(n, f, a, d, k, obj, dbg) %#ok<INUSD>
%FIR_QP_CVX  Drop-in for the toolbox's fir_qp_cvx.m (same signature, defaults, status strings).  The two CVX programs
%     scalar obj (fir_qp_cvx.m:145-166):   min E_total + obj*Peak   s.t. ||A_i x - Hd_i|| <= D_i (bands), ||A_i x|| <= 1 + 5 max(d)
%                                          (transitions), ||(x_i, x_{n+i})|| <= Peak, ||x|| <= E_total
%     obj = [o1 o2] (:170-191):            min delta + o1*E_total + o2*Peak   s.t. ||A_i x - Hd_i|| <= D_i*delta, ||A_i x|| <= 1.1, ...
%  go to libmbrf's first-order solver on the GPU (fir_solve_mex, method 'pdhg') with E_total, Peak and delta eliminated:
%  E_total = ||x|| is a norm term, obj*Peak = obj*max_i ||(x_i, x_{n+i})|| a group block over identity rows, the response
%  constraints are disk rows, delta a centred group block over the band rows scaled by 1/D_i (include/mbrf.h, mbrf_pdhg_blocks).
if nargin < 5 || isempty(k),   k = 100; end                                        % :28-31
if nargin < 6 || isempty(obj), obj = 0; end
if ~any(numel(obj) == [1 2]), error('invalid input of obj'); end                   % :193-195
minimax = numel(obj) == 2;
f = reshape(f, 1, []) * pi;   a = reshape(a, 1, []);   d = reshape(d, 1, []);      % :34
w = sort([linspace(-pi, pi, n * 10), f]);                                          % :35-38
nbands = numel(f) / 2;
inband = false(size(w));   Mb = [];   Db = [];   bandidx = [];
for b = 1:nbands                                                                   % :46-61
    e0 = f(2*b-1);   e1 = f(2*b);
    sel = find(w >= e0 & w <= e1);
    if e0 == e1, amp = repmat(a(2*b-1), size(sel));
    else,        amp = a(2*b-1) + (a(2*b) - a(2*b-1)) * (w(sel) - e0) / (e1 - e0); end
    bandidx = [bandidx, sel];  Mb = [Mb, amp];  Db = [Db, repmat(d(b), size(sel))];  %#ok<AGROW>
    inband(sel) = true;
end
wband = w(bandidx);   wtran = w(~inband);                                          % :75-76
Hd = Mb .* exp(1i * (k * wband.^2 - wband * (n - 1) / 2));                         % :113-121
wall = [wband, wtran];   m = numel(wall);   nb = numel(wband);
centre = [Hd, zeros(1, numel(wtran))];
radius = [Db, repmat(1 + 5 * max(d), 1, numel(wtran))];                            % :151,156
N = 2*n;   M = 2*m + 2*n;
w_row = [reshape([wall; wall], 1, []), zeros(1, 2*n)];
row_phase = [repmat([0, pi/2], 1, m), zeros(1, 2*n)];                              % [cos sin; -sin cos], :96-109
row_scale = [ones(1, 2*m), zeros(1, 2*n)];
col_type = [ones(1, n), 2*ones(1, n)];   col_kappa = [0:n-1, 0:n-1];   col_amp = ones(1, N);
entries = [(2*m + (1:2*n)).', reshape([1:n; n + (1:n)], [], 1), ones(2*n, 1)];     % identity rows, pair i = (x_i, x_{n+i}), :126-139
lo = -inf(M, 1);   hi = inf(M, 1);
blocks = zeros(1, 9);
if minimax
    row_scale(1:2:2*nb) = 1 ./ Db;   row_scale(2:2:2*nb) = 1 ./ Db;               % band rows / D_i: ||.|| <= delta, :176
    lo(1:2:2*nb) = real(Hd) ./ Db;   lo(2:2:2*nb) = imag(Hd) ./ Db;
    lo(2*nb+1:2*m) = 0;   hi(2*nb+1:2:2*m) = 1.1;                                  % transition disks, :180
    blocks([8 9]) = [1, nb];   blocks([3 4]) = [2*nb + 1, m - nb];
    gw = obj(2);   lam = obj(1);   g2 = 1;
else
    lo(1:2:2*m) = real(centre);   lo(2:2:2*m) = imag(centre);   hi(1:2:2*m) = radius;
    blocks([3 4]) = [1, m];
    gw = obj(1);   lam = 1;   g2 = 0;
end
big = 2 * max(radius) + 2 * max(abs(centre)) + 2 * minimax;
blocks([5 6]) = [2*m + 1, n];   blocks(7) = N;
% the solver sees the objective divided by its largest weight (obj = 1e6 in dzrf_mb.m:211-213: "minimise Peak" with the energy
% as a 1e-6 tie-break; a first-order method needs O(1) weights); the reported objective is scaled back
oscale = max([1, gw, lam]);
block_w = [0; gw; lam; g2] / oscale;
[z, info] = fir_solve_mex(0, w_row, row_phase, row_scale, col_type, col_kappa, col_amp, entries, [], [], zeros(N, 1), lo, hi, ...
                          -big * ones(N, 1), big * ones(N, 1), [], [], [400000, 64, 8e-7, 1e-4, 5e-5], blocks, block_w);
info([3 4 7]) = info([3 4 7]) * oscale;
if info(1) ~= 1                                                                    % :200-206
    status = 'Failed';   h = [];
    return
end
status = 'Solved';
h = z(1:n) + 1i * z(n+1:2*n);                                                      % :209
h = h(:).';
end

function [h, status] = fir_qp(n, f, a, d, min_order, min_tran, min_peak, dbg)
%FIR_QP  Drop-in for the toolbox's fir_qp.m (same signature and defaults): low-pass design with a symmetric pass band
%  f = [-fp, fp, fs, f4], every probe fir_ap_cvx(n, f, a, d, 1e5) (fir_qp.m:47,68,92,107,130), optional bisection on the
%  transition width (:57-96, threshold 1e-3) and on the order (:103-134), optional zero flipping (:136-146).
%  The transition search solves the 7 probes the next three bisection steps can ask for as ONE batch on the GPU
%  (fir_ap_cvx_batch) and then walks the tree with the reference's decisions.
if nargin < 4, error('not enough input'); end
if nargin < 5, min_order = 0; end
if nargin < 6, min_tran = 0; end
if nargin < 7, min_peak = 0; end
if nargin < 8, dbg = 0; end
lambda = 1e5;   df_thre = 0.001;   Peak = 1e-3;                                    % :36-37, fir_ap_cvx.m:33
[h, status] = fir_ap_cvx(n, f, a, d, lambda);                                       % :47
if strcmp(status, 'Failed'), error('original parameters are too tight'); end       % :48-50
if min_tran > 0
    centre = (f(3) + f(2)) / 2;   top = (f(3) - f(2)) / 2;   bot = 0;               % :58-60
    edges = @(dfv) [-(centre - dfv), centre - dfv, centre + dfv, f(4)];            % :64-67
    done = false;
    while ~done
        span = top - bot;
        fc = cell(1, 7);
        for j = 1:7, fc{j} = edges(bot + span * j / 8); end
        [hs, sts] = fir_ap_cvx_batch(n, fc, a, d, lambda, Peak);
        lo = 0;   hi = 8;
        for level = 1:3
            mid = (lo + hi) / 2;                                                   % df_mid = (df_top + df_bot)/2, :63
            if strcmp(sts{mid}, 'Failed'), lo = mid;                               % :69-71
            else, h = hs{mid};  status = sts{mid};  hi = mid; end                  % :72-76
            if span * (hi - lo) / 8 < df_thre, done = true; break; end             % :78-80
        end
        top = bot + span * hi / 8;   bot = bot + span * lo / 8;
    end
    if ~(min_tran > 0 && min_tran <= 1), error('invalid input of min_tran'); end   % :97-99
    f = edges(((f(3) - f(2)) / 2) * (1 - min_tran) + top * min_tran);              % :90-96
    [h, status] = fir_ap_cvx(n, f, a, d, lambda);
elseif min_tran ~= 0
    error('invalid input of min_tran');
end
if min_order > 0                                                                    % :103-123
    n_top = n;   n_bot = 2;
    while true
        n_mid = ceil((n_top + n_bot) / 2);
        [h0, st0] = fir_ap_cvx(n_mid, f, a, d, lambda);
        if strcmp(st0, 'Failed'), n_bot = n_mid; else, h = h0;  status = st0;  n_top = n_mid; end
        if n_top - n_bot == 1, break; end
    end
    if min_order > 0 && min_order < 1                                               % :128-131
        [h, status] = fir_ap_cvx(ceil(n * (1 - min_order) + n_top * min_order), f, a, d, lambda);
    elseif min_order ~= 1
        error('invalid input of min_order');                                        % :132-134
    end
elseif min_order ~= 0
    error('invalid input of min_order');
end
if min_peak, h = fir_flip_zero(h, dbg); end                                         % :136-146
end

function [h, status, n_op, f_op] = fir_ap(n, f, a, d, Peak, min_order, min_tran, min_peak, dbg) %#ok<INUSD>
%FIR_AP  Drop-in for the toolbox's fir_ap.m (same signature): bisection on the transition width (fir_ap.m:57-134) and / or on
%  the order (:137-176), every probe a fir_ap_cvx design with lambda = 0.1 solved on the GPU.  The transition search probes
%  the next three levels of the reference's bisection tree (7 band-edge expansions, same n) as ONE batch and then walks the
%  tree with the reference's decisions; the order search solves one order per call (different n = different matrices).
%  min_peak (fir_flip_zero) is not part of the accelerated path.
if nargin < 5 || isempty(Peak), Peak = 1e-3; end
if nargin < 6 || isempty(min_order), min_order = 0; end
if nargin < 7 || isempty(min_tran), min_tran = 0; end
if nargin < 8 || isempty(min_peak), min_peak = 0; end
if min_peak, error('fir_flip_zero (min_peak) is outside the accelerated path'); end
f = reshape(f, 1, []);
lambda = 0.1;   df_thre = 0.0005;                                          % fir_ap.m:45-46
n_op = n;   f_op = f;
[h, status] = probe(n, f);                                                 % :51
if strcmp(status, 'Failed'), error('original parameters are too tight'); end   % :52-54
if min_tran > 0
    df_min = min(f(3:2:end-1) - f(2:2:end-2));                             % :60-61
    bot = 0;   top = df_min / 2;                                           % :62-63
    done = false;
    while ~done
        span = top - bot;
        fs = cell(1, 7);
        for j = 1:7, fs{j} = widen(bot + span * j / 8); end
        [hs, sts] = fir_ap_cvx_batch(n, fs, a, d, lambda, Peak);
        lo_j = 0;   hi_j = 8;
        for level = 1:3
            mid = (lo_j + hi_j) / 2;                                       % f_add_mid = (bot+top)/2, :79
            if strcmp(sts{mid}, 'Failed')
                hi_j = mid;                                                % :86-88
            else
                h = hs{mid};   status = sts{mid};   lo_j = mid;            % :89-93
            end
            if span * (hi_j - lo_j) / 8 < df_thre, done = true; break; end % :100-102
        end
        top = bot + span * hi_j / 8;   bot = bot + span * lo_j / 8;
    end
    if ~(min_tran > 0 && min_tran <= 1), error('invalid input of min_tran'); end   % :131-133
    fa = bot * min_tran;                                                   % :110
    [h0, st0] = probe(n, widen(fa));                                       % :116
    if strcmp(st0, 'Failed'), fa = bot; else, h = h0; status = st0; end    % :117-121
    f = widen(fa);   f_op = f;
end
if min_order > 0
    n_top = n;   n_bot = 2;                                                % :140-141
    while n_top - n_bot > 1                                                % :143-162
        n_mid = ceil((n_top + n_bot) / 2);
        [h0, st0] = probe(n_mid, f);
        if strcmp(st0, 'Failed'), n_bot = n_mid; else, n_top = n_mid; h = h0; status = st0; end
    end
    if min_order == 1
        n_op = n_top;                                                      % :164-166
    elseif min_order > 0 && min_order < 1
        n_new = ceil(n * (1 - min_order) + n_top * min_order);             % :169-173
        [h, status] = probe(n_new, f);   n_op = n_new;
    else
        error('invalid input of min_order');                               % :174-176
    end
end

    function fn = widen(fa)
        fn = f;   fn(1:2:end) = fn(1:2:end) - fa;   fn(2:2:end) = fn(2:2:end) + fa;    % :71-73
    end
    function [hh, st] = probe(nn, ff)
        % one fir_ap_cvx design; an iteration-limit ending (info(1) == 3: undecided, not infeasible) is reported
        [hq, sq, inf8] = fir_ap_cvx_batch(nn, {ff}, a, d, lambda, Peak);
        hh = hq{1};   st = sq{1};
        if inf8(1) == 3
            warning('mbrf:UndecidedProbe', 'fir_ap_cvx probe at n = %d ended at the iteration limit: treated as Failed', nn);
        end
    end
end

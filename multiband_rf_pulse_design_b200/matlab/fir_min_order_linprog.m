function [h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg, pick_longer)
%FIR_MIN_ORDER_LINPROG  Drop-in for ss/fir_min_order_linprog.m:54-234: bisection on the half-length, odd lengths 2k-1 first
%  (:98-147), then even lengths 2k capped by the best odd one (:151-211), every probe one fir_linprog solve on the GPU; returns
%  the SHORTER of the two (:220-228).  pick_longer = true (used by fir_min_order.m) keeps ss/fir_min_order.m's selection of
%  the longer one (:222-226).  A probe that ends at the solver's iteration limit (info(1) == 3: undecided, NOT infeasible) is
%  treated as 'Failed' with a warning, as the reference treats every exitflag ~= 1.
if nargin < 5 || isempty(even_odd) || ~any(even_odd == [1 2]), even_odd = 0; end   % :65-69
if nargin < 6, dbg = 0; end
if nargin < 7, pick_longer = false; end
n_odd_max = 2 * floor((n - 1) / 2) + 1;                                            % :78-79
n_even_max = 2 * floor(n / 2);
hbest_odd = [];   hbest_even = [];
if even_odd ~= 2
    hbest_odd = bisect((n_odd_max + 1) / 2, @(k) 2*k - 1);
end
if even_odd ~= 1
    n_top = n_even_max / 2;
    if ~isempty(hbest_odd), n_top = min(n_top, (numel(hbest_odd) + 1) / 2); end
    hbest_even = bisect(n_top, @(k) 2*k);
end
h = [];   status = 'Failed';
if isempty(hbest_odd) && isempty(hbest_even), return; end
status = 'Solved';
if pick_longer
    if numel(hbest_odd) > numel(hbest_even), h = hbest_odd; else, h = hbest_even; end
elseif isempty(hbest_odd)
    h = hbest_even;
elseif isempty(hbest_even) || numel(hbest_odd) < numel(hbest_even)
    h = hbest_odd;
else
    h = hbest_even;
end

    function hbest = bisect(n_top, tap_of)
        % the reference's loop, :91-147 / :152-211: state (n_bot, n_top, n_cur), one probe per step
        hbest = [];   n_bot = 1;   n_cur = n_top;
        while n_top - n_bot > 1
            [hc, st, inf8] = fir_linprog(tap_of(n_cur), f, a, d, hbest, dbg);
            if ~isempty(inf8) && inf8(1) == 3
                warning('mbrf:UndecidedProbe', 'fir_linprog probe at n = %d ended at the iteration limit: treated as Failed', tap_of(n_cur));
            end
            if strcmp(st, 'Solved')
                hbest = hc;   n_top = n_cur;
                if n_top == n_bot + 1, n_cur = n_bot; else, n_cur = ceil((n_top + n_bot) / 2); end   % :131-136
            else
                n_bot = n_cur;   n_cur = ceil((n_bot + n_top) / 2);                                   % :141-142
            end
        end
    end
end

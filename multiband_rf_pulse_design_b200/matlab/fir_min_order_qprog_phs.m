function [h, status] = fir_min_order_qprog_phs(n, f, a, d, even_odd, dbg)
%FIR_MIN_ORDER_QPROG_PHS  Drop-in for ss/fir_min_order_qprog_phs.m: the bisection of fir_min_order_linprog (odd lengths, then
%  even lengths capped by the best odd one, the shorter of the two returned) with fir_qprog_phs probes (:100,142).
if nargin < 5, even_odd = 0; end
if nargin < 6, dbg = 0; end
[h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg, false, @fir_qprog_phs);
end

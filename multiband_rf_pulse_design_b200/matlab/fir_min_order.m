function [h, status] = fir_min_order(n, f, a, d, even_odd, a_min, dbg) %#ok<INUSL>
%FIR_MIN_ORDER  Drop-in for ss/fir_min_order.m:55-230.  The reference probes with fir_pm -> cfirpm (closed-source
%  Parks-McClellan, ss/fir_pm.m:175); here the same bisection -- and the same selection of the LONGER of the odd / even
%  answers (:222-226) -- runs on linear-programming feasibility probes (fir_linprog on the GPU).  a_min is fir_pm's
%  minimum-amplitude option and has no LP counterpart: accepted and ignored.
if nargin < 5, even_odd = []; end
if nargin < 7, dbg = 0; end
[h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg, true);
end

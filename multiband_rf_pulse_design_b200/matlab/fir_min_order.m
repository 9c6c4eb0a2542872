function [h, status] = fir_min_order(n, f, a, d, even_odd, a_min, dbg)
%FIR_MIN_ORDER  Drop-in for ss/fir_min_order.m:55-230.  The reference probes with fir_pm -> cfirpm (closed-source
%  Parks-McClellan, ss/fir_pm.m:175); here the same bisection -- and the same selection of the LONGER of the odd / even
%  answers (:222-226) -- runs on linear-programming feasibility probes (fir_linprog on the GPU).  a_min is fir_pm's lower
%  bound of the response in the transition regions (ss/fir_pm.m:42-43,102; default min(0, min(a - d)), which is also the LP's
%  default): a given value replaces that bound in every probe.
if nargin < 5, even_odd = []; end
if nargin < 6, a_min = []; end
if nargin < 7, dbg = 0; end
[h, status] = fir_min_order_linprog(n, f, a, d, even_odd, dbg, true, @(nt, f_, a_, d_, h0, dbg_) fir_linprog(nt, f_, a_, d_, h0, dbg_, a_min));
end

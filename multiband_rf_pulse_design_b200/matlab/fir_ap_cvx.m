function [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
% fir_ap_cvx - drop-in for the reference's fir_ap_cvx.m with the CVX solve replaced by libmbrf's
% batched restarted PDHG on the GPU (fir_pdhg_mex).  Same signature, defaults, status strings and
% spectral factorisation; the problem is assembled exactly as fir_ap_cvx.m:44-142 does.
%
%   [h, status] = fir_ap_cvx(n, f, a, d, obj, Peak, dbg)
if nargin < 4,   error('not enough input');  end;
if nargin <= 4,  obj = 0;     end;
if nargin <= 5,  Peak = 1e-3; end;
if nargin <= 6,  dbg = 0;     end; %#ok<NASGU>
if obj < 0, error('invalid input of obj'); end;

f = f(:).' * pi;  a = a(:).';  d = d(:).';
oversamp = 15;  m = 2 * n * oversamp;
w = sort([linspace(-pi,pi,m) f]);
idx_band = []; U_band = []; L_band = [];
for band = 1:length(f)/2,
    idx = find( (w >= f(band*2-1)) & (w <= f(band*2)) );
    idx_band = [idx_band idx]; %#ok<AGROW>
    if (f(band*2-1) == f(band*2)), amp = a(band*2-1) * ones(size(idx));
    else amp = a(band*2-1) + (a(band*2)-a(band*2-1)) * ((w(idx) - f(band*2-1))/(f(band*2)-f(band*2-1))); end;
    U_band = [U_band (amp + d(band))]; L_band = [L_band (amp - d(band))]; %#ok<AGROW>
end;
idx_tmp = ones(1,length(w)); idx_tmp(idx_band) = 0; idx_tran = find(idx_tmp == 1);
U_tran = max(U_band)*ones(1,length(idx_tran)); L_tran = min(0, min(L_band))*ones(1,length(idx_tran));
w = [w(idx_band) w(idx_tran)];  m = length(w);
U_b = [U_band U_tran].^2;  L_b = [L_band L_tran];  L_b(L_b < 0) = 0;  L_b = L_b.^2;  L_b(L_b < 1e-20) = 1e-20;
idx_stop = find( sqrt(U_b) < (min(sqrt(U_b))+1e-2) );  ns = length(idx_stop);

% canonical form: z = x (2n-1); the stop rows are appended a second time as the solver's "simplex block":
% obj*ripple_stop with A_U(idx_stop,:)*x <= ripple_stop  ==  obj * max_i (A x)_i over idx_stop
w_row = [w w(idx_stop)];
col_type = [0 ones(1,n-1) 2*ones(1,n-1)];  col_kappa = [0 1:n-1 1:n-1];  col_amp = [1 2*ones(1,2*n-2)];
N = 2*n-1;  c = zeros(N,1);  c(1) = 1;
lo = [L_b -inf(1,ns)].';  hi = [U_b zeros(1,ns)].';
bl = -inf(N,1);  bu = inf(N,1);  bl(1) = -n*Peak;  bu(1) = n*Peak;  tmax = max(U_b(idx_stop));
rho = ((n-(2:n)+1) * Peak).';
[z, info] = fir_pdhg_mex(w_row, [], col_type, col_kappa, col_amp, 0, 2:n, n+1:2*n-1, c, lo, hi, bl, bu, rho, ...
                         n*Peak + obj*tmax, [200000 64 8e-7 1e-4 5e-5], [m+1 ns obj]);
if info(1) == 1, status = 'Solved'; else status = 'Failed'; h = []; return; end;
x = z(1:2*n-1);
r = [x(1);  x(2:n) + sqrt(-1) * x((n+1):(2*n-1))];  r = [conj(r(end:-1:2)); r];
h = fmp2_local(r);
end

function hmp = fmp2_local(h)    % fir_ap_cvx.m:262-283
h = transpose(h(:));  l = length(h);
lp = 8*exp(ceil(log(l)/log(2))*log(2));
hp = [zeros(1,ceil((lp-l)/2)) h zeros(1,floor((lp-l)/2))];
hpf = fftshift(fft(fftshift(hp)));
x = sqrt(abs(hpf));  nn = length(x);  xlf = fft(log(x));
xlfp = zeros(1,nn);  xlfp(1) = xlf(1);  xlfp(2:(nn/2)) = 2*xlf(2:(nn/2));  xlfp(nn/2+1) = xlf(nn/2+1);
hpmp = ifft(fftshift(conj(exp(ifft(xlfp)))));
hmp = hpmp(1:(l+1)/2);
end

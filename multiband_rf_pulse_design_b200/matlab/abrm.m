function [a b] = abrm(rf,g,x,y)
%  [a b] = abrm(rf,[g],[x [,y])
%
%  Drop-in for rf_tools/abrm.m: same arguments, defaults and output convention
%  (abrm.m:26-64), computed on the GPU by libmbrf through the abrx MEX gateway
%  (convention 1 = abrm).  Like the original there is no phi == 0 guard: such a
%  sample yields NaN.
if (nargin == 2),
  x = g;
  g = ones(1,length(rf))*2*pi/length(rf);
  y = 0;
elseif (nargin == 3),
  y = 0;
end;
[a, b] = abrx(rf(:).', complex(real(g(:).'), imag(g(:).')), x(:).', y(:).', 1);
if (nargout == 1),
  a = [a b];
end;

function [a, b] = abrm(rf, g, x, y)
%ABRM  Drop-in for rf_tools/abrm.m on top of libmbrf (GPU):  [a b] = abrm(rf, [g,] x [, y])
%
%  Same call forms and output convention as the toolbox function: with two arguments the second one is the
%  position vector and the gradient defaults to one full cycle over the pulse (2*pi/N per sample); a missing y means a
%  one-dimensional profile.  The recursion itself runs in the abrx MEX gateway with convention 1 (the abrm form of
%  alpha/beta).  As in the original, a sample whose total rotation angle is zero produces NaN.
switch nargin
  case 2
    pos = g;
    npts = numel(rf);
    grad = repmat(2*pi/npts, 1, npts);
    ypos = 0;
  case 3
    pos = x;  grad = g;  ypos = 0;
  otherwise
    pos = x;  grad = g;  ypos = y;
end
grad = reshape(grad, 1, []);
[a, b] = abrx(reshape(rf, 1, []), complex(real(grad), imag(grad)), reshape(pos, 1, []), reshape(ypos, 1, []), 1);
if nargout < 2
  a = [a b];       % single output: alpha and beta side by side
end

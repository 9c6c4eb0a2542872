function h_new = fir_flip_zero(h, dbg)
%FIR_FLIP_ZERO  Drop-in for the toolbox's fir_flip_zero.m: among the filters that share |H| with h (passband zeros reflected
%  about the unit circle in every tried combination), return the one with the smallest peak amplitude.
%  The reference expands each candidate with poly() in a MATLAB loop (fir_flip_zero.m:66-93); here all candidates are expanded
%  at once on the GPU (flip_zero_mex -> mbrf_flip_zero_batch), one thread block per pattern.  roots() and the pattern table
%  stay here.  Pattern order, the 2^12 cap and the random sampling above 12 passband zeros follow :45-64.
if nargin < 2, dbg = 0; end
Z = roots(h);                                                                      % :25
off = abs(Z) > 1 + 1e-2 | abs(Z) < 1 - 1e-2;                                       % passband zeros, :28
idx_pb = find(off);
nz = numel(idx_pb);
cap = 2^12;
if nz == 0
    mask = zeros(0, 1);
elseif nz <= 19
    cols = 0:2^nz - 1;                                                             % column c of combination_2power (:119-138):
    if nz > 12                                                                     % zero r flips iff bit nz-r of c is clear
        cols = sort(randperm(2^nz, cap)) - 1;                                      % :55-58
    end
    mask = zeros(nz, numel(cols));
    for r = 1:nz
        mask(r, :) = 1 - bitand(bitshift(cols, -(nz - r)), 1);
    end
else
    mask = round(rand(nz, cap));                                                   % :60-63
end
[h_new, ~, ~, ~] = flip_zero_mex(Z, idx_pb, mask, sum(h));
if isreal(h)                                                                       % poly() returns real coefficients for
    sel = logical(mask(:, 1) * 0);                                                 % conjugate-closed zeros
    [~, best] = flip_zero_mex(Z, idx_pb, mask, sum(h));
    if nz, sel = logical(mask(:, best)); end
    Zb = Z;   Zb(idx_pb(sel)) = (1 ./ abs(Z(idx_pb(sel)))) .* exp(1i * angle(Z(idx_pb(sel))));
    if isequal(sort(Zb(imag(Zb) > 0)), sort(conj(Zb(imag(Zb) < 0)))), h_new = real(h_new); end
end
if dbg >= 1                                                                        % :101-103
    fprintf('reduce peak amplitude from %6.4f to %6.4f by %6.4f\n', max(abs(h)), max(abs(h_new)), ...
            (max(abs(h)) - max(abs(h_new))) / max(abs(h)));
end
end

function [rot, trans, state] = find_transform_matrix(pset1, pset2)
% Signature of mex_files/RANSAC_CALCULATION/find_transform_matrix.m:2 of 3PRE, computed by libpre3.so.
[rot, trans, state] = find_transform_matrix_mex(pset1, pset2);
end

function [hypothesis_support, positions_li_inliers_id, positions_li_inliers_euc] = compute_hypothesis_support_fast(xi, cam, state_vector_pattern, z_id, z_euc, threshold)
% Same signature as compute_hypothesis_support_fast.m:27 of 3PRE; runs on the GPU through
% compute_hypothesis_support_fast_mex (libpre3.so).  Put this directory before the reference's on the path.
[hypothesis_support, positions_li_inliers_id, positions_li_inliers_euc] = ...
    compute_hypothesis_support_fast_mex(xi, cam, state_vector_pattern, z_id, z_euc, threshold);
end

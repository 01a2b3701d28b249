function features_info = ransac_hypotheses(filter, features_info, cam, varargin)
% Same signature as ransac_hypotheses.m:27 of 3PRE.  The hypothesis loop (select -> 1/3-match EKF update
% -> support -> adaptive stop) runs on the GPU through ransac_hypotheses_mex (libpre3.so); this shim only
% unpacks the @ekf_filter object and the features_info struct array (fields are reachable through the
% get_* methods only) and writes low_innovation_inlier / StatData back like the reference does.
% Optional extras: ransac_hypotheses(..., sel_or_seed, n_hyp, adaptive), sel = 3 x H 1-based positions.
global StatData
x = get_x_k_km1(filter); P = get_p_k_km1(filter); std_z = get_std_z(filter);
F = length(features_info); n = length(x);
type = zeros(1, F); pos = zeros(1, F); has_z = zeros(1, F); ic = zeros(1, F); li0 = zeros(1, F);
z = zeros(2, F); h = zeros(2, F); Hcam = zeros(2, 13, F); Hfeat = zeros(2, 6, F); R = zeros(2, 2, F);
position = 14;                                   % generate_state_vector_pattern.m:30
for i = 1:F
    if strcmp(features_info(i).type, 'cartesian'), type(i) = 1; nf = 3; else, nf = 6; end
    pos(i) = position; position = position + nf;
    has_z(i) = ~isempty(features_info(i).z);
    ic(i) = isequal(features_info(i).individually_compatible, 1);
    if isfield(features_info, 'low_innovation_inlier') && ~isempty(features_info(i).low_innovation_inlier)
        li0(i) = features_info(i).low_innovation_inlier;
    end
    if has_z(i)
        z(:, i) = features_info(i).z(1:2); h(:, i) = features_info(i).h(:);
        Hi = full(features_info(i).H);
        Hcam(:, :, i) = Hi(:, 1:13); Hfeat(:, 1:nf, i) = Hi(:, pos(i):pos(i) + nf - 1);
        Hi(:, 1:13) = 0; Hi(:, pos(i):pos(i) + nf - 1) = 0;
        if any(Hi(:)), error('pre3:ekf', 'H has non-zeros outside the camera and the feature''s own block'); end
        R(:, :, i) = features_info(i).R;
    end
end
if position - 1 ~= n, error('pre3:ekf', 'state size does not match the features'); end
[li, stats] = ransac_hypotheses_mex(x, P, std_z, cam, type, pos, has_z, ic, z, h, Hcam, Hfeat, R, li0, varargin{:});
if stats(3) > 0                                  % set_as_most_supported_hypothesis.m:32-53
    for i = 1:F
        if has_z(i), features_info(i).low_innovation_inlier = li(i); end
    end
end
disp(['FOR DEBUG - RANSAC !PRE: ', num2str(stats(1))])   % ransac_hypotheses.m:83
StatData.RANSAC_ITER = stats(1);
StatData.RANSAC_HYP_SUPPORT = stats(2);
end

function [rot, phi, theta, psi, trans, error, pnum, op_num, sta, varargout] = vodometry_dr_ye(file1, file2, varargin)
% Same signature and outputs as code_from_dr_ye/vodometry_dr_ye.m:5 of 3PRE.  Reading the two SR4000 frames,
% SIFT extraction and the confidence filter are the reference's own functions (they are not on the GPU path);
% descriptor matching goes through the siftmatch MEX of this directory and the RANSAC iterations, selection,
% refit and residual statistics (vodometry_dr_ye.m:147-220) through vodometry_dr_ye_mex (libpre3.so).
% Put this directory before the reference's on the MATLAB path.
% Optional name/value extras after file2: 'samples' (4 x H, 1-based draws) or 'seed'.
global myCONFIG
extra = {};
i = 1;
while i <= numel(varargin)
    if ischar(varargin{i}) && any(strcmpi(varargin{i}, {'samples', 'seed'}))
        extra = {varargin{i + 1}};
        i = i + 2;
    else
        i = i + 1;  % the reference's debug flag: plotting is not reproduced here
    end
end
error = 0;
RANSAC_STAT = struct('nFeatures1', 0, 'nF1_Confidence_Filtered', 0, 'nFeatures2', 0, 'nF2_Confidence_Filtered', 0, ...
    'nMatches', 0, 'nIterationRansac', 0, 'InlierRatio', 0, 'nSupport', 0, 'ErrorMean', 0, 'ErrorVariance', 0, ...
    'SolutionState', 0);
[x1, y1, z1, confidence_map1, img1] = read_sr4000_data_dr_ye(file1);
[frm1, des1] = sift(img1);
RANSAC_STAT.nFeatures1 = size(frm1, 2);
frm1(1:2, :) = frm1(1:2, :) + 1;
if myCONFIG.FLAGS.CONFIDENCE_MAP
    [frm1, des1] = confidence_filtering(frm1, des1, confidence_map1);
end
RANSAC_STAT.RawFrames1 = frm1;
RANSAC_STAT.RawDescriptor1 = des1;
RANSAC_STAT.nF1_Confidence_Filtered = size(frm1, 2);
[x2, y2, z2, confidence_map2, img2] = read_sr4000_data_dr_ye(file2);
[frm2, des2] = sift(img2);
RANSAC_STAT.nFeatures2 = size(frm2, 2);
frm2(1:2, :) = frm2(1:2, :) + 1;
if myCONFIG.FLAGS.CONFIDENCE_MAP
    [frm2, des2] = confidence_filtering(frm2, des2, confidence_map2);
end
RANSAC_STAT.RawFrames2 = frm2;
RANSAC_STAT.RawDescriptor2 = des2;
RANSAC_STAT.nF2_Confidence_Filtered = size(frm2, 2);
match = siftmatch(des1, des2);
pnum = size(match, 2);
RANSAC_STAT.nMatches = pnum;
rot = zeros(3); phi = 0.0; theta = 0.0; psi = 0.0; trans = 0.0; op_num = 0; sta = 0;
varargout = {[], [], RANSAC_STAT};
if pnum < 4
    fprintf('too few sift points for ransac.\n');
    error = 1;
    RANSAC_STAT.SolutionState = 4;
    varargout{3} = RANSAC_STAT;
    return;
end
pset1 = lookup_points(frm1, match(1, :), x1, y1, z1);
pset2 = lookup_points(frm2, match(2, :), x2, y2, z2);
[rot, trans, sta, op_num, good, st] = vodometry_dr_ye_mex(pset1, pset2, match, extra{:});
RANSAC_STAT.nIterationRansac = st.nIterationRansac;
if sta == 4
    fprintf('no consensus found, ransac fails.\n');
    RANSAC_STAT.SolutionState = 4;
    phi = 0.0; theta = 0.0; psi = 0.0; trans = 0.0;
    varargout{3} = RANSAC_STAT;
    return;
end
op_match = match(:, good);
op_pset1 = pset1(:, good);
op_pset2 = pset2(:, good);
RANSAC_STAT.nSupport = op_num;
RANSAC_STAT.ErrorMean = st.ErrorMean;
RANSAC_STAT.ErrorStd = st.ErrorStd;
RANSAC_STAT.SolutionState = sta;
RANSAC_STAT.GoodFrames1 = frm1(:, op_match(1, :));
RANSAC_STAT.GoodDescriptor1 = des1(:, op_match(1, :));
RANSAC_STAT.GoodFrames2 = frm2(:, op_match(2, :));
RANSAC_STAT.GoodDescriptor2 = des2(:, op_match(2, :));
RANSAC_STAT.InlierRatio = (op_num / pnum) * 100;
varargout = {op_pset1, op_pset2, RANSAC_STAT};
if sta < 1
    error = 2;
    phi = 0.0; theta = 0.0; psi = 0.0; trans = 0.0;
else
    e_ = R2e(rot);
    phi = e_(1); theta = e_(2); psi = e_(3);
end
if sta == -1
    disp('??????? RANSAC FAILED ?????')
end
end

function pset = lookup_points(frm, idx, x, y, z)
% [-x(ROW,COL); -y(ROW,COL); z(ROW,COL)] at the rounded frame position of every matched feature
col = round(frm(1, idx));
row = round(frm(2, idx));
lin = sub2ind(size(x), row, col);
pset = [-x(lin); -y(lin); z(lin)];
pset = reshape(pset, 3, []);
end

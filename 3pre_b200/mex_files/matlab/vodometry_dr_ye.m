function [rot, phi, theta, psi, trans, error, pnum, op_num, sta, varargout] = vodometry_dr_ye(file1, file2, varargin)
% Drop-in for code_from_dr_ye/vodometry_dr_ye.m:5 of 3PRE (same inputs, same twelve outputs, same RANSAC_STAT fields).
% What stays in MATLAB: reading the two SR4000 frames, SIFT extraction and the confidence filter (the reference's own
% read_sr4000_data_dr_ye / sift / confidence_filtering -- none of them is on the GPU path).  What runs in libpre3.so:
% descriptor matching (the siftmatch MEX of this directory) and everything from the RANSAC iterations to the residual
% statistics (vodometry_dr_ye.m:147-220) in ONE call of vodometry_dr_ye_mex.
% Put this directory before the reference's on the MATLAB path.
% Optional trailing name/value pair: 'samples', S (4 x H, 1-based draws) or 'seed', s.  Any other trailing argument is
% taken as the reference's debug flag; its figures are not reproduced.
global myCONFIG
rng_arg = {};
for a = 1:numel(varargin)
    if ischar(varargin{a}) && a < numel(varargin) && any(strcmpi(varargin{a}, {'samples', 'seed'}))
        rng_arg = varargin(a + 1);
    end
end
use_conf = myCONFIG.FLAGS.CONFIDENCE_MAP;
A = prepare_frame(file1, use_conf);
B = prepare_frame(file2, use_conf);
match = siftmatch(A.des, B.des);
pnum = size(match, 2);

% defaults = the values the reference returns on its two failure paths (:152-160, :187-194)
rot = zeros(3); phi = 0.0; theta = 0.0; psi = 0.0; trans = 0.0; error = 0; op_num = 0; sta = 0;
stat = struct('nFeatures1', A.n_raw, 'nF1_Confidence_Filtered', size(A.frm, 2), ...
              'nFeatures2', B.n_raw, 'nF2_Confidence_Filtered', size(B.frm, 2), ...
              'nMatches', pnum, 'nIterationRansac', 0, 'InlierRatio', 0, 'nSupport', 0, ...
              'ErrorMean', 0, 'ErrorVariance', 0, 'SolutionState', 0, ...
              'RawFrames1', A.frm, 'RawDescriptor1', A.des, 'RawFrames2', B.frm, 'RawDescriptor2', B.des);
support1 = [];
support2 = [];
if pnum < 4
    fprintf('too few sift points for ransac.\n');
    error = 1;
    stat.SolutionState = 4;
else
    P1 = points_of(A, match(1, :));
    P2 = points_of(B, match(2, :));
    [rot, trans, sta, op_num, good, core] = vodometry_dr_ye_mex(P1, P2, match, rng_arg{:});
    stat.nIterationRansac = core.nIterationRansac;
    if sta == 4
        fprintf('no consensus found, ransac fails.\n');
        stat.SolutionState = 4;
        trans = 0.0;
    else
        support1 = P1(:, good);
        support2 = P2(:, good);
        stat.nSupport = op_num;
        stat.ErrorMean = core.ErrorMean;
        stat.ErrorStd = core.ErrorStd;
        stat.SolutionState = sta;
        stat.GoodFrames1 = A.frm(:, match(1, good));
        stat.GoodDescriptor1 = A.des(:, match(1, good));
        stat.GoodFrames2 = B.frm(:, match(2, good));
        stat.GoodDescriptor2 = B.des(:, match(2, good));
        stat.InlierRatio = 100 * op_num / pnum;
        if sta < 1
            error = 2;
            trans = 0.0;
            if sta == -1
                disp('??????? RANSAC FAILED ?????')
            end
        else
            eul = R2e(rot);
            phi = eul(1); theta = eul(2); psi = eul(3);
        end
    end
end
varargout = {support1, support2, stat};
end

function F = prepare_frame(file, use_conf)
% one SR4000 frame: maps, SIFT frames shifted to 1-based pixel positions, optional confidence filter
[F.x, F.y, F.z, conf, img] = read_sr4000_data_dr_ye(file);
[F.frm, F.des] = sift(img);
F.n_raw = size(F.frm, 2);
F.frm(1:2, :) = F.frm(1:2, :) + 1;
if use_conf
    [F.frm, F.des] = confidence_filtering(F.frm, F.des, conf);
end
end

function P = points_of(F, idx)
% 3 x numel(idx): [-x; -y; z] of the maps at the rounded positions of features idx (camera coordinates)
lin = sub2ind(size(F.x), round(F.frm(2, idx)), round(F.frm(1, idx)));
P = [-F.x(lin); -F.y(lin); F.z(lin)];
P = reshape(P, 3, []);
end

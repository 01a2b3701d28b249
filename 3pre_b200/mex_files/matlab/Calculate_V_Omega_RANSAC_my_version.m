function [T, q, R, State_RANSAC] = Calculate_V_Omega_RANSAC_my_version(stepPre, stepCurrent)
% Signature of Calculate_V_Omega_RANSAC_my_version.m:1 of 3PRE.  The on-disk memoisation
% (RANSAC5_step_%d_%d.mat, :4-20) stays as in the reference; on a cache miss the pair is solved
% by the GPU drop-ins (siftmatch MEX + RANSAC_CALC_VER2 MEX) through the reference's own
% RANSAC_CALC_SAVE_SR4000 / SIFT_match_save, which call `siftmatch` and `RANSAC_CALC_VER2` by name.
global myCONFIG
file = sprintf('%s/RANSAC_pose_shift/RANSAC5_step_%d_%d.mat', myCONFIG.PATH.DATA_FOLDER, stepPre, stepCurrent);
if ~exist(file, 'file') || myCONFIG.FLAGS.RECALCULATE
    RANSAC_CALC_SAVE_SR4000(stepPre, stepCurrent);
end
S = load(file, 'T_RANSAC', 'R_RANSAC', 'State_RANSAC');
T = S.T_RANSAC; R = S.R_RANSAC; State_RANSAC = S.State_RANSAC;
q = R2q(R);
end

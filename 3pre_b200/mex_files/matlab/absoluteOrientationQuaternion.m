function [s, R, T, err] = absoluteOrientationQuaternion(A, B, doScale)
% Signature of absoluteOrientationQuaternion.m:28 of 3PRE, computed by libpre3.so.
if nargin < 3
    doScale = 1;  % the reference's default (:32-34)
end
[s, R, T, err] = absoluteOrientationQuaternion_mex(A, B, doScale);
end

function [R, T, error, BestFit, State_RANSAC] = RANSAC_CALC_VER2(Ya, Yb, options, varargin)
% Same signature as mex_files/RANSAC_CALCULATION/RANSAC_CALC_VER2.m:2 of 3PRE; the loop runs on
% the GPU through RANSAC_CALC_VER2_mex (libpre3.so).  Put this directory before the reference's
% on the MATLAB path.  varargin{1:2} (Za, Zb) are accepted and ignored like a pass-through;
% optional name/value extras: 'samples' (k x H, 1-based), 'seed', 'k', 'adaptive', 'method'.
extra = {[], [], [], []};  % samples|seed, k, adaptive, method
i = 1;
while i <= numel(varargin)
    if ischar(varargin{i})
        switch lower(varargin{i})
            case {'samples', 'seed'}, extra{1} = varargin{i + 1};
            case 'k',                 extra{2} = varargin{i + 1};
            case 'adaptive',          extra{3} = varargin{i + 1};
            case 'method',            extra{4} = varargin{i + 1};
        end
        i = i + 2;
    else
        i = i + 1;  % Za, Zb
    end
end
[R, T, error, BestFit, State_RANSAC] = RANSAC_CALC_VER2_mex(Ya, Yb, options, extra{:});
end

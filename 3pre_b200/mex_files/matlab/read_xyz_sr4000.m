function [x, y, z, confidence_map, varargout] = read_xyz_sr4000(scan_file_name_prefix, my_k)
% Drop-in for read_xyz_sr4000.m:1 of 3PRE: same inputs and outputs (incl. the optional fifth output, the time stamp in
% row 721 of the file, -1 when absent) and the same xyz_%04d.mat side cache.  The three 3 x 3 Gaussian filters run on
% the GPU (read_xyz_sr4000_mex -> libpre3.so).  Put this directory before the reference's on the MATLAB path.
frame = load(fullfile(scan_file_name_prefix, sprintf('d1_%04d.dat', my_k)));
[x, y, z, confidence_map] = read_xyz_sr4000_mex(frame(:, 1:176));
if nargout > 4
    stamp = -1;
    if size(frame, 1) == 721
        stamp = frame(721, 1);
    end
    varargout{1} = stamp;
end
cache = fullfile(scan_file_name_prefix, 'xyz_data', sprintf('xyz_%04d.mat', my_k));
if ~exist(cache, 'file')
    save(cache, 'x', 'y', 'z', 'confidence_map')
end
end

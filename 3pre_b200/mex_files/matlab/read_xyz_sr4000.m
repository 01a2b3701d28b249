function [x, y, z, confidence_map, varargout] = read_xyz_sr4000(scan_file_name_prefix, my_k)
% Same signature as read_xyz_sr4000.m:1 of 3PRE.  The file is loaded as in the reference (:2-3); the three Gaussian
% filters (:8-21) run on the GPU through read_xyz_sr4000_mex (libpre3.so); the time-stamp output (:36-43) and the
% xyz_%04d.mat cache (:49-52) are kept.  Put this directory before the reference's on the MATLAB path.
s = sprintf('%s/d1_%04d.dat', scan_file_name_prefix, my_k);
sr_data = load(s);
[x, y, z, confidence_map] = read_xyz_sr4000_mex(sr_data(:, 1:176));
if nargout == 5
    if size(sr_data, 1) == 721
        varargout{1} = sr_data(721, 1);
    else
        varargout{1} = -1;
    end
end
s1 = sprintf('%s/xyz_data/xyz_%04d.mat', scan_file_name_prefix, my_k);
if ~exist(s1, 'file')
    save(s1, 'x', 'y', 'z', 'confidence_map')
end
end

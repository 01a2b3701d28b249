"""Independent numpy / LAPACK restatement of the RANSAC path ("oracle A").

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/pre3_oracle.c: it uses
LAPACK svd / eigh the way the MATLAB reference uses svd / eig, so it checks the
Jacobi-based C oracle and the CUDA kernels to the north-star tolerance
(1e-9 rad / 1e-9 m) rather than bit-for-bit.  Arrays: points are (N,3).
`M/` = /root/reference/matlab_code/.
"""
from __future__ import annotations

import numpy as np


def find_transform_matrix(pset1, pset2, threshold=1e-11):
    """M/mex_files/RANSAC_CALCULATION/find_transform_matrix.m:9-42 (pset1 ~ rot*pset2 + trans);
    threshold=1e-14 is M/code_from_dr_ye/find_transform_matrix_dr_ye.m."""
    p1 = np.asarray(pset1, float).T  # 3 x n
    p2 = np.asarray(pset2, float).T
    n = p2.shape[1]
    ct1 = p1.sum(1) / n
    ct2 = p2.sum(1) / n
    q1 = p1 - ct1[:, None]
    q2 = p2 - ct2[:, None]
    H = q2 @ q1.T
    U, S, Vt = np.linalg.svd(H)
    V = Vt.T
    sv = np.abs(S)
    Xq = V @ U.T
    mdet = np.linalg.det(Xq)
    if round(mdet) == 1:
        rot = Xq
        return rot, ct1 - rot @ ct2, 1
    if round(mdet) == -1:
        zn = np.nonzero(sv < threshold)[0]
        if zn.size == 1:
            V[:, zn] = -V[:, zn]
            rot = V @ U.T
            return rot, ct1 - rot @ ct2, 2
        return H, np.zeros(3), -1
    return H, np.zeros(3), 0


def horn(A, B, do_scale=True):
    """M/absoluteOrientationQuaternion.m:56-127 (B ~ s*R*A + T) with the eigenvector of
    the LARGEST eigenvalue."""
    A = np.asarray(A, float).T
    B = np.asarray(B, float).T
    n = A.shape[1]
    Ca = A.mean(1)
    Cb = B.mean(1)
    An = A - Ca[:, None]
    Bn = B - Cb[:, None]
    M = np.zeros((4, 4))
    for i in range(n):
        a = np.r_[0.0, An[:, i]]
        b = np.r_[0.0, Bn[:, i]]
        Ma = np.array([[a[0], -a[1], -a[2], -a[3]], [a[1], a[0], a[3], -a[2]], [a[2], -a[3], a[0], a[1]], [a[3], a[2], -a[1], a[0]]])
        Mb = np.array([[b[0], -b[1], -b[2], -b[3]], [b[1], b[0], -b[3], b[2]], [b[2], b[3], b[0], -b[1]], [b[3], -b[2], b[1], b[0]]])
        M += Ma.T @ Mb
    w, E = np.linalg.eigh(0.5 * (M + M.T))
    e = E[:, np.argmax(w)]
    M1 = np.array([[e[0], -e[1], -e[2], -e[3]], [e[1], e[0], e[3], -e[2]], [e[2], -e[3], e[0], e[1]], [e[3], e[2], -e[1], e[0]]])
    M2 = np.array([[e[0], -e[1], -e[2], -e[3]], [e[1], e[0], -e[3], e[2]], [e[2], e[3], e[0], -e[1]], [e[3], -e[2], e[1], e[0]]])
    R = (M1.T @ M2)[1:, 1:]
    if do_scale:
        a = sum(Bn[:, i] @ R @ An[:, i] for i in range(n))
        b = sum(Bn[:, i] @ Bn[:, i] for i in range(n))
        s = b / a
    else:
        s = 1.0
    T = Cb - s * R @ Ca
    err = sum(np.linalg.norm(B[:, i] - (s * R @ A[:, i] + T)) for i in range(n))
    return s, R, T, err


def score(R, T, Ya, Yb, thr):
    """RANSAC_CALC_VER2.m:121-125."""
    Ya = np.asarray(Ya, float)
    Yb = np.asarray(Yb, float)
    res = Yb @ np.asarray(R).T + np.asarray(T)[None, :] - Ya
    nr = np.sqrt(res[:, 0] ** 2 + res[:, 1] ** 2 + res[:, 2] ** 2)
    mask = nr < thr
    return int(mask.sum()), mask, float(nr[mask].sum()), nr


def ransac_ver2(Ya, Yb, samples, max_iteration=2000, adaptive=True, method=0, distance_threshold=0.05):
    """RANSAC_CALC_VER2.m:43-201 with supplied sample sets (H,k) 0-based."""
    Ya = np.asarray(Ya, float)
    Yb = np.asarray(Yb, float)
    N = Ya.shape[0]
    k = samples.shape[1]
    if method == 0:
        j = int(np.argmin(Yb[:, 2]))
        thr = 0.01 * np.sqrt(Yb[j, 0] ** 2 + Yb[j, 1] ** 2 + Yb[j, 2] ** 2)
        mult = 5
    else:
        thr = distance_threshold
        mult = 1
    n_iter = float(max_iteration)
    max_support = 5
    it = 1
    rec = []
    h = 0
    while it < min(n_iter, max_iteration) and h < samples.shape[0]:
        s = samples[h]
        h += 1
        if method == 0:
            Rm, Tm, st = find_transform_matrix(Ya[s], Yb[s])
            if st == -1:
                continue
        else:
            _, Rm, Tm, _ = horn(Yb[s], Ya[s], False)
        c, mask, es, _ = score(Rm, Tm, Ya, Yb, thr)
        rec.append((c, es, h - 1, Rm, Tm, mask))
        if c >= max_support:
            max_support = c
            if adaptive:
                w = (c / N) ** k
                with np.errstate(divide="ignore"):
                    n_iter = mult * np.ceil(np.log(0.01) / np.log(1 - w)) if w < 1 else 0.0
        it += 1
    card = np.array([r[0] for r in rec])
    es = np.array([r[1] for r in rec], float)
    mx = card.max()
    e1 = np.where((card != mx) | (card == 0), 10000.0, es)
    best = int(np.argmin(e1))
    c, _, hs, Rm, Tm, mask = rec[best]
    if method == 0:
        R, T, st = find_transform_matrix(Ya[mask], Yb[mask])
    else:
        _, R, T, _ = horn(Yb[mask], Ya[mask], False)
        st = 1
    return dict(R=R, T=T, state=st, best_fit=int(mx), best_sample=hs, mask=mask, n_iter=len(rec), thr=thr, error_sum=es[best])


def dr_ye_sampler(match, rand):
    """M/code_from_dr_ye/ransac_dr_ye.m:28-48, line by line.  match: (pnum,2) feature ids; rand(): a callable
    returning the next uniform of the stream.  Returns the four 0-based match indices in draw order."""
    m = np.asarray(match).T  # 2 x pnum, as in the reference
    pnum = m.shape[1]

    def draw():
        v = (pnum - 1) * rand() + 1
        return int(np.floor(v + 0.5)) - 1  # MATLAB round (half away from zero, v > 0), 0-based

    n = [draw() for _ in range(4)]
    d1 = (m[0, n[0]] == m[0, n[1]]) or (m[1, n[0]] == m[1, n[1]])
    d2 = ((m[0, n[0]] == m[0, n[2]]) or (m[0, n[1]] == m[0, n[2]]) or (m[1, n[0]] == m[1, n[2]])
          or (m[1, n[1]] == m[1, n[2]]))
    d3 = ((m[0, n[0]] == m[0, n[3]]) or (m[0, n[1]] == m[1, n[3]]) or (m[0, n[2]] == m[0, n[3]])
          or (m[1, n[0]] == m[0, n[3]]) or (m[1, n[1]] == m[1, n[3]]) or (m[1, n[2]] == m[1, n[3]]))
    while n[1] == n[0] or d1:
        n[1] = draw()
        d1 = (m[0, n[0]] == m[0, n[1]]) or (m[1, n[0]] == m[1, n[1]])
    while n[2] == n[0] or n[2] == n[1] or d2:
        n[2] = draw()
        d2 = ((m[0, n[0]] == m[0, n[2]]) or (m[0, n[1]] == m[0, n[2]]) or (m[1, n[0]] == m[1, n[2]])
              or (m[1, n[1]] == m[1, n[2]]))
    while n[3] == n[0] or n[3] == n[1] or n[3] == n[2] or d3:
        n[3] = draw()
        d3 = ((m[0, n[0]] == m[0, n[3]]) or (m[0, n[1]] == m[1, n[3]]) or (m[0, n[2]] == m[0, n[3]])
              or (m[1, n[0]] == m[0, n[3]]) or (m[1, n[1]] == m[1, n[3]]) or (m[1, n[2]] == m[1, n[3]]))
    return n


def vodometry_dr_ye(Ya, Yb, samples, max_iteration=700):
    """RANSAC part of M/code_from_dr_ye/vodometry_dr_ye.m:147-220 with ransac_dr_ye.m:20-71 as the loop body;
    Ya = pset1, Yb = pset2 (N,3); samples (H,4) supplied draws.  LAPACK svd for the fits.  Returns a dict."""
    from math import comb
    Ya = np.asarray(Ya, float)
    Yb = np.asarray(Yb, float)
    pnum = Ya.shape[0]
    out = {"status": 0, "op_num": 0, "best_sample": -1}
    if pnum < 4:
        out["status"] = 1
        return out
    nrm = np.sqrt(Yb[:, 2] ** 2 + Yb[:, 1] ** 2 + Yb[:, 0] ** 2)
    far = nrm > 0.4
    if not far.any():
        out["status"] = 5
        return out
    minZ = Yb[far, 2].min()
    j = np.nonzero(Yb[:, 2] == minZ)[0][0]
    dist = np.sqrt(Yb[j, 0] ** 2 + Yb[j, 1] ** 2 + Yb[j, 2] ** 2)
    rst = min(max_iteration, comb(pnum, 4))
    L = min(rst, len(samples))
    cnums = np.full(len(samples), -1, np.int64)
    maxc, nit = 0, rst
    for i in range(L):
        s = np.asarray(samples[i])
        rot, trans, _ = find_transform_matrix(Ya[s], Yb[s])
        d = ((Yb @ rot.T + trans - Ya) ** 2).sum(1)
        cnums[i] = int((d < 0.001 * dist).sum())
        if cnums[i] > maxc:
            maxc = cnums[i]
            with np.errstate(divide="ignore"):
                nit = 5 * np.ceil(np.log(0.01) / np.log(1 - (maxc / pnum) ** 4))
    best = int(np.argmax(cnums[:L])) if L else -1
    out.update(counts=cnums, best_sample=best, op_num=int(cnums[best]) if L else 0, thr=0.001 * dist, dist=dist,
               n_loops=L, n_iteration_ransac=int(min(rst, nit)))
    if out["op_num"] < 3:
        out["status"] = 4
        return out
    s = np.asarray(samples[best])
    rot, trans, _ = find_transform_matrix(Ya[s], Yb[s])
    mask = ((Yb @ rot.T + trans - Ya) ** 2).sum(1) < 0.001 * dist
    R, T, sta = find_transform_matrix(Ya[mask], Yb[mask], threshold=1e-14)
    en = np.sqrt(((Yb[mask] @ R.T + T - Ya[mask]) ** 2).sum(1))
    out.update(mask=mask, R=R, T=T, state=sta, R_hyp=rot, T_hyp=trans, error_mean=en.mean(),
               error_std=en.std(ddof=1) if en.size > 1 else 0.0)
    return out


def fspecial_gaussian3(sigma):
    """fspecial('gaussian', [3 3], sigma) (Image Processing Toolbox)."""
    xx, yy = np.meshgrid(np.arange(-1, 2), np.arange(-1, 2))
    h = np.exp(-(xx * xx + yy * yy) / (2.0 * sigma * sigma))
    h[h < np.finfo(float).eps * h.max()] = 0
    return h / h.sum()


def read_xyz_sr4000(sr_data, sigma=2.0, boundary="same"):
    """M/read_xyz_sr4000.m:8-26 (sigma 2, imfilter 'same' = zero padding) or
    M/code_from_dr_ye/read_sr4000_data_dr_ye.m:8,27-36,88-90 (sigma 1, 'replicate') on the rows x 176 matrix
    `load` returned.  scipy.ndimage.correlate stands for imfilter.  Returns x, y, z, confidence_map (or None)."""
    from scipy import ndimage
    sr = np.asarray(sr_data, float)
    G = fspecial_gaussian3(sigma)
    mode = {"same": "constant", "replicate": "nearest"}[boundary]
    z = ndimage.correlate(sr[0:144, :176], G, mode=mode, cval=0.0)
    x = ndimage.correlate(sr[144:288, :176], G, mode=mode, cval=0.0)
    y = ndimage.correlate(sr[288:432, :176], G, mode=mode, cval=0.0)
    cm = sr[576:720, :176] if sr.shape[0] >= 720 else None
    return x, y, z, cm


def sift_extract_xyz(sr_data, frames):
    """The per-feature loop of M/SIFT_extract_save.m:55-56,75-88 with M/inittialize_depth_my_version.m:16,31-85.
    frames: 4 x K as sift returns them (0-based x, y).  Returns (xyz_data 3 x K with NaN columns, idxRemain 0-based)."""
    x, y, z, cm = read_xyz_sr4000(sr_data, 2.0, "same")
    fr = np.array(frames, float)
    fr[:2] += 1
    mc = np.nanmax(cm) if cm is not None else None
    out = np.full((3, fr.shape[1]), np.nan)
    remain = []
    for i in range(fr.shape[1]):
        r = int(np.floor(fr[1, i] + 0.5)) - 1
        c = int(np.floor(fr[0, i] + 0.5)) - 1
        xf, yf, zf = x[r, c], y[r, c], z[r, c]
        if np.isnan(xf):
            continue
        df = np.sqrt(xf ** 2 + yf ** 2 + zf ** 2)
        if df < 0.4 or (cm is not None and cm[r, c] <= 0.5 * mc):
            continue
        out[:, i] = [-xf, -yf, zf]
        remain.append(i)
    return out, np.array(remain, int)


def rot_angle(Ra, Rb):
    """Geodesic distance between two rotations, radians."""
    D = np.asarray(Ra) @ np.asarray(Rb).T
    c = np.clip((np.trace(D) - 1) / 2, -1, 1)
    # for tiny angles use the skew part (acos loses precision near 1)
    sk = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    s = np.linalg.norm(sk)
    return float(np.arctan2(s, c))

"""Independent numpy / LAPACK restatement of the 1-point-RANSAC EKF hypothesis path ("oracle A").

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/pre3_oracle_ekf.c: dense H rows,
numpy sin / cos, numpy.linalg.inv and BLAS products the way the MATLAB reference uses its
built-ins, MATLAB-shaped arrays (column vectors, n x 4 logical pattern).  Checks the C oracle
and the CUDA kernels to tolerance (states to 1e-9) and the supports / masks wherever no residual
lies within 1e-9 px of its threshold.  `M/` = /root/reference/matlab_code/.
"""
from __future__ import annotations

import numpy as np


def m(a):
    """M/m.m:32-34."""
    theta, phi = a[0], a[1]
    cphi = np.cos(phi)
    return np.vstack([cphi * np.sin(theta), -np.sin(phi), cphi * np.cos(theta)])


def q2r(q):
    """M/q2r.m:29-36."""
    r, x, y, z = q
    return np.array([[r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)],
                     [2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)],
                     [2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z]])


def distort_fm(uv, cam):
    """M/distort_fm_my_version.m:44-61."""
    xu = (uv[0] - cam["Cx"]) / cam["f"]
    yu = (uv[1] - cam["Cy"]) / cam["f"]
    ru = np.sqrt(xu * xu + yu * yu)
    D = 1 + cam["k1"] * ru ** 2 + cam["k2"] * ru ** 4
    return np.vstack([xu * D * cam["f"] + cam["Cx"], yu * D * cam["f"] + cam["Cy"]])


def generate_state_vector_pattern(types, has_z, z, n):
    """M/generate_state_vector_pattern.m:27-53.  types: 0 inverse depth / 1 cartesian; z: (F,2)."""
    pattern = np.zeros((n, 4))
    position = 13  # 0-based 14
    z_id, z_euc = [], []
    for i, t in enumerate(types):
        if t == 0:
            if has_z[i]:
                pattern[position:position + 3, 0] = 1
                pattern[position + 3:position + 5, 1] = 1
                pattern[position + 5, 2] = 1
                z_id.append(z[i])
            position += 6
        else:
            if has_z[i]:
                pattern[position:position + 3, 3] = 1
                z_euc.append(z[i])
            position += 3
    z_id = np.array(z_id, float).reshape(-1, 2).T
    z_euc = np.array(z_euc, float).reshape(-1, 2).T
    return pattern, z_id, z_euc


def compute_hypothesis_support_fast(xi, cam, pattern, z_id, z_euc, threshold, return_residuals=False):
    """M/compute_hypothesis_support_fast.m:27-116.  z_id: 2 x n_id, z_euc: 2 x n_euc."""
    xi = np.asarray(xi, float).ravel()
    support = 0
    li_id = np.zeros(0, bool)
    li_euc = np.zeros(0, bool)
    res_all = []
    rotcw = q2r(xi[3:7]).T
    if z_id.size:
        n_id = z_id.shape[1]
        ri = xi[pattern[:, 0].astype(bool)].reshape(n_id, 3).T
        ang = xi[pattern[:, 1].astype(bool)].reshape(n_id, 2).T
        rho = xi[pattern[:, 2].astype(bool)]
        mi = m(ang)
        d = (ri - xi[0:3, None]) * rho[None, :]
        hc = rotcw @ (d + mi)
        hn = hc[0:2] / hc[2]
        h_image = cam["f"] * hn + np.array([[cam["Cx"]], [cam["Cy"]]])
        nu = z_id - distort_fm(h_image, cam)
        residuals = np.sqrt(nu[0] ** 2 + nu[1] ** 2)
        li_id = residuals < (np.nanmin(residuals) + threshold if np.isfinite(residuals).any() else np.nan)
        support += int(li_id.sum())
        res_all.append(residuals)
    if z_euc.size:
        n_euc = z_euc.shape[1]
        xyz = xi[pattern[:, 3].astype(bool)].reshape(n_euc, 3).T
        hc = rotcw @ (xyz - xi[0:3, None])
        hn = hc[0:2] / hc[2]
        h_image = cam["f"] * hn + np.array([[cam["Cx"]], [cam["Cy"]]])
        nu = z_euc - distort_fm(h_image, cam)
        residuals = np.sqrt(nu[0] ** 2 + nu[1] ** 2)
        li_euc = residuals < threshold
        support += int(li_euc.sum())
        res_all.append(residuals)
    if return_residuals:
        return support, li_id, li_euc, (np.concatenate(res_all) if res_all else np.zeros(0))
    return support, li_id, li_euc


def dense_H(fr, i):
    """features_info(i).H as the full 2 x n matrix (M/calculate_Hi_inverse_depth_my_version.m:44-49)."""
    H = np.zeros((2, fr.n))
    H[:, 0:13] = fr.Hcam[i].T      # stored (13, 2) = 2 x 13 column-major
    nf = 6 if fr.type[i] == 0 else 3
    H[:, fr.pos[i]:fr.pos[i] + nf] = fr.Hfeat[i].T[:, :nf]
    return H


def hypothesis_state(fr, sel):
    """M/ransac_hypotheses.m:51-63 for the matches `sel` (0-based feature indices)."""
    hi = np.concatenate([fr.h[i] for i in sel])
    zi = np.concatenate([fr.z[i] for i in sel])
    Hi = np.vstack([dense_H(fr, i) for i in sel])
    R = np.zeros((2 * len(sel), 2 * len(sel)))
    for a, i in enumerate(sel):
        R[2 * a:2 * a + 2, 2 * a:2 * a + 2] = fr.R[i].T  # stored column-major
    S = Hi @ fr.P @ Hi.T + R
    K = fr.P @ Hi.T @ np.linalg.inv(S)
    return fr.x + K @ (zi - hi)


def n_hyp_rule(support, num_ic):
    """M/ransac_hypotheses.m:77-78 with MATLAB's complex log for negative arguments."""
    epsilon = 1 - support / num_ic
    a = 1 - (1 - epsilon)
    with np.errstate(divide="ignore"):
        v = np.log(complex(1 - 0.99)) / np.log(complex(a)) if a != 0 else complex(0.0)
    return float(np.ceil(v.real))


def ransac_hypotheses(fr, sel, n_hyp_init=1000):
    """M/ransac_hypotheses.m:27-85.  sel: (H, 3) supplied selections (0-based feature indices)."""
    pattern, z_id, z_euc = generate_state_vector_pattern(fr.type, fr.has_z, fr.z, fr.n)
    num_ic = int(fr.ic.sum())
    if num_ic == 0:
        raise IndexError("select_random_match: no individually compatible match")
    mm = 3 if num_ic > 3 else 1
    n_hyp = float(n_hyp_init)
    max_support = 0
    li = fr.li0.copy()
    best = -1
    supports = []
    id_feats = [i for i in range(fr.F) if fr.has_z[i] and fr.type[i] == 0]
    euc_feats = [i for i in range(fr.F) if fr.has_z[i] and fr.type[i] == 1]
    for i in range(min(n_hyp_init, len(sel))):
        if n_hyp == 0:
            break
        xi = hypothesis_state(fr, list(sel[i][:mm]))
        sup, li_id, li_euc = compute_hypothesis_support_fast(xi, fr.cam, pattern, z_id, z_euc, fr.std_z)
        supports.append(sup)
        if sup > max_support:
            max_support = sup
            best = i
            li[id_feats] = li_id
            li[euc_feats] = li_euc
            n_hyp = n_hyp_rule(sup, num_ic)
        if n_hyp <= mm:
            break
    return dict(li=li, n_hyp=n_hyp, max_support=max_support, best_hyp=best, supports=np.array(supports),
                n_evaluated=len(supports), m=mm, num_ic=num_ic)


# ---------------------------------------------------------------------------------------------------
# SURVEY.md 8f rank 3 (first part): the EKF partial updates around ransac_hypotheses
# ---------------------------------------------------------------------------------------------------
def normJac(q):
    """M/normJac.m:1-16."""
    r, x, y, z = q
    return (r * r + x * x + y * y + z * z) ** (-1.5) * np.array(
        [[x * x + y * y + z * z, -r * x, -r * y, -r * z],
         [-x * r, r * r + y * y + z * z, -x * y, -x * z],
         [-y * r, -y * x, r * r + x * x + z * z, -y * z],
         [-z * r, -z * x, -z * y, r * r + x * x + y * y]])


def update(x_km1_k, p_km1_k, H, R, z, h):
    """[x_k_k, p_k_k, K] = update(x_km1_k, p_km1_k, H, R, z, h)   (M/update.m:27-56), dense numpy / LAPACK."""
    if len(z) == 0:
        return x_km1_k.copy(), p_km1_k.copy(), 0
    S = H @ p_km1_k @ H.T + R
    K = p_km1_k @ H.T @ np.linalg.inv(S)
    x = x_km1_k + K @ (z - h)
    p = p_km1_k - K @ S @ K.T
    p = 0.5 * p + 0.5 * p.T
    J = normJac(x[3:7])
    n = p.shape[0]
    p = np.block([[p[0:3, 0:3], p[0:3, 3:7] @ J.T, p[0:3, 7:n]],
                  [J @ p[3:7, 0:3], J @ p[3:7, 3:7] @ J.T, J @ p[3:7, 7:n]],
                  [p[7:n, 0:3], p[7:n, 3:7] @ J.T, p[7:n, 7:n]]])
    x = x.copy()
    x[3:7] = x[3:7] / np.linalg.norm(x[3:7])
    return x, p, K


def ekf_update_inliers(fr, flags, x=None, P=None):
    """M/@ekf_filter/ekf_update_li_inliers.m:15-29 (flags = low_innovation_inlier, x/P = x_k_km1/p_k_km1) and
    ekf_update_hi_inliers.m:18-32 (flags = high_innovation_inlier, x/P = x_k_k/p_k_k): stack z, h, H of the flagged
    features in feature order, R = eye, update()."""
    x = fr.x if x is None else x
    P = fr.P if P is None else P
    sel = [i for i in range(fr.F) if flags[i] == 1]
    if not sel:
        return update(x, P, np.zeros((0, fr.n)), np.zeros((0, 0)), np.zeros(0), np.zeros(0))[:2]
    z = np.concatenate([fr.z[i] for i in sel])
    h = np.concatenate([fr.h[i] for i in sel])
    H = np.vstack([dense_H(fr, i) for i in sel])
    xk, pk, _ = update(x, P, H, np.eye(len(z)), z, h)
    return xk, pk


def rescue_hi_inliers(fr, p_k_k, li, h=None):
    """The test of M/@ekf_filter/rescue_hi_inliers.m:35-46 for features that are individually compatible but not
    low-innovation inliers: nu' * inv(H p_k_k H') * nu < chi2inv(0.95, 2) = 5.9915.  (The re-prediction of h and H at
    x_k_k, :32-33, is the caller's; pass the re-predicted h, or None for fr.h.)  Returns the high_innovation_inlier
    flags (-1 where the reference leaves the field untouched)."""
    h = fr.h if h is None else h
    out = np.full(fr.F, -1, np.int32)
    for i in range(fr.F):
        if fr.ic[i] == 1 and li[i] == 0:
            Hi = dense_H(fr, i)
            Si = Hi @ p_k_k @ Hi.T
            nu = fr.z[i] - h[i]
            out[i] = 1 if nu @ np.linalg.inv(Si) @ nu < 5.9915 else 0
    return out


# ------------------------------------------------------------------------------------------
# re-prediction at x_k_k (rescue_hi_inliers.m:32-33): independent restatement with numpy / LAPACK
# ------------------------------------------------------------------------------------------
def _h_of(x, cam, ty, ps):
    """hi_inverse_depth.m:27-86 / hi_cartesian.m:27-80 without the visibility tests: the distorted pixel."""
    Rcw = q2r(x[3:7]).T
    y = x[ps:ps + (6 if ty == 0 else 3)]
    if ty == 0:
        v = (y[:3] - x[:3]) * y[5] + m(y[3:5])[:, 0]
    else:
        v = y - x[:3]
    hrl = Rcw @ v
    uv = np.array([cam["Cx"] + hrl[0] / hrl[2] * cam["f"], cam["Cy"] + hrl[1] / hrl[2] * cam["f"]])
    return distort_fm(uv, cam)[:, 0], hrl


def predict_and_derivatives(x, cam, n_rows, n_cols, types, pos, has_h, h_in):
    """predict_camera_measurements.m:27-68 + calculate_derivatives.m:27-59.  Returns h (F,2), has_h (F,), predicted (F,),
    H (F,2,n) dense (zero rows where h is empty)."""
    x = np.asarray(x, np.float64)
    F, n = len(types), len(x)
    h = np.array(h_in, np.float64).copy()
    has = np.array(has_h, bool).copy()
    pred = np.zeros(F, bool)
    H = np.zeros((F, 2, n))
    Rrw = np.linalg.inv(q2r(x[3:7]))
    f, Cx, Cy, k1, k2 = cam["f"], cam["Cx"], cam["Cy"], cam["k1"], cam["k2"]
    for i in range(F):
        ty, ps = int(types[i]), int(pos[i])
        hd, hrl = _h_of(x, cam, ty, ps)
        ax, ay = np.degrees(np.arctan2(hrl[0], hrl[2])), np.degrees(np.arctan2(hrl[1], hrl[2]))
        ok = (-60 <= ax <= 60) and (-60 <= ay <= 60) and (0 < hd[0] < n_cols) and (0 < hd[1] < n_rows)
        if ok:
            h[i], has[i], pred[i] = hd, True, True
        if not has[i]:
            continue
        y = x[ps:ps + (6 if ty == 0 else 3)]
        rho = y[5] if ty == 0 else 1.0
        a = ((y[:3] - x[:3]) * y[5] + m(y[3:5])[:, 0]) if ty == 0 else (y - x[:3])
        hc = Rrw @ a
        dhu = np.array([[f / hc[2], 0, -hc[0] * f / hc[2] ** 2], [0, f / hc[2], -hc[1] * f / hc[2] ** 2]])
        u, v = h[i]
        r2 = ((u - Cx) ** 2 + (v - Cy) ** 2) / f ** 2
        g = k1 + 2 * k2 * r2
        D = 1 + k1 * r2 + k2 * r2 * r2
        J = np.array([[D + (u - Cx) * g * 2 * (u - Cx) / f ** 2, (u - Cx) * g * 2 * (v - Cy) / f ** 2],
                      [(v - Cy) * g * 2 * (u - Cx) / f ** 2, D + (v - Cy) * g * 2 * (v - Cy) / f ** 2]])
        dh = np.linalg.inv(np.linalg.inv(J)) @ dhu
        q0, qx, qy, qz = x[3], -x[4], -x[5], -x[6]
        dR = [np.array([[2 * q0, -2 * qz, 2 * qy], [2 * qz, 2 * q0, -2 * qx], [-2 * qy, 2 * qx, 2 * q0]]),
              np.array([[2 * qx, 2 * qy, 2 * qz], [2 * qy, -2 * qx, -2 * q0], [2 * qz, 2 * q0, -2 * qx]]),
              np.array([[-2 * qy, 2 * qx, 2 * q0], [2 * qx, 2 * qy, 2 * qz], [-2 * q0, 2 * qz, -2 * qy]]),
              np.array([[-2 * qz, -2 * q0, 2 * qx], [2 * q0, -2 * qz, 2 * qy], [2 * qx, 2 * qy, 2 * qz]])]
        dq = np.stack([d @ a for d in dR], 1) @ np.diag([1.0, -1, -1, -1])
        H[i, :, 0:3] = dh @ (-Rrw * rho)
        H[i, :, 3:7] = dh @ dq
        if ty == 0:
            th, ph = y[3], y[4]
            dy = np.column_stack([rho * Rrw, Rrw @ np.array([np.cos(ph) * np.cos(th), 0, -np.cos(ph) * np.sin(th)]),
                                  Rrw @ np.array([-np.sin(ph) * np.sin(th), -np.cos(ph), -np.sin(ph) * np.cos(th)]),
                                  Rrw @ (y[:3] - x[:3])])
            H[i, :, ps:ps + 6] = dh @ dy
        else:
            H[i, :, ps:ps + 3] = dh @ Rrw
    return h, has, pred, H


def numeric_H(x, cam, ty, ps, eps=1e-6):
    """Central-difference Jacobian of the predicted (distorted) pixel with respect to the state."""
    x = np.asarray(x, np.float64)
    H = np.zeros((2, len(x)))
    for k in list(range(7)) + list(range(ps, ps + (6 if ty == 0 else 3))):
        hi, lo = x.copy(), x.copy()
        hi[k] += eps
        lo[k] -= eps
        H[:, k] = (_h_of(hi, cam, ty, ps)[0] - _h_of(lo, cam, ty, ps)[0]) / (2 * eps)
    return H

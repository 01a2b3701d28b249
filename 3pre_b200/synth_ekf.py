"""Seeded synthetic EKF frames for BASELINE.json config 4: the inputs
ransac_hypotheses (M/ransac_hypotheses.m:27) sees at one time step of the 1-point-RANSAC EKF.

Per frame: a 13-state camera (r 3, q 4, v 3, w 3) + n_id inverse-depth features (6 states:
anchor 3, azimuth, elevation, inverse depth) + n_euc cartesian features (3 states), so
n = 13 + 6 n_id + 3 n_euc (200 inverse-depth features -> n = 1213).  The prior is the truth plus
a low-rank pose error (1 cm, 0.5 deg) coupled into the feature states, P = A A' + D is the
matching covariance; measurements are the distorted projections of the truth + N(0, 0.25 px),
a fraction of them gross outliers (+-10..40 px); h is the prediction at the prior and H its
Jacobian (central differences; H is an INPUT of the path, the reference computes it in
calculate_derivatives.m, outside the path).  Camera: M/initialize_cam.m:52-76.

Written with torch so that the same code makes a handful of frames on the CPU for the parity
tests and thousands on the device for bench.py.  All matrices are stored in MATLAB's
column-major order (a "2 x 13" block is a (13, 2) C-contiguous array).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .synth import CX, CY, F_PX, K1_DIST, K2_DIST

CAM = dict(f=F_PX, Cx=CX, Cy=CY, k1=K1_DIST, k2=K2_DIST)


def _q2r(q):
    import torch
    r, x, y, z = q.unbind(-1)
    return torch.stack([
        torch.stack([r * r + x * x - y * y - z * z, 2 * (x * y - r * z), 2 * (z * x + r * y)], -1),
        torch.stack([2 * (x * y + r * z), r * r - x * x + y * y - z * z, 2 * (y * z - r * x)], -1),
        torch.stack([2 * (z * x - r * y), 2 * (y * z + r * x), r * r - x * x - y * y + z * z], -1)], -2)


def _project(xv, y, typ):
    """Distorted pixel of every feature.  xv (Fr,13), y (Fr,F,6), typ (Fr,F) -> (Fr,F,2)."""
    import torch
    r = xv[:, None, 0:3]
    Rwc = _q2r(xv[:, 3:7])  # (Fr,3,3)
    theta, phi, rho = y[..., 3], y[..., 4], y[..., 5]
    cphi = torch.cos(phi)
    mi = torch.stack([cphi * torch.sin(theta), -torch.sin(phi), cphi * torch.cos(theta)], -1)
    v_id = (y[..., 0:3] - r) * rho[..., None] + mi
    v_euc = y[..., 0:3] - r
    v = torch.where((typ == 0)[..., None], v_id, v_euc)
    hc = torch.einsum("fji,fnj->fni", Rwc, v)  # rotcw * v
    u = F_PX * hc[..., 0] / hc[..., 2] + CX
    w = F_PX * hc[..., 1] / hc[..., 2] + CY
    xu, yu = (u - CX) / F_PX, (w - CY) / F_PX
    ru2 = xu * xu + yu * yu
    D = 1 + K1_DIST * ru2 + K2_DIST * ru2 * ru2
    return torch.stack([xu * D * F_PX + CX, yu * D * F_PX + CY], -1)


def make_ekf_frames(Fr, seed, device="cpu", n_id=200, n_euc=0, outlier_ratio=0.20, meas_noise=0.25, std_z=1.0,
                    drop_z=0.0, drop_ic=0.0, interleave=False, asym=0.0):
    """Fr frames as a dict of torch tensors (float64 unless noted):
      x (Fr,n)  P (Fr,n,n) [P[f,c,r] = p(r,c): column-major]  type/pos (Fr,F) int32  has_z/ic/li0 (Fr,F) uint8
      z,h (Fr,F,2)  Hcam (Fr,F,13,2)  Hfeat (Fr,F,6,2)  R (Fr,F,2,2)  x_true (Fr,n)  outlier (Fr,F) bool
    plus python scalars n, F, std_z."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = torch.float64
    F = n_id + n_euc

    def randn(*s):
        return torch.randn(*s, device=device, generator=g, dtype=f64)

    def rand(*s):
        return torch.rand(*s, device=device, generator=g, dtype=f64)

    typ = torch.cat([torch.zeros(n_id, dtype=torch.int32), torch.ones(n_euc, dtype=torch.int32)]).to(device)
    if interleave and n_euc > 0:
        perm = torch.randperm(F, generator=torch.Generator().manual_seed(int(seed) + 1)).to(device)
        typ = typ[perm]
    typ = typ[None, :].expand(Fr, F).contiguous()
    nf = torch.where(typ == 0, 6, 3)
    pos = (13 + torch.cumsum(nf, 1) - nf).to(torch.int32)
    n = 13 + 6 * n_id + 3 * n_euc

    # ---- truth ---------------------------------------------------------------------------------
    xv = torch.zeros(Fr, 13, device=device, dtype=f64)
    xv[:, 0:3] = 0.05 * randn(Fr, 3)
    axis = randn(Fr, 3)
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = rand(Fr) * np.deg2rad(5.0)
    xv[:, 3] = torch.cos(ang / 2)
    xv[:, 4:7] = torch.sin(ang / 2)[:, None] * axis
    xv[:, 7:10] = 0.1 * randn(Fr, 3)
    xv[:, 10:13] = 0.05 * randn(Fr, 3)
    u = 8 + rand(Fr, F) * 160
    w = 8 + rand(Fr, F) * 128
    depth = 0.8 + 4.2 * rand(Fr, F)
    pc = torch.stack([(u - CX) / F_PX * depth, (w - CY) / F_PX * depth, depth], -1)
    Rwc = _q2r(xv[:, 3:7])
    pw = torch.einsum("fij,fnj->fni", Rwc, pc) + xv[:, None, 0:3]
    anchor = 0.1 * randn(Fr, F, 3)
    d = pw - anchor
    rho = 1.0 / d.norm(dim=-1)
    mdir = d * rho[..., None]
    theta = torch.atan2(mdir[..., 0], mdir[..., 2])
    phi = torch.atan2(-mdir[..., 1], torch.sqrt(mdir[..., 0] ** 2 + mdir[..., 2] ** 2))
    y_id = torch.cat([anchor, theta[..., None], phi[..., None], rho[..., None]], -1)
    y_euc = torch.cat([pw, torch.zeros(Fr, F, 3, device=device, dtype=f64)], -1)
    y_true = torch.where((typ == 0)[..., None], y_id, y_euc)

    fi = torch.arange(Fr, device=device)[:, None].expand(Fr, F)

    def pack(xv_, y_):
        x_ = torch.zeros(Fr, n, device=device, dtype=f64)
        x_[:, :13] = xv_
        for k in range(6):
            ok = k < nf
            x_[fi[ok], (pos.long() + k)[ok]] = y_[..., k][ok]
        return x_

    # ---- prior: truth + A w + sqrt(D) e,  P = A A' + D -------------------------------------------
    A = torch.zeros(Fr, n, 6, device=device, dtype=f64)
    A[:, 0, 0] = A[:, 1, 1] = A[:, 2, 2] = 0.01
    qr, qx, qy, qz = xv[:, 3], xv[:, 4], xv[:, 5], xv[:, 6]
    Q = 0.5 * torch.stack([torch.stack([-qx, -qy, -qz], -1), torch.stack([qr, -qz, qy], -1),
                           torch.stack([qz, qr, -qx], -1), torch.stack([-qy, qx, qr], -1)], -2)
    A[:, 3:7, 3:6] = Q * np.deg2rad(0.5)
    couple = torch.tensor([2e-3, 2e-3, 2e-3, 5e-4, 5e-4, 1e-3], device=device, dtype=f64)
    D = torch.full((Fr, n), 1e-8, device=device, dtype=f64)
    D[:, 7:13] = 1e-4
    for k in range(6):
        ok = k < nf
        rows = couple[k] * randn(Fr, F, 6)
        A[fi[ok], (pos.long() + k)[ok], :] = rows[ok]
    wv = randn(Fr, 6)
    x_true = pack(xv, y_true)
    x = x_true + torch.einsum("fnk,fk->fn", A, wv) + torch.sqrt(D) * randn(Fr, n) * 0.5
    P = A @ A.transpose(1, 2) + torch.diag_embed(D)
    if asym:
        P = P + asym * torch.triu(randn(Fr, n, n), 1) * P.diagonal(dim1=1, dim2=2).mean(1)[:, None, None]

    def unpack(x_):
        y_ = torch.zeros(Fr, F, 6, device=device, dtype=f64)
        for k in range(6):
            ok = k < nf
            y_[..., k][ok] = x_[fi[ok], (pos.long() + k)[ok]]
        return x_[:, :13], y_

    # ---- measurements, prediction, Jacobians ---------------------------------------------------
    z = _project(xv, y_true, typ) + meas_noise * randn(Fr, F, 2)
    outlier = rand(Fr, F) < outlier_ratio
    off = (10 + 30 * rand(Fr, F, 2)) * torch.where(rand(Fr, F, 2) < 0.5, -1.0, 1.0)
    z = torch.where(outlier[..., None], z + off, z)
    xv0, y0 = unpack(x)
    h = _project(xv0, y0, typ)
    eps = 1e-6
    Hcam = torch.zeros(Fr, F, 13, 2, device=device, dtype=f64)
    for k in range(7):  # velocities do not enter the measurement
        dp = torch.zeros(13, device=device, dtype=f64)
        dp[k] = eps
        Hcam[:, :, k, :] = (_project(xv0 + dp, y0, typ) - _project(xv0 - dp, y0, typ)) / (2 * eps)
    Hfeat = torch.zeros(Fr, F, 6, 2, device=device, dtype=f64)
    for k in range(6):
        dp = torch.zeros(6, device=device, dtype=f64)
        dp[k] = eps
        J = (_project(xv0, y0 + dp, typ) - _project(xv0, y0 - dp, typ)) / (2 * eps)
        Hfeat[:, :, k, :] = J * (k < nf)[..., None]
    R = torch.zeros(Fr, F, 2, 2, device=device, dtype=f64)
    R[..., 0, 0] = R[..., 1, 1] = std_z * std_z
    has_z = (rand(Fr, F) >= drop_z).to(torch.uint8)
    ic = (has_z.bool() & (rand(Fr, F) >= drop_ic)).to(torch.uint8)
    return dict(x=x.contiguous(), P=P.transpose(1, 2).contiguous(), type=typ, pos=pos.contiguous(), has_z=has_z, ic=ic,
                li0=torch.zeros(Fr, F, dtype=torch.uint8, device=device), z=z.contiguous(), h=h.contiguous(),
                Hcam=Hcam.contiguous(), Hfeat=Hfeat.contiguous(), R=R.contiguous(), x_true=x_true, outlier=outlier,
                n=n, F=F, std_z=float(std_z))


@dataclass
class EkfFrame:
    """One frame as numpy arrays (what the oracle and the MATLAB-shaped mirror take)."""
    n: int
    F: int
    x: np.ndarray      # (n,)
    P: np.ndarray      # (n,n) mathematical matrix: P[r,c] = p(r,c)
    type: np.ndarray   # (F,) int32
    pos: np.ndarray    # (F,) int32, 0-based
    has_z: np.ndarray  # (F,) uint8
    ic: np.ndarray     # (F,) uint8
    li0: np.ndarray    # (F,) uint8
    z: np.ndarray      # (F,2)
    h: np.ndarray      # (F,2)
    Hcam: np.ndarray   # (F,13,2)
    Hfeat: np.ndarray  # (F,6,2)
    R: np.ndarray      # (F,2,2) column-major blocks
    cam: dict
    std_z: float
    outlier: np.ndarray


def frame(batch, i) -> EkfFrame:
    t = lambda k: batch[k][i].detach().cpu().numpy()
    return EkfFrame(n=batch["n"], F=batch["F"], x=t("x"), P=t("P").T.copy(), type=t("type"), pos=t("pos"),
                    has_z=t("has_z"), ic=t("ic"), li0=t("li0"), z=t("z"), h=t("h"), Hcam=t("Hcam"), Hfeat=t("Hfeat"),
                    R=t("R"), cam=dict(CAM), std_z=batch["std_z"], outlier=t("outlier"))


def make_selections(fr_ic, H, seed):
    """H supplied match selections for one frame: rows of 3 distinct IC feature indices in random
    order (what select_random_match.m:40-58 draws), int32 (H,3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ic_list = np.flatnonzero(fr_ic)
    out = np.zeros((H, 3), np.int32)
    if len(ic_list) == 0:
        return out
    for i in range(H):
        k = min(3, len(ic_list))
        out[i, :k] = rng.permutation(ic_list)[:k]
    return out


def batch_to_numpy(batch):
    """torch batch -> the dict of numpy arrays Context.ransac_hypotheses_batch takes."""
    out = {k: (v.detach().cpu().numpy() if hasattr(v, "detach") else v) for k, v in batch.items()}
    out["cam"] = dict(CAM)
    return out


def to_features_info(fr: EkfFrame):
    """MATLAB-shaped inputs of ransac_hypotheses for one frame: (filter, features_info, cam)."""
    feats = []
    for i in range(fr.F):
        nf = 6 if fr.type[i] == 0 else 3
        H = np.zeros((2, fr.n))
        H[:, :13] = fr.Hcam[i].T
        H[:, fr.pos[i]:fr.pos[i] + nf] = fr.Hfeat[i].T[:, :nf]
        feats.append(dict(type="inversedepth" if fr.type[i] == 0 else "cartesian",
                          z=fr.z[i].reshape(2, 1) if fr.has_z[i] else np.zeros((0, 0)), h=fr.h[i].reshape(1, 2), H=H,
                          R=fr.R[i].T.copy(), individually_compatible=int(fr.ic[i]),
                          low_innovation_inlier=int(fr.li0[i])))
    filt = dict(x_k_km1=fr.x.reshape(-1, 1), p_k_km1=fr.P, std_z=fr.std_z)
    return filt, feats, dict(fr.cam)

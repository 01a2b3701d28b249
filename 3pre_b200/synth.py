"""Seeded synthetic SR4000 frame pairs.

The reference ships no dataset (its d1_%04d.dat logs are external), so every
config of BASELINE.json is driven by this generator:

* frame geometry 176x144, pinhole f=250.57731, Cx=91.69, Cy=72.27
  (M/initialize_cam.m:52-53,73-76); scene depth z ~ U[0.8, 5.0] m, points stored in
  the reference's "code coordinates" [-x, -y, z] (M/inittialize_depth_my_version.m:85);
* known rigid motion: rotation axis uniform on S^2, angle ~ U[0, 5 deg], |t| ~ U[0, 0.10] m;
  inliers Ya = R*Yb + t + N(0, (2 mm)^2); outliers (fraction rho) get an independent point;
* descriptors: |N(0,1)|^128, L2-normalised, clamped at 0.2, renormalised, rounded to
  float32 and stored as double (M/sift/siftdescriptor.c:500-527); a planted
  correspondence re-uses the raw vector plus N(0, 0.02^2) noise before normalisation.

numpy path = parity tests (deterministic, PCG64(seed)); torch path = large batches
generated on the device for bench.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

W, H_IMG = 176, 144
F_PX, CX, CY = 250.57731, 91.69, 72.27
K1_DIST, K2_DIST = -0.84656, 0.53701  # M/initialize_cam.m:64-65


def _normalise_desc(raw):
    d = raw / np.linalg.norm(raw, axis=-1, keepdims=True)
    d = np.minimum(d, 0.2)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return d.astype(np.float32).astype(np.float64)


def _scene_points(rng, n):
    u = rng.uniform(0, W - 1, n)
    v = rng.uniform(0, H_IMG - 1, n)
    z = rng.uniform(0.8, 5.0, n)
    x = (u - CX) * z / F_PX
    y = (v - CY) * z / F_PX
    return np.stack([-x, -y, z], axis=1)


def random_motion(rng, max_angle_deg=5.0, max_t=0.10):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0, max_angle_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    t = rng.normal(size=3)
    t *= rng.uniform(0, max_t) / np.linalg.norm(t)
    return R, t


@dataclass
class Correspondences:
    Ya: np.ndarray  # (N,3) previous frame
    Yb: np.ndarray  # (N,3) current frame
    R: np.ndarray
    t: np.ndarray
    inlier: np.ndarray  # planted inlier flags


def make_correspondences(seed, N=300, outlier_ratio=0.30, noise=0.002):
    """Already-matched 3-D correspondences Ya ~ R*Yb + t (RANSAC_CALC_VER2 inputs)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, t = random_motion(rng)
    Yb = _scene_points(rng, N)
    Ya = Yb @ R.T + t + rng.normal(scale=noise, size=(N, 3))
    n_out = int(round(outlier_ratio * N))
    out_idx = rng.permutation(N)[:n_out]
    Ya[out_idx] = _scene_points(rng, n_out)
    inl = np.ones(N, bool)
    inl[out_idx] = False
    return Correspondences(Ya, Yb, R, t, inl)


def make_samples(seed, H, N, k):
    """Sample-index sets (H,k) int32, 0-based, each row k distinct ascending indices --
    what get_rand(k, N) yields (M/Common/get_rand.m:43-48) -- supplied to oracle and GPU alike."""
    rng = np.random.Generator(np.random.PCG64(seed))
    keys = rng.random((H, N))
    idx = np.argpartition(keys, k - 1, axis=1)[:, :k]
    return np.sort(idx, axis=1).astype(np.int32)


def make_draws(seed, H, N, k=4):
    """Draws (H,k) int32, 0-based, k distinct indices per row in DRAW order (not sorted) -- what
    num_rs(1..4) of M/code_from_dr_ye/ransac_dr_ye.m:28-48 holds once its re-draw loops have ended."""
    rng = np.random.Generator(np.random.PCG64(seed))
    keys = rng.random((H, N))
    return np.argsort(keys, axis=1)[:, :k].astype(np.int32)


@dataclass
class FramePair:
    desc1: np.ndarray  # (K1,128) float64 (float32-valued)  previous frame
    desc2: np.ndarray  # (K2,128)                           current frame
    xyz1: np.ndarray   # (K1,3)
    xyz2: np.ndarray   # (K2,3)
    R: np.ndarray
    t: np.ndarray
    corr: np.ndarray   # (n_corr,2) planted (k1,k2)
    inlier: np.ndarray  # (n_corr,) planted geometric inliers


def make_frame_pair(seed, K1=512, K2=512, n_corr=300, outlier_ratio=0.30, noise=0.002, desc_noise=0.02):
    """One SR4000-shaped frame pair in the SCAN_SIFT layout (M/SIFT_match_save.m:8-14)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, t = random_motion(rng)
    raw1 = np.abs(rng.normal(size=(K1, 128)))
    raw2 = np.abs(rng.normal(size=(K2, 128)))
    xyz2 = _scene_points(rng, K2)
    xyz1 = _scene_points(rng, K1)
    i1 = rng.permutation(K1)[:n_corr]
    i2 = rng.permutation(K2)[:n_corr]
    raw2[i2] = np.abs(raw1[i1] + rng.normal(scale=desc_noise, size=(n_corr, 128)))
    xyz1[i1] = xyz2[i2] @ R.T + t + rng.normal(scale=noise, size=(n_corr, 3))
    n_out = int(round(outlier_ratio * n_corr))
    inl = np.ones(n_corr, bool)
    inl[rng.permutation(n_corr)[:n_out]] = False
    xyz1[i1[~inl]] = _scene_points(rng, n_out)
    return FramePair(_normalise_desc(raw1), _normalise_desc(raw2), xyz1, xyz2, R, t,
                     np.stack([i1, i2], 1).astype(np.int32), inl)


def to_uint8(desc):
    """uint8(512*descr) as in M/sift/sift_demo2.m:93-94."""
    return np.clip(np.floor(512.0 * desc + 0.5), 0, 255).astype(np.uint8)


def make_batch_torch(P, seed, device, K1=512, K2=512, n_corr=300, outlier_ratio=0.30, noise=0.002,
                     desc_noise=0.02, dtype="float64"):
    """P frame pairs generated on `device` with torch (same distributions as
    make_frame_pair, different RNG stream).  Returns dict of tensors:
    desc1 (P,K1,128), desc2 (P,K2,128), xyz1 (P,K1,3), xyz2 (P,K2,3), R (P,3,3), t (P,3)."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = torch.float64

    def scene(n):
        u = torch.rand(P, n, device=device, generator=g, dtype=f64) * (W - 1)
        v = torch.rand(P, n, device=device, generator=g, dtype=f64) * (H_IMG - 1)
        z = torch.rand(P, n, device=device, generator=g, dtype=f64) * 4.2 + 0.8
        return torch.stack([-(u - CX) * z / F_PX, -(v - CY) * z / F_PX, z], dim=2)

    def norm_desc(raw):
        d = raw / raw.norm(dim=-1, keepdim=True)
        d = d.clamp(max=0.2)
        d = d / d.norm(dim=-1, keepdim=True)
        return d.to(torch.float32)

    axis = torch.randn(P, 3, device=device, generator=g, dtype=f64)
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = torch.rand(P, device=device, generator=g, dtype=f64) * np.deg2rad(5.0)
    Kx = torch.zeros(P, 3, 3, device=device, dtype=f64)
    Kx[:, 0, 1], Kx[:, 0, 2] = -axis[:, 2], axis[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = axis[:, 2], -axis[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -axis[:, 1], axis[:, 0]
    eye = torch.eye(3, device=device, dtype=f64).expand(P, 3, 3)
    R = eye + torch.sin(ang)[:, None, None] * Kx + (1 - torch.cos(ang))[:, None, None] * (Kx @ Kx)
    t = torch.randn(P, 3, device=device, generator=g, dtype=f64)
    t = t * (torch.rand(P, 1, device=device, generator=g, dtype=f64) * 0.10) / t.norm(dim=1, keepdim=True)

    raw1 = torch.randn(P, K1, 128, device=device, generator=g, dtype=torch.float32).abs()
    raw2 = torch.randn(P, K2, 128, device=device, generator=g, dtype=torch.float32).abs()
    xyz1 = scene(K1)
    xyz2 = scene(K2)
    i1 = torch.rand(P, K1, device=device, generator=g).argsort(dim=1)[:, :n_corr]
    i2 = torch.rand(P, K2, device=device, generator=g).argsort(dim=1)[:, :n_corr]
    src = torch.gather(raw1, 1, i1[:, :, None].expand(P, n_corr, 128))
    pert = (src + desc_noise * torch.randn(P, n_corr, 128, device=device, generator=g, dtype=torch.float32)).abs()
    raw2.scatter_(1, i2[:, :, None].expand(P, n_corr, 128), pert)
    yb = torch.gather(xyz2, 1, i2[:, :, None].expand(P, n_corr, 3))
    ya = yb @ R.transpose(1, 2) + t[:, None, :] + noise * torch.randn(P, n_corr, 3, device=device, generator=g, dtype=f64)
    n_out = int(round(outlier_ratio * n_corr))
    if n_out > 0:  # the first n_out planted correspondences (random features anyway) become outliers
        ya[:, :n_out] = scene(n_out)
    xyz1.scatter_(1, i1[:, :, None].expand(P, n_corr, 3), ya)
    d1 = norm_desc(raw1)
    d2 = norm_desc(raw2)
    if dtype == "float64":
        d1, d2 = d1.to(f64), d2.to(f64)
    return dict(desc1=d1.contiguous(), desc2=d2.contiguous(), xyz1=xyz1.contiguous(), xyz2=xyz2.contiguous(), R=R, t=t)


def make_sequence_torch(F, seed, device, K=512, n_corr=300, outlier_ratio=0.30, noise=0.002, desc_noise=0.02,
                        dtype="float64"):
    """F consecutive synthetic SR4000 frames (desc (F,K,128), xyz (F,K,3)): frame f+1 re-observes n_corr random
    features of frame f (descriptor + N(0, desc_noise^2) before normalisation, 3-D point moved by the rigid
    motion of the step, a fraction replaced by unrelated points = outliers); its other K - n_corr features are
    new.  Same distributions as make_batch_torch, which draws the two frames of every pair independently.
    Returns dict desc, xyz, R (F-1,3,3), t (F-1,3) with xyz[f] ~ R[f] * xyz[f+1] + t[f] on the inliers."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = torch.float64
    P = F - 1

    def scene(*shape):
        u = torch.rand(*shape, device=device, generator=g, dtype=f64) * (W - 1)
        v = torch.rand(*shape, device=device, generator=g, dtype=f64) * (H_IMG - 1)
        z = torch.rand(*shape, device=device, generator=g, dtype=f64) * 4.2 + 0.8
        return torch.stack([-(u - CX) * z / F_PX, -(v - CY) * z / F_PX, z], dim=-1)

    axis = torch.randn(P, 3, device=device, generator=g, dtype=f64)
    axis = axis / axis.norm(dim=1, keepdim=True)
    ang = torch.rand(P, device=device, generator=g, dtype=f64) * np.deg2rad(5.0)
    Kx = torch.zeros(P, 3, 3, device=device, dtype=f64)
    Kx[:, 0, 1], Kx[:, 0, 2] = -axis[:, 2], axis[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = axis[:, 2], -axis[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -axis[:, 1], axis[:, 0]
    eye = torch.eye(3, device=device, dtype=f64).expand(P, 3, 3)
    R = eye + torch.sin(ang)[:, None, None] * Kx + (1 - torch.cos(ang))[:, None, None] * (Kx @ Kx)
    t = torch.randn(P, 3, device=device, generator=g, dtype=f64)
    t = t * (torch.rand(P, 1, device=device, generator=g, dtype=f64) * 0.10) / t.norm(dim=1, keepdim=True)

    raw = torch.randn(F, K, 128, device=device, generator=g, dtype=torch.float32).abs()
    xyz = scene(F, K)
    i_prev = torch.rand(P, K, device=device, generator=g).argsort(dim=1)[:, :n_corr]   # features of frame f
    i_next = torch.rand(P, K, device=device, generator=g).argsort(dim=1)[:, :n_corr]   # their slots in frame f+1
    pert = desc_noise * torch.randn(P, n_corr, 128, device=device, generator=g, dtype=torch.float32)
    pnoise = noise * torch.randn(P, n_corr, 3, device=device, generator=g, dtype=f64)
    n_out = int(round(outlier_ratio * n_corr))
    outl = scene(P, max(n_out, 1))
    Rt = R.transpose(1, 2)
    for f in range(P):  # frame f+1 depends on frame f
        src = raw[f, i_prev[f]]
        raw[f + 1, i_next[f]] = (src + pert[f]).abs()
        ya = xyz[f, i_prev[f]]                       # previous-frame points
        yb = (ya - t[f]) @ R[f] + pnoise[f]          # R^T (ya - t): the same points seen from frame f+1
        if n_out > 0:
            yb[:n_out] = outl[f, :n_out]
        xyz[f + 1, i_next[f]] = yb
    d = raw / raw.norm(dim=-1, keepdim=True)
    d = d.clamp(max=0.2)
    d = (d / d.norm(dim=-1, keepdim=True)).to(torch.float32)
    if dtype == "float64":
        d = d.to(f64)
    return dict(desc=d.contiguous(), xyz=xyz.contiguous(), R=R, t=t)


def make_sr_frames(seed, F, K, rows=720, n_nan=300, n_near=200, n_lowconf=2000):
    """Synthetic SR4000 frames as `load('d1_%04d.dat')` returns them (M/read_xyz_sr4000.m:3,10-12,26): F matrices of
    rows x 176 doubles [z; x; y; amplitude; confidence(; time stamp)], returned as (F,176,rows) C-contiguous (= column-
    major rows x 176), plus sift-style frames (F,K,4): 0-based (x, y), scale, orientation.  Some pixels are NaN, some
    closer than 0.4 m, some of low confidence; some features sit exactly on .5 positions (round half away)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sr = np.zeros((F, rows, 176))
    for f in range(F):
        u, v = np.meshgrid(np.arange(176.0), np.arange(144.0))
        z = 2.5 + 1.5 * np.sin(u / 40.0 + f) * np.cos(v / 30.0) + rng.normal(scale=0.01, size=(144, 176))
        x = -(u - 91.69) * z / 250.577
        y = -(v - 72.27) * z / 250.577
        pix = rng.permutation(144 * 176)
        near = pix[:n_near]
        z.flat[near] *= 0.05
        x.flat[near] *= 0.05
        y.flat[near] *= 0.05
        nan = pix[n_near:n_near + n_nan]
        x.flat[nan] = np.nan
        z.flat[nan[::3]] = np.nan
        sr[f, 0:144], sr[f, 144:288], sr[f, 288:432] = z, x, y
        sr[f, 432:576] = rng.uniform(0, 20000, size=(144, 176))
        if rows >= 720:
            cm = rng.uniform(20000, 65535, size=(144, 176))
            cm.flat[pix[n_near + n_nan:n_near + n_nan + n_lowconf]] = rng.uniform(0, 30000, n_lowconf)
            sr[f, 576:720] = cm
        if rows == 721:
            sr[f, 720, 0] = 1000.0 + f
    frames = np.zeros((F, K, 4))
    frames[:, :, 0] = rng.uniform(0.0, 175.0, size=(F, K))
    frames[:, :, 1] = rng.uniform(0.0, 143.0, size=(F, K))
    half = rng.random((F, K)) < 0.1
    frames[:, :, 0] = np.where(half, np.floor(frames[:, :, 0]) + 0.5, frames[:, :, 0]).clip(0, 174.5)
    frames[:, :, 1] = np.where(half, np.floor(frames[:, :, 1]) + 0.5, frames[:, :, 1]).clip(0, 142.5)
    frames[:, :, 2] = rng.uniform(1, 4, size=(F, K))
    frames[:, :, 3] = rng.uniform(-np.pi, np.pi, size=(F, K))
    return np.ascontiguousarray(sr.transpose(0, 2, 1)), frames

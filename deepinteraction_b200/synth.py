"""Seeded synthetic nuScenes-like frames for tests and benchmarks.

There is no dataset or checkpoint offline, so every measurement uses frames made
here: a ring of pinhole cameras around the LiDAR origin, a polar "lidar-like" (or
dense) point cloud, and deterministic 0.6 m pillars (first <=20 points per cell in
point order, cells sorted by (b, y, x)) -- the inputs the reference's detector
hands to the hot path (reference models/detectors/deepinteraction.py:132-149).

``sanitize`` removes points whose projection lies within ``margin`` pixels of a
decision boundary of the reference's geometry (strict in-image test, feature-pixel
truncation ``(u/W*w).long()``, z > 1e-5).  Those are measure-zero ties where two
correct fp32 evaluations may legitimately disagree; removing them makes parity
tests bit-stable without changing the workload statistics.
"""
import math

import numpy as np
import torch

PC_RANGE = (-54.0, -54.0, -5.0, 54.0, 54.0, 3.0)


def camera_rig(num_views, in_hw, radius=1.0, height=1.5, focal_frac=0.79):
    """lidar2img (V,4,4) float64 for V pinhole cameras at yaw k*360/V."""
    H, W = in_hw
    K = np.eye(4)
    K[0, 0] = K[1, 1] = focal_frac * W
    K[0, 2], K[1, 2] = W / 2.0, H / 2.0
    base = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])   # cam (x right,y down,z fwd) -> lidar
    mats = []
    for k in range(num_views):
        yaw = 2.0 * math.pi * k / num_views
        c, s = math.cos(yaw), math.sin(yaw)
        Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        R_c2l = Rz @ base
        t = Rz @ np.array([radius, 0.0, height])
        E = np.eye(4)
        E[:3, :3] = R_c2l.T
        E[:3, 3] = -R_c2l.T @ t
        mats.append(K @ E)
    return np.stack(mats)


def camera_extras(num_views, in_hw, radius=1.0, height=1.5, focal_frac=0.79):
    """cam2lidar (V,4,4) and cam_intrinsic (V,3,3) of the same rig (img_metas keys read by the ++ polar block,
    reference fusion_transformerv4.py:521-531)."""
    H, W = in_hw
    base = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    c2l, Ks = [], []
    for k in range(num_views):
        yaw = 2.0 * math.pi * k / num_views
        c, s = math.cos(yaw), math.sin(yaw)
        Rz = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        M = np.eye(4)
        M[:3, :3] = Rz @ base
        M[:3, 3] = Rz @ np.array([radius, 0.0, height])
        c2l.append(M)
        Ks.append(np.array([[focal_frac * W, 0.0, W / 2.0], [0.0, focal_frac * W, H / 2.0], [0.0, 0.0, 1.0]]))
    return np.stack(c2l), np.stack(Ks)


def make_points(n, rng, mode='lidar', rmax=76.0):
    theta = rng.uniform(0.0, 2.0 * math.pi, n)
    if mode == 'lidar':
        r = 1.0 + rng.exponential(8.0, n)
    elif mode == 'dense':
        r = rng.uniform(1.0, rmax, n)
    else:
        raise ValueError(mode)
    x, y = r * np.cos(theta), r * np.sin(theta)
    z = rng.uniform(-3.0, 1.0, n)
    pts = np.stack([x, y, z, rng.uniform(0, 1, n), rng.uniform(0, 1, n)], 1)
    keep = (np.abs(x) < 54.0) & (np.abs(y) < 54.0)
    return pts[keep].astype(np.float32)


def sanitize(points, lidar2img, in_hw, feat_hw, margin=2e-3, pc_range=PC_RANGE):
    """Drop points that sit on a geometric decision boundary for any camera (float64)."""
    H, W = in_hw
    h, w = feat_hw
    p4 = np.concatenate([points[:, :3].astype(np.float64), np.ones((len(points), 1))], 1)
    bad = np.zeros(len(points), bool)
    for M in lidar2img:
        cam = p4 @ M.T
        z = cam[:, 2]
        bad |= np.abs(z - 1e-5) < 1e-3
        zz = np.maximum(z, 1e-5)
        u, v = cam[:, 0] / zz, cam[:, 1] / zz
        vis = (z > 1e-5) & (u > -8) & (u < W + 8) & (v > -8) & (v < H + 8)
        fu, fv = u / W * w, v / H * h                  # feature-pixel coordinate that gets truncated
        near = (np.abs(fu - np.round(fu)) < margin) | (np.abs(fv - np.round(fv)) < margin)
        bad |= vis & near
    return points[~bad]


def pillarize(points_list, pillar=0.6, max_pts=20, pc_range=PC_RANGE):
    """-> pillars (P,max_pts,5) f32, coors (P,4) i32 [b,0,y,x], num_points (P,) i32."""
    nx = int(round((pc_range[3] - pc_range[0]) / pillar))
    ny = int(round((pc_range[4] - pc_range[1]) / pillar))
    P_all, C_all, N_all = [], [], []
    for b, pts in enumerate(points_list):
        ix = np.floor((pts[:, 0].astype(np.float64) - pc_range[0]) / pillar).astype(np.int64)
        iy = np.floor((pts[:, 1].astype(np.float64) - pc_range[1]) / pillar).astype(np.int64)
        ok = (ix >= 0) & (ix < nx) & (iy >= 0) & (iy < ny) & (pts[:, 2] > pc_range[2]) & (pts[:, 2] < pc_range[5])
        idx = np.nonzero(ok)[0]
        cell = iy[idx] * nx + ix[idx]
        order = np.argsort(cell, kind='stable')
        cell_s, idx_s = cell[order], idx[order]
        uniq, start, cnt = np.unique(cell_s, return_index=True, return_counts=True)
        P = len(uniq)
        pil = np.zeros((P, max_pts, pts.shape[1]), np.float32)
        rank = np.arange(len(cell_s)) - np.repeat(start, cnt)
        keep = rank < max_pts
        pid = np.repeat(np.arange(P), cnt)
        pil[pid[keep], rank[keep]] = pts[idx_s[keep]]
        coors = np.stack([np.full(P, b), np.zeros(P, np.int64), uniq // nx, uniq % nx], 1).astype(np.int32)
        P_all.append(pil)
        C_all.append(coors)
        N_all.append(np.minimum(cnt, max_pts).astype(np.int32))
    return np.concatenate(P_all), np.concatenate(C_all), np.concatenate(N_all)


AUG_META = dict(   # SURVEY.md 8(d) config-3 record; also used to test the affine folding
    pcd_rotation=[[math.cos(0.3), -math.sin(0.3), 0.0], [math.sin(0.3), math.cos(0.3), 0.0], [0.0, 0.0, 1.0]],
    pcd_scale_factor=1.05, pcd_trans=[0.2, -0.1, 0.05], pcd_horizontal_flip=True, pcd_vertical_flip=False,
    transformation_3d_flow=['R', 'S', 'T', 'HF'])


def make_frame_batch(seed, batch=1, num_views=6, in_hw=(448, 800), stride=4, c_img=256, c_pts=512,
                     bev_hw=(180, 180), n_points=250000, cloud='lidar', aug=False, device='cpu', sanitize_pts=True):
    """One batch of synthetic hot-path inputs.  sanitize_pts=False keeps the points that sit within 2e-3 feature pixels
    of a decision boundary of the reference geometry (near ties: two correct fp32 evaluations may then disagree on an
    index, so parity is a tolerance statement, not a bit-exact one -- tests/test_gpu_full.py).

    Returns dict(img_feats (B*V,c_img,h,w), pts_feats (B,c_pts,Y,X), img_metas, pts_metas)."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    h, w = in_hw[0] // stride, in_hw[1] // stride
    rig = camera_rig(num_views, in_hw)
    c2l, intr = camera_extras(num_views, in_hw)
    img_metas, pts = [], []
    for b in range(batch):
        p = make_points(n_points, rng, cloud)
        if sanitize_pts:
            p = sanitize(p, rig, in_hw, (h, w))
        pts.append(p)
        meta = dict(lidar2img=[m.astype(np.float32) for m in rig], input_shape=in_hw,
                    img_shape=[(in_hw[0], in_hw[1], 3)] * num_views, box_type_3d=None,
                    cam2lidar=[m.astype(np.float32) for m in c2l], cam_intrinsic=[m.astype(np.float32) for m in intr])
        if aug:
            meta.update(AUG_META)
        img_metas.append(meta)
    if aug:   # the stored cloud is the AUGMENTED one (the hot path undoes the aug before projecting)
        from math import cos, sin
        R = np.asarray(AUG_META['pcd_rotation'], np.float64)
        for b in range(batch):
            q = pts[b].copy().astype(np.float64)
            q[:, :3] = q[:, :3] @ R
            q[:, :3] *= AUG_META['pcd_scale_factor']
            q[:, :3] += np.asarray(AUG_META['pcd_trans'])
            q[:, 1] = -q[:, 1]
            pts[b] = q.astype(np.float32)
    pillars, coors, npts = pillarize(pts)
    out = dict(
        img_feats=torch.randn(batch * num_views, c_img, h, w, generator=g),
        pts_feats=torch.randn(batch, c_pts, bev_hw[0], bev_hw[1], generator=g),
        img_metas=img_metas,
        pts_metas=dict(pillars=torch.from_numpy(pillars), pillar_coors=torch.from_numpy(coors),
                       pillars_num_points=torch.from_numpy(npts), pts=[torch.from_numpy(p) for p in pts]))
    if device != 'cpu':
        out = to_device(out, device)
    return out


def to_device(frame, device):
    pm = frame['pts_metas']
    return dict(img_feats=frame['img_feats'].to(device), pts_feats=frame['pts_feats'].to(device),
                img_metas=frame['img_metas'],
                pts_metas=dict(pillars=pm['pillars'].to(device), pillar_coors=pm['pillar_coors'].to(device),
                               pillars_num_points=pm['pillars_num_points'].to(device),
                               pts=[p.to(device) for p in pm['pts']]))


def randomize_norm_stats(module, seed):
    """Give BN/LN/bias parameters non-trivial values so folding paths are exercised
    (SURVEY.md 8(d) 'Weights')."""
    g = torch.Generator().manual_seed(seed)
    import torch.nn as nn
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
                if m.affine:
                    m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            elif isinstance(m, nn.LayerNorm):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        for name, p in module.named_parameters():
            if name.endswith('bias') and p.dim() == 1 and 'bn' not in name and 'norm' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return module

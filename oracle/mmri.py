"""Oracle: MMRI encoder (multi-modal representational interaction), base model.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain-PyTorch restatement of the
reference modules (paths relative to the reference's projects/mmdet3d_plugin/):

* models/utils/encoder_utils.py            ConvBNReLU :11-34, window ops :36-81,
  LocalContextAttentionBlock :84-135, BEVWarp :137-199, MMRI_P2I :202-213,
  MMRI_I2P :216-320
* models/utils/ops/locatt_ops/kernels.cuh  cc2k :4-42, ck2c_ori :44-80  (window-op
  semantics, incl. the out-of-bounds rules)
* models/necks/deepinteraction_encoder.py  encoder layer :8-33, encoder :35-85

Module/parameter names equal the reference's so a reference state_dict loads
unchanged.  The window ops are written as 81 shifted multiply-accumulates (the
reference's CUDA extension has no CPU path).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .geometry import apply_3d_transformation
from . import depth_completion

torch.backends.mha.set_fastpath_enabled(False)

PC_RANGE = (-54.0, -54.0, -5.0, 54.0, 54.0, 3.0)   # hard-coded in the reference, encoder_utils.py:190


class ConvBNReLU(nn.Module):
    """encoder_utils.py:11-34.  Conv has a bias only when there is no norm."""

    def __init__(self, cin, cout, kernel_size=3, norm=True, act=True, affine=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, 1, (kernel_size - 1) // 2, bias=not norm)
        if norm:
            self.bn = nn.BatchNorm2d(cout, affine=affine)
        self.use_norm, self.use_act = norm, act

    def forward(self, x):
        x = self.conv(x)
        if self.use_norm:
            x = self.bn(x)
        return F.relu(x) if self.use_act else x


def window_similarity(q, k, ksize):
    """locatt ``similar_forward`` (kernels.cuh:4-42): y[n,h,w,t] = sum_c q[n,c,h,w] *
    k[n,c,h+dy,w+dx], t=(dy+r)*ksize+(dx+r); out-of-bounds taps give 0."""
    N, C, H, W = q.shape
    r = ksize // 2
    kp = F.pad(k, (r, r, r, r))
    out = q.new_empty(N, H, W, ksize * ksize)
    for t in range(ksize * ksize):
        dy, dx = divmod(t, ksize)
        out[..., t] = (q * kp[:, :, dy:dy + H, dx:dx + W]).sum(1)
    return out


def window_weighting(v, w, ksize):
    """locatt ``weighting_forward`` (kernels.cuh:44-80): y[n,c,h,w] = sum_t
    v[n,c,h+dy,w+dx] * w[n,h,w,t], out-of-bounds taps skipped."""
    N, C, H, W = v.shape
    r = ksize // 2
    vp = F.pad(v, (r, r, r, r))
    out = torch.zeros_like(v)
    for t in range(ksize * ksize):
        dy, dx = divmod(t, ksize)
        out += vp[:, :, dy:dy + H, dx:dx + W] * w[..., t].unsqueeze(1)
    return out


class LocalContextAttentionBlock(nn.Module):
    """encoder_utils.py:84-135: q = MLP2(target), k = MLP2(source), v = MLP1(source);
    softmax(similar(q,k)/sqrt(C)) over the 9x9 window; weighting with v."""

    def __init__(self, cin, cout, kernel_size, last_affine=True):
        super().__init__()
        self.kernel_size = kernel_size
        self.query_project = nn.Sequential(ConvBNReLU(cin, cout, 1), ConvBNReLU(cout, cout, 1))
        self.key_project = nn.Sequential(ConvBNReLU(cin, cout, 1), ConvBNReLU(cout, cout, 1))
        self.value_project = ConvBNReLU(cin, cout, 1, affine=last_affine)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)

    def forward(self, target, source, chunk=1):
        outs = []
        if self.training:
            chunk = target.shape[0]                     # batch statistics are over the whole call, as in the reference
        for i in range(0, target.shape[0], chunk):      # chunking only bounds CPU memory
            q = self.query_project(target[i:i + chunk])
            k = self.key_project(source[i:i + chunk])
            v = self.value_project(source[i:i + chunk])
            w = window_similarity(q, k, self.kernel_size)
            w = F.softmax(w / math.sqrt(k.size(1)), -1)
            outs.append(window_weighting(v, w, self.kernel_size))
        return torch.cat(outs, 0)


def _lidar2img(img_metas, like):
    return like.new_tensor(np.asarray([m['lidar2img'] for m in img_metas]))


def project_points(pts3, proj, in_hw):
    """Shared by BEVWarp (:157-171) and MMRI_I2P (:281-296): pts3 (N,3), proj (V,4,4)
    -> pixel coords (V,N,2), depth (V,N), strict in-image mask (V,N)."""
    p4 = torch.cat((pts3, torch.ones_like(pts3[..., :1])), -1)[None, :, :, None]      # (1,N,4,1)
    cam = torch.matmul(proj[:, None], p4).squeeze(-1)                                   # (V,N,4)
    z = cam[..., 2]
    eps = 1e-5
    uv = cam[..., 0:2] / torch.maximum(z, torch.ones_like(z) * eps).unsqueeze(-1)
    H, W = in_hw
    nx = (uv[..., 0] / W - 0.5) * 2
    ny = (uv[..., 1] / H - 0.5) * 2
    mask = (z > eps) & (nx > -1.0) & (nx < 1.0) & (ny > -1.0) & (ny < 1.0)
    return uv, z, mask, torch.stack((nx, ny), -1)


def sparse_depth_maps(pts3, proj, in_hw, feat_hw):
    """encoder_utils.py:172-174.  Duplicate pixels: the LAST point in order wins
    (CPU index_put_ behaviour)."""
    uv, z, mask, _ = project_points(pts3, proj, in_hw)
    V = proj.shape[0]
    H, W = in_hw
    h, w = feat_hw
    dm = pts3.new_zeros(V, h, w)
    for v in range(V):
        m = mask[v]
        r = (uv[v, m, 1] / H * h).long()
        c = (uv[v, m, 0] / W * w).long()
        # dm[v, r, c] = z[v, m] with duplicates resolved as a serial index_put_ does (last wins),
        # independent of how many CPU threads torch uses
        lin = r * w + c
        zz = z[v, m]
        winner = torch.full((h * w,), -1, dtype=torch.long)
        winner.scatter_reduce_(0, lin, torch.arange(lin.numel()), 'amax', include_self=True)
        hit = winner >= 0
        flat = dm[v].view(-1)
        flat[hit] = zz[winner[hit]]
    return dm


def lift_pixels(depth, img2lidar, in_hw, img_meta):
    """encoder_utils.py:183-194: (V,h,w) depth -> BEV sampling grid (V,h,w,2) in
    [-1,1] and the strict in-range mask."""
    V, h, w = depth.shape
    H, W = in_hw
    xs = torch.linspace(0, W - 1, w, dtype=torch.float32).view(1, 1, w).expand(V, h, w)
    ys = torch.linspace(0, H - 1, h, dtype=torch.float32).view(1, h, 1).expand(V, h, w)
    xyd = torch.stack((xs * depth, ys * depth, depth, torch.ones_like(depth)), -1)
    xyz = img2lidar.view(V, 1, 1, 4, 4).matmul(xyd.unsqueeze(-1)).squeeze(-1)[..., :3]
    xyz = apply_3d_transformation(xyz.reshape(-1, 3), img_meta, reverse=False).view(V, h, w, 3)
    lo = xyz.new_tensor(PC_RANGE[:3])
    hi = xyz.new_tensor(PC_RANGE[3:])
    mask = ((xyz > lo) & (xyz < hi)).all(-1)
    g = (xyz[..., :2] - lo[:2]) / (hi[:2] - lo[:2])
    return (g - 0.5) * 2, mask, xyz


class BEVWarp(nn.Module):
    """encoder_utils.py:137-199."""

    def forward(self, lidar_feats, img_feats, img_metas, pts_metas, return_aux=False):
        B, V, C, h, w = img_feats.shape
        l2i = _lidar2img(img_metas, img_feats)
        i2l = torch.inverse(l2i)
        outs, aux = [], []
        for b in range(B):
            in_hw = tuple(img_metas[b]['input_shape'][:2])
            pts3 = apply_3d_transformation(pts_metas['pts'][b][..., :3], img_metas[b], reverse=True)
            dm = sparse_depth_maps(pts3, l2i[b], in_hw, (h, w))
            dense = torch.stack([dm.new_tensor(depth_completion.fill_in_multiscale(dm[v].numpy()))
                                 for v in range(V)])
            grid, mask, xyz = lift_pixels(dense, i2l[b], in_hw, img_metas[b])
            warped = F.grid_sample(lidar_feats[b:b + 1].expand(V, -1, -1, -1), grid,
                                   mode='bilinear', padding_mode='zeros', align_corners=False)
            warped = warped * mask.unsqueeze(1)
            outs.append(warped)
            aux.append(dict(sparse=dm, dense=dense, grid=grid, mask=mask, xyz=xyz))
        out = torch.stack(outs, 0)
        return (out, aux) if return_aux else out


class MMRI_P2I(nn.Module):
    """encoder_utils.py:202-213."""

    def __init__(self, cin, cout, kernel_size):
        super().__init__()
        self.Warp = BEVWarp()
        self.Local = LocalContextAttentionBlock(cin, cout, kernel_size)

    def forward(self, lidar_feats, img_feats, img_metas, pts_metas):
        warped = self.Warp(lidar_feats, img_feats, img_metas, pts_metas)
        B, V, C, h, w = warped.shape
        return self.Local(img_feats.reshape(B * V, C, h, w), warped.reshape(B * V, C, h, w)).view(B, V, C, h, w)


class MMRI_I2P(nn.Module):
    """encoder_utils.py:216-320.  The reference buckets pillars by valid-key count
    (``group_attn`` :226-255) purely to bound padding; the result equals one dense
    masked single-head attention, which is what is evaluated here (pillar chunks
    bound memory only)."""

    def __init__(self, pts_channels, img_channels, dropout):
        super().__init__()
        self.pts_channels, self.img_channels = pts_channels, img_channels
        self.learnedAlign = nn.MultiheadAttention(pts_channels, 1, dropout=dropout, kdim=img_channels,
                                                  vdim=img_channels, batch_first=True)

    def forward(self, lidar_feat, img_feat, img_metas, pts_metas, chunk=2048):
        B = len(img_metas)
        out = torch.zeros_like(lidar_feat)
        l2i = _lidar2img(img_metas, lidar_feat)
        coors_all = pts_metas['pillar_coors']
        for b in range(B):
            sel = coors_all[:, 0] == b
            pillars = pts_metas['pillars'][sel]
            coors = coors_all[sel].long()
            npts = pts_metas['pillars_num_points'][sel].long()
            P, T, _ = pillars.shape
            V = l2i.shape[1]
            pts3 = apply_3d_transformation(pillars.reshape(P * T, -1)[:, :3], img_metas[b], reverse=True)
            in_hw = tuple(img_metas[b]['input_shape'][:2])
            _, _, mask, grid = project_points(pts3, l2i[b], in_hw)          # (V,PT), (V,PT,2)
            mask = mask.view(V, P, T).permute(1, 2, 0)                      # (P,T,V)
            mask = mask & (torch.arange(T)[None, :, None] < npts[:, None, None])
            mask = mask.reshape(P, T * V)                                   # key index = point*V + cam (:298,309)
            q_all = lidar_feat[b][:, coors[:, 2], coors[:, 3]].t()          # (P,C)   (:313)
            res = lidar_feat.new_zeros(P, self.pts_channels)
            valid = mask.any(1)
            idx = valid.nonzero().squeeze(1)
            for s in range(0, idx.numel(), chunk):
                ii = idx[s:s + chunk]
                n = ii.numel()
                g = grid.view(V, P, T, 2)[:, ii].reshape(V, n * T, 1, 2)
                samp = F.grid_sample(img_feat[b], g, mode='bilinear', padding_mode='zeros',
                                     align_corners=False).squeeze(-1)      # (V,C,nT)
                kv = samp.permute(2, 0, 1).reshape(n, T * V, self.img_channels)
                att = self.learnedAlign(q_all[ii].unsqueeze(1), kv, kv,
                                        attn_mask=~mask[ii].unsqueeze(1))[0]
                res[ii] = att.squeeze(1)
            out[b][:, coors[:, 2], coors[:, 3]] = res.t()
        return out


class DeepInteractionEncoderLayer(nn.Module):
    """deepinteraction_encoder.py:8-33."""

    def __init__(self, c):
        super().__init__()
        self.I2P_block = MMRI_I2P(c, c, 0.1)
        self.P_IML = LocalContextAttentionBlock(c, c, 9)
        self.P_out_proj = ConvBNReLU(2 * c, c, 1, act=False)
        self.P_integration = ConvBNReLU(2 * c, c, 1, act=False)
        self.P2I_block = MMRI_P2I(c, c, 9)
        self.I_IML = LocalContextAttentionBlock(c, c, 9)
        self.I_out_proj = ConvBNReLU(2 * c, c, 1, act=False)
        self.I_integration = ConvBNReLU(2 * c, c, 1, act=False)

    def forward(self, img_feat, lidar_feat, img_metas, pts_metas, return_parts=False):
        B = lidar_feat.shape[0]
        BN, C, h, w = img_feat.shape
        img5 = img_feat.view(B, -1, C, h, w)
        i2p = self.I2P_block(lidar_feat, img5, img_metas, pts_metas)
        p2p = self.P_IML(lidar_feat, lidar_feat)
        p_aug = self.P_out_proj(torch.cat((i2p, p2p), 1))
        new_lidar = self.P_integration(torch.cat((p_aug, lidar_feat), 1))
        p2i = self.P2I_block(lidar_feat, img5, img_metas, pts_metas)
        i2i = self.I_IML(img_feat, img_feat)
        i_aug = self.I_out_proj(torch.cat((p2i.view(BN, -1, h, w), i2i), 1))
        new_img = self.I_integration(torch.cat((i_aug, img_feat), 1))
        if return_parts:
            return new_img, new_lidar, dict(i2p=i2p, p2p=p2p, p2i=p2i.view(BN, -1, h, w), i2i=i2i)
        return new_img, new_lidar


class DeepInteractionEncoder(nn.Module):
    """deepinteraction_encoder.py:35-85."""

    def __init__(self, num_layers=2, in_channels_img=64, in_channels_pts=128 * 3, hidden_channel=128,
                 bn_momentum=0.1, bias='auto'):
        super().__init__()
        use_bias = True if bias == 'auto' else bool(bias)     # mmcv build_conv_layer passes 'auto' through (truthy)
        self.shared_conv_pts = nn.Conv2d(in_channels_pts, hidden_channel, 3, padding=1, bias=use_bias)
        self.shared_conv_img = nn.Conv2d(in_channels_img, hidden_channel, 3, padding=1, bias=use_bias)
        self.num_layers = num_layers
        self.fusion_blocks = nn.ModuleList(DeepInteractionEncoderLayer(hidden_channel) for _ in range(num_layers))
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.momentum = bn_momentum

    def forward(self, img_feats, pts_feats, img_metas, pts_metas):
        img = self.shared_conv_img(img_feats)
        pts = self.shared_conv_pts(pts_feats)
        pts_conv = pts.clone()
        for blk in self.fusion_blocks:
            img, pts = blk(img, pts, img_metas, pts_metas)
        return img, [pts_conv, pts]

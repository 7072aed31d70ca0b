"""Oracle: the ++ ("deformable") MMRI encoder, DeepInteraction++ (BASELINE.json config 4).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain-PyTorch restatement, fp32, CPU-runnable, of

* models/necks/fusion_transformerv4.py   FusionTransformerv4 :25-138, DeepInteractionLayer :142-218,
  MMRI_P2I (BEVWarp + 1-level deformable attention) :220-240, MMRI_I2P (+ residual) :242-364,
  TransSinePositionalEncoding :367-485, MMRI_I2P_Polar :487-640, FlashMultiheadAttention :716-759
* the mmcv 1.3.18 bricks those classes are assembled from, which are NOT under /root/reference and are therefore
  restated from their published behaviour (SURVEY.md Appendix C.4 / C.5 -- "parity unpinned" for them):
  MultiScaleDeformableAttention (value_proj, sampling_offsets / attention_weights from the query, softmax over
  levels x points, bilinear zero-padded sampling at ref + offset / (W_l, H_l), output_proj, + identity),
  FFN (Linear-ReLU-Linear + identity), BaseTransformerLayer (attentions / ffns / norms ModuleLists).
* flash-attn 0.2.2 (fp16 kernel) is evaluated as exact fp32 softmax attention here; the reference's fp16 rounding
  (FlashAttention.fp16_enabled, fusion_transformerv4.py:665-667) is NOT reproduced -- documented tolerance 2e-3 on
  the polar block against an fp16-emulating stub.

Module / parameter names equal the reference's (incl. mmcv's `attentions.N`, `ffns.N.layers.0.0`, `norms.N`,
`transformer_layers.decoder.layers.0...`) so a released ++ checkpoint loads unchanged.
"""
import copy
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geometry import apply_3d_transformation
from .mmri import BEVWarp, MMRI_I2P as _BaseI2P, _lidar2img

PC_RANGE = (-54.0, -54.0, -5.0, 54.0, 54.0, 3.0)


# ---------------------------------------------------------------------------------------------------------------
# mmcv 1.3.18 bricks (Appendix C.4 / C.5)
# ---------------------------------------------------------------------------------------------------------------
def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value (bs, sum HW, heads, d); sampling_locations (bs, nq, heads, L, P, 2) in [0,1]; weights (bs,nq,heads,L,P)."""
    bs, _, heads, d = value.shape
    _, nq, _, L, P, _ = sampling_locations.shape
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    value_list = value.split(sizes, dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        v = value_list[lvl].flatten(2).transpose(1, 2).reshape(bs * heads, d, int(h), int(w))
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)                         # (bs*heads, nq, P, 2)
        sampled.append(F.grid_sample(v, g, mode='bilinear', padding_mode='zeros', align_corners=False))
    w = attention_weights.transpose(1, 2).reshape(bs * heads, 1, nq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * w).sum(-1).view(bs, heads * d, nq)
    return out.transpose(1, 2).contiguous()


class MultiScaleDeformableAttention(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, im2col_step=64, dropout=0.1,
                 batch_first=False, norm_cfg=None, init_cfg=None):
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.batch_first = batch_first
        self.dropout = nn.Dropout(dropout)
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid.view(-1)
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if identity is None:
            identity = query
        if query_pos is not None:
            query = query + query_pos
        assert self.batch_first
        bs, nq, _ = query.shape
        nv = value.shape[1]
        value = self.value_proj(value).view(bs, nv, self.num_heads, -1)
        off = self.sampling_offsets(query).view(bs, nq, self.num_heads, self.num_levels, self.num_points, 2)
        aw = self.attention_weights(query).view(bs, nq, self.num_heads, self.num_levels * self.num_points).softmax(-1)
        aw = aw.view(bs, nq, self.num_heads, self.num_levels, self.num_points)
        shapes = [(int(h), int(w)) for h, w in spatial_shapes]
        normalizer = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype)
        loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
        out = ms_deform_attn_core(value, shapes, loc, aw)
        return self.dropout(self.output_proj(out)) + identity


class FFN(nn.Module):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0., **kw):
        super().__init__()
        assert num_fcs == 2
        self.embed_dims = embed_dims
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    def forward(self, x, identity=None):
        return (x if identity is None else identity) + self.layers(x)


# ---------------------------------------------------------------------------------------------------------------
# fusion_transformerv4.py
# ---------------------------------------------------------------------------------------------------------------
class MMRI_P2I_Deform(nn.Module):
    """fusion_transformerv4.py:220-240 (registered as `MMRI_P2I` in the ATTENTION registry)."""

    def __init__(self, embed_dims, batch_first=True):
        super().__init__()
        self.embed_dims = embed_dims
        self.Warp = BEVWarp()
        self.Local = MultiScaleDeformableAttention(embed_dims, num_levels=1, batch_first=batch_first)

    def forward(self, img_feats, lidar_feats, img_metas, pts_metas, reference_points=None, **kw):
        B = lidar_feats.size(0)
        _, C, H, W = img_feats.shape
        warped = self.Warp(lidar_feats, img_feats.reshape(B, -1, C, H, W), img_metas, pts_metas)
        q = img_feats.flatten(-2).permute(0, 2, 1)
        v = warped.reshape(-1, C, H * W).permute(0, 2, 1)
        out = self.Local(query=q, value=v, reference_points=reference_points, spatial_shapes=[(H, W)])
        return out.permute(0, 2, 1).reshape(-1, C, H, W)


class MMRI_I2P_Res(_BaseI2P):
    """fusion_transformerv4.py:242-364: the base MMRI_I2P (same projection / masking / single-head attention; the
    group_attn bucketing :262-293 is a padding device) plus the residual `+ lidar_feat` (:364)."""

    def __init__(self, embed_dims, dropout, batch_first=True, fp16_enabled=False, flash_attn=False,
                 group_attn_enabled=False):
        assert not flash_attn
        super().__init__(embed_dims, embed_dims, dropout)
        self.embed_dims = embed_dims

    def forward(self, lidar_feat, img_feat, img_metas, pts_metas, **kw):
        B = lidar_feat.size(0)
        _, C, H, W = img_feat.shape
        return super().forward(lidar_feat, img_feat.reshape(B, -1, C, H, W), img_metas, pts_metas) + lidar_feat


def sine_pos(x_range, y_range, num_feats, temperature=10000.):
    """TransSinePositionalEncoding.forward, normalize=False, z_pos=None (:420-485): -> (bs, 2*num_feats, y_len, x_len)."""
    x_len, y_len = x_range.shape[-1], y_range.shape[-1]
    x_embed = x_range.unsqueeze(-2).repeat(1, y_len, 1)
    y_embed = y_range.unsqueeze(-1).repeat(1, 1, x_len)
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x, pos_y = x_embed[:, :, :, None] / dim_t, y_embed[:, :, :, None] / dim_t
    B, H, W = x_range.shape[0], y_len, x_len
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class _MHA(nn.Module):
    """FlashMultiheadAttention (:716-759) with exact fp32 softmax attention (scale 1/sqrt(head_dim), no mask)."""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, q, k, v):
        wq, wk, wv = self.in_proj_weight.chunk(3)
        bq, bk, bv = self.in_proj_bias.chunk(3)
        h = self.num_heads
        split = lambda t: t.view(t.shape[0], t.shape[1], h, -1).transpose(1, 2)
        q, k, v = split(F.linear(q, wq, bq)), split(F.linear(k, wk, bk)), split(F.linear(v, wv, bv))
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), -1) @ v
        return self.out_proj(a.transpose(1, 2).reshape(a.shape[0], -1, self.embed_dim))


class _DecoderLayer(nn.Module):
    """nn.TransformerDecoderLayer (post-norm, ReLU) with the two attentions replaced (:762-768)."""

    def __init__(self, d, nhead, ff):
        super().__init__()
        self.self_attn, self.multihead_attn = _MHA(d, nhead), _MHA(d, nhead)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)

    def forward(self, tgt, memory):
        x = self.norm1(tgt + self.self_attn(tgt, tgt, tgt))
        x = self.norm2(x + self.multihead_attn(x, memory, memory))
        return self.norm3(x + self.linear2(F.relu(self.linear1(x))))


class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.norm = nn.LayerNorm(d)


class _Decoder(nn.Module):
    def __init__(self, d, nhead, ff):
        super().__init__()
        self.layers = nn.ModuleList([_DecoderLayer(d, nhead, ff)])
        self.norm = nn.LayerNorm(d)


class _Transformer(nn.Module):
    """nn.Transformer(num_encoder_layers=0, custom_decoder=...): the (layer-less) encoder still applies its final
    LayerNorm to src (Appendix C.6); state_dict keys `encoder.norm.*`, `decoder.layers.0.*`, `decoder.norm.*`."""

    def __init__(self, d, nhead, ff):
        super().__init__()
        self.encoder, self.decoder = _Norm(d), _Decoder(d, nhead, ff)

    def forward(self, src, tgt):
        memory = self.encoder.norm(src)
        return self.decoder.norm(self.decoder.layers[0](tgt, memory))


class MMRI_I2P_Polar(nn.Module):
    """fusion_transformerv4.py:487-640: per camera, every image COLUMN (H tokens) is decoded into a polar RAY
    (R = 60 radius bins) by one transformer decoder layer whose queries are the BEV features sampled along that
    column's viewing ray; the rays are then resampled onto the BEV grid (mean over 10 heights of the projected
    pixel / radius) and averaged over the cameras that see a cell; + residual."""

    def __init__(self, embed_dims, dropout, batch_first=True, radius_range=(1., 61., 1.0), num_decoder_layers=1,
                 pc_range=PC_RANGE):
        super().__init__()
        assert num_decoder_layers == 1
        self.embed_dims, self.radius_range, self.pc_range = embed_dims, list(radius_range), list(pc_range)
        self.radius = int((radius_range[1] - radius_range[0]) / radius_range[-1])
        self.im_scale = 4.
        self.transformer_layers = _Transformer(embed_dims, 8, embed_dims * 4)

    def forward(self, lidar_feat, img_feat, img_metas, pts_metas, **kw):
        B = lidar_feat.size(0)
        _, C, H, W = img_feat.shape
        R = self.radius
        img_feat = img_feat.reshape(B, -1, C, H, W)
        out = torch.zeros_like(lidar_feat)
        visibles = torch.zeros_like(out[:, 0:1])
        l2i = _lidar2img(img_metas, lidar_feat)
        cam2lidar = lidar_feat.new_tensor([m['cam2lidar'] for m in img_metas])
        pc = self.pc_range
        f32 = dict(dtype=torch.float32)
        for cam in range(img_feat.shape[1]):
            feat = img_feat[:, cam]
            xr = torch.arange(0., float(W), 1.).unsqueeze(0).repeat(B, 1)
            img_pos = sine_pos(xr, torch.arange(0., float(H), 1.).unsqueeze(0).repeat(B, 1), C // 2)
            ray_pos = sine_pos(xr, torch.arange(0., float(R), 1.).unsqueeze(0).repeat(B, 1), C // 2)
            cam_coors = torch.stack([xr + 0.5, torch.zeros_like(xr) + H // 2, torch.ones_like(xr), torch.ones_like(xr)], 1)
            cam_coors[:, :2] *= self.im_scale
            img2lidar = torch.linalg.inv(l2i[:, cam])
            ray_dirs = torch.bmm(img2lidar, cam_coors)[:, :2] - cam2lidar[:, cam, :2, -1:]
            ray_dirs = ray_dirs / ray_dirs.norm(dim=1, p=2, keepdim=True)
            depths = torch.arange(self.radius_range[0], self.radius_range[1], self.radius_range[2]) + self.radius_range[2] / 2
            centers = (depths[None, None, :, None].to(ray_dirs) * ray_dirs[:, :, None]).permute(0, 2, 3, 1)   # b r w c
            norm_xy = []
            for b in range(B):
                c3 = torch.cat([centers[b].reshape(R * W, 2), torch.zeros(R * W, 1)], -1)
                c3 = apply_3d_transformation(c3, img_metas[b], reverse=False).view(R, W, 3)
                norm_xy.append(torch.stack([(c3[..., 0] - pc[0]) / (pc[3] - pc[0]), (c3[..., 1] - pc[1]) / (pc[4] - pc[1])], -1))
            norm_xy = torch.stack(norm_xy, 0)
            polar_query = F.grid_sample(lidar_feat, norm_xy * 2 - 1, align_corners=False) + ray_pos      # (B,C,R,W)
            rays = polar_query.permute(2, 0, 3, 1).flatten(1, 2).transpose(0, 1)                        # (B*W, R, C)
            cols = (feat + img_pos).permute(2, 0, 3, 1).flatten(1, 2).transpose(0, 1)                   # (B*W, H, C)
            bev_out = self.transformer_layers(cols, rays).view(B, W, R, C).permute(0, 3, 2, 1)          # (B,C,R,W)
            xs, ys, zs = lidar_feat.shape[-2], lidar_feat.shape[-1], 10
            by, bx, bz = torch.meshgrid(torch.linspace(0, xs - 1, xs) + 0.5, torch.linspace(0, ys - 1, ys) + 0.5,
                                        torch.linspace(0, zs - 1, zs) + 0.5, indexing='ij')
            bx = bx / xs * (pc[3] - pc[0]) + pc[0]
            by = by / ys * (pc[4] - pc[1]) + pc[1]
            bz = bz / zs * (pc[5] - pc[2]) + pc[2]
            bev_pts = torch.stack([bx, by, bz], -1).reshape(-1, 3)
            reaug = torch.stack([apply_3d_transformation(bev_pts, img_metas[b], reverse=True) for b in range(B)], 0)
            hom = torch.cat([reaug, torch.ones_like(reaug[..., :1])], -1).transpose(1, 2)               # (B,4,N)
            xyz = torch.bmm(l2i[:, cam], hom)[:, :3].transpose(1, 2)
            eps = 1e-5
            mask = xyz[..., 2:3] > eps
            xy = xyz[..., 0:2] / torch.maximum(xyz[..., 2:3], torch.ones_like(xyz[..., 2:3]) * eps)
            in_hw = img_metas[0]['input_shape']
            xy = torch.stack([xy[..., 0] / in_hw[1], xy[..., 1] / in_hw[0]], -1) * 2 - 1
            mask = mask & (xy[..., 0:1] > -1.0) & (xy[..., 0:1] < 1.0) & (xy[..., 1:2] > -1.0) & (xy[..., 1:2] < 1.0)
            radius_map = torch.norm(hom[:, :2, :] - cam2lidar[:, cam, :2, -1:], dim=1)
            norm_r = (2 * (radius_map - self.radius_range[0]) / self.radius - 1).clamp(-1, 1)
            loc = torch.stack([xy[..., 0], norm_r], -1).view(B, ys, xs, zs, 2).mean(dim=3)
            m = mask.view(B, ys, xs, zs, 1).sum(dim=3).permute(0, 3, 1, 2) > 0
            out = out + F.grid_sample(bev_out, loc, align_corners=False) * m
            visibles = visibles + m
        visibles[visibles == 0] = 1
        return out / visibles + lidar_feat


ATTENTIONS = dict(MultiScaleDeformableAttention=MultiScaleDeformableAttention, MMRI_P2I=MMRI_P2I_Deform,
                  MMRI_I2P=MMRI_I2P_Res, MMRI_I2P_Polar=MMRI_I2P_Polar)


class DeepInteractionLayer(nn.Module):
    """fusion_transformerv4.py:142-218 on top of mmcv's BaseTransformerLayer (post-norm: operation_order[0] != 'norm')."""

    def __init__(self, attn_cfgs, ffn_cfgs, operation_order=None, norm_cfg=None, batch_first=True, **kw):
        super().__init__()
        self.operation_order = tuple(operation_order)
        self.pre_norm = self.operation_order[0] == 'norm'
        assert not self.pre_norm and batch_first
        self.attentions = nn.ModuleList()
        for cfg in attn_cfgs:
            cfg = dict(cfg)
            self.attentions.append(ATTENTIONS[cfg.pop('type')](**cfg))
        self.embed_dims = self.attentions[0].embed_dims
        ffn_cfg = {k: v for k, v in dict(ffn_cfgs).items() if k != 'type'}
        self.ffns = nn.ModuleList(FFN(**copy.deepcopy(ffn_cfg)) for _ in range(self.operation_order.count('ffn')))
        self.norms = nn.ModuleList(nn.LayerNorm(self.embed_dims) for _ in range(self.operation_order.count('norm')))
        self.scale = nn.Parameter(torch.ones(1))

    def forward(self, query, value, ms_query, reference_points, spatial_shapes, level_start_index, img_metas, pts_metas):
        q_h, q_w = query.shape[-2:]
        v_h, v_w = value.shape[-2:]
        query = query.flatten(-2).permute(0, 2, 1)
        value = value.flatten(-2).reshape(-1, self.embed_dims, v_h, v_w)
        ni = ai = fi = 0
        self_feat = None
        for layer in self.operation_order[:-2]:
            if layer == 'self_attn':
                query = self.attentions[ai](query=query, value=ms_query, reference_points=reference_points,
                                            spatial_shapes=spatial_shapes, level_start_index=level_start_index)
                ai += 1
                self_feat = query
            elif layer == 'norm':
                query = self.norms[ni](query)
                ni += 1
            elif layer == 'cross_attn':
                q4 = query.permute(0, 2, 1).reshape(-1, self.embed_dims, q_h, q_w)
                q4 = self.attentions[ai](q4, value, img_metas=img_metas, pts_metas=pts_metas,
                                         reference_points=reference_points[:, :, 0:1, :], spatial_shapes=spatial_shapes,
                                         level_start_index=level_start_index)
                query = q4.reshape(-1, self.embed_dims, q_h * q_w).permute(0, 2, 1)
                ai += 1
            elif layer == 'ffn':
                query = self.ffns[fi](query)
                fi += 1
        for layer in self.operation_order[-2:]:
            if layer == 'norm':
                self_feat = self.norms[ni](self_feat)
                ni += 1
            elif layer == 'ffn':
                self_feat = self.ffns[fi](self_feat)
                fi += 1
        query = self_feat + self.scale * query
        return query.permute(0, 2, 1).reshape(-1, self.embed_dims, q_h, q_w)


def reference_points(H, W):
    """FusionTransformerv4.get_reference_points :129-138: pixel centres normalised by (W, H) -> (1, H*W, 2) = (x, y)."""
    ry, rx = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing='ij')
    return torch.stack((rx.reshape(-1)[None] / W, ry.reshape(-1)[None] / H), -1)


class FusionTransformerv4(nn.Module):
    """fusion_transformerv4.py:25-138.  forward(img_feats: list of levels (B*V,Ci,h_l,w_l), pts_feats: list
    [concat (B, Cp*num_lidar_maps, Y, X), map_1 (B,Cp,Y,X), ...], img_metas, pts_metas)
    -> (new_img (B*V,C,h_0,w_0), [pts_conv, new_pts])."""

    def __init__(self, num_layers=2, num_lidar_maps=2, in_channels_img=64, in_channels_pts=128 * 3, hidden_channel=128,
                 bn_momentum=0.1, bias='auto', img_transformerlayers=None, pts_transformerlayers=None):
        super().__init__()
        use_bias = True if bias == 'auto' else bool(bias)
        self.shared_conv_pts = nn.Conv2d(in_channels_pts * num_lidar_maps, hidden_channel, 3, padding=1, bias=use_bias)
        self.multi_scale_conv_img = nn.Conv2d(in_channels_img, hidden_channel, 3, padding=1, bias=use_bias)
        self.multi_scale_conv_pts = nn.Conv2d(in_channels_pts, hidden_channel, 3, padding=1, bias=use_bias)
        self.num_layers = num_layers
        mk = lambda cfg: DeepInteractionLayer(**{k: v for k, v in copy.deepcopy(dict(cfg)).items() if k != 'type'})
        self.img_fusion_blocks = nn.ModuleList(mk(img_transformerlayers) for _ in range(num_layers))
        self.pts_fusion_blocks = nn.ModuleList(mk(pts_transformerlayers) for _ in range(num_layers))

    def forward(self, img_feats, pts_feats, img_metas, pts_metas):
        pts_feats = list(pts_feats)
        pts_conv = self.shared_conv_pts(pts_feats.pop(0))
        ms_img = [self.multi_scale_conv_img(f) for f in img_feats]
        ms_pts = [self.multi_scale_conv_pts(f) for f in pts_feats]
        new_img, new_pts = ms_img[0], ms_pts[0]

        def flat(ms):
            shapes = [tuple(f.shape[-2:]) for f in ms]
            return torch.cat([f.flatten(-2).permute(0, 2, 1) for f in ms], -2), shapes
        img_flat, shapes_img = flat(ms_img)
        pts_flat, shapes_pts = flat(ms_pts)
        ref_img = reference_points(*new_img.shape[-2:]).unsqueeze(-2).repeat(1, 1, len(ms_img), 1)
        ref_pts = reference_points(*new_pts.shape[-2:]).unsqueeze(-2).repeat(1, 1, len(ms_pts), 1)
        for i in range(self.num_layers):
            t_img = self.img_fusion_blocks[i](new_img, new_pts, img_flat, ref_img, shapes_img, None, img_metas, pts_metas)
            t_pts = self.pts_fusion_blocks[i](new_pts, new_img, pts_flat, ref_pts, shapes_pts, None, img_metas, pts_metas)
            new_img, new_pts = t_img, t_pts
        return new_img, [pts_conv, new_pts]

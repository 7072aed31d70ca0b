"""Oracle: MMPI decoder (multi-modal predictive interaction), base model, forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain-PyTorch restatement of the
reference modules (paths relative to the reference's projects/mmdet3d_plugin/):

* models/utils/decoder_utils.py   PositionEmbeddingLearned :16-32,
  TransformerDecoderLayer :35-113, MultiheadAttention / multi_head_attention_forward
  :116-495, FFN :498-581, DynamicConv :584-629, ImageRCNNBlock :632-761,
  PointRCNNBlock :765-841
* models/dense_heads/deepinteraction_decoder.py   __init__ :21-160, forward :201-313
* core/bbox/coders/transfusion_bbox_coder.py      decode :39-126

mmcv's ConvModule / build_conv_layer are replaced by equivalent torch modules with
the same attribute (and therefore state_dict) names: ``conv``/``bn`` inside a
ConvModule, plain ``nn.Conv*`` for build_conv_layer (bias='auto' -> bias=True).
detectron2's ROIPooler is ``oracle.geometry.roi_align``; mmdet3d's box corners and
``apply_3d_transformation`` are in ``oracle.geometry``.
"""
import copy

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .geometry import apply_3d_transformation, lidar_box_corners, roi_align

torch.backends.mha.set_fastpath_enabled(False)


class ConvModule1d(nn.Module):
    """mmcv 1.3.18 ConvModule(conv_cfg=Conv1d, norm_cfg=BN1d): conv(no bias)-bn-relu."""

    def __init__(self, cin, cout, k=1):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, k, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm1d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class ConvModule2d(nn.Module):
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class PositionEmbeddingLearned(nn.Module):
    """decoder_utils.py:16-32."""

    def __init__(self, cin, c):
        super().__init__()
        self.position_embedding_head = nn.Sequential(nn.Conv1d(cin, c, 1), nn.BatchNorm1d(c), nn.ReLU(inplace=True),
                                                     nn.Conv1d(c, c, 1))

    def forward(self, xyz):
        return self.position_embedding_head(xyz.transpose(1, 2).contiguous())


class MultiheadAttention(nn.Module):
    """decoder_utils.py:116-495 (packed in-projection path; q scaled by
    head_dim**-0.5 AFTER the bias add, :407)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value):
        L, N, E = query.shape
        S = key.shape[0]
        h = self.num_heads
        d = E // h
        w, b = self.in_proj_weight, self.in_proj_bias
        q = F.linear(query, w[:E], b[:E]) * float(d) ** -0.5
        k = F.linear(key, w[E:2 * E], b[E:2 * E])
        v = F.linear(value, w[2 * E:], b[2 * E:])
        q = q.contiguous().view(L, N * h, d).transpose(0, 1)
        k = k.contiguous().view(S, N * h, d).transpose(0, 1)
        v = v.contiguous().view(S, N * h, d).transpose(0, 1)
        att = F.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
        out = torch.bmm(att, v).transpose(0, 1).contiguous().view(L, N, E)
        return self.out_proj(out)


class TransformerDecoderLayer(nn.Module):
    """decoder_utils.py:35-113 (eval: dropouts are identities)."""

    def __init__(self, d_model, nhead, dim_ff, self_posembed, cross_posembed):
        super().__init__()
        self.self_attn = MultiheadAttention(d_model, nhead)
        self.multihead_attn = MultiheadAttention(d_model, nhead)
        self.linear1 = nn.Linear(d_model, dim_ff)
        self.linear2 = nn.Linear(dim_ff, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.self_posembed = self_posembed
        self.cross_posembed = cross_posembed

    def forward(self, query, key, query_pos, key_pos):
        qpe = self.self_posembed(query_pos).permute(2, 0, 1)
        kpe = self.cross_posembed(key_pos).permute(2, 0, 1)
        query = query.permute(2, 0, 1)
        key = key.permute(2, 0, 1)
        x = query + qpe
        query = self.norm1(query + self.self_attn(x, x, x))
        kk = key + kpe
        query = self.norm2(query + self.multihead_attn(query + qpe, kk, kk))
        query = self.norm3(query + self.linear2(F.relu(self.linear1(query))))
        return query.permute(1, 2, 0)


class PredictionFFN(nn.Module):
    """decoder_utils.py:498-581 (class ``FFN``): per head ConvModule(in->64) then Conv1d(64->k)."""

    def __init__(self, cin, heads, head_conv=64):
        super().__init__()
        self.heads = heads
        for name, (classes, num_conv) in heads.items():
            layers, c = [], cin
            for _ in range(num_conv - 1):
                layers.append(ConvModule1d(c, head_conv))
                c = head_conv
            layers.append(nn.Conv1d(head_conv, classes, 1))
            setattr(self, name, nn.Sequential(*layers))

    def forward(self, x):
        return {name: getattr(self, name)(x) for name in self.heads}


class TransFusionBBoxCoder:
    """core/bbox/coders/transfusion_bbox_coder.py:8-126 (decode, filter=False path)."""

    def __init__(self, pc_range, out_size_factor, voxel_size, post_center_range=None, score_threshold=None,
                 code_size=8):
        self.pc_range, self.out_size_factor, self.voxel_size = pc_range, out_size_factor, voxel_size
        self.post_center_range, self.score_threshold, self.code_size = post_center_range, score_threshold, code_size

    def decode_boxes(self, rot, dim, center, height, vel):
        """-> (B, P, 7 or 9) boxes (x, y, z_bottom, dx, dy, dz, yaw[, vx, vy])."""
        cx = center[:, 0:1] * self.out_size_factor * self.voxel_size[0] + self.pc_range[0]
        cy = center[:, 1:2] * self.out_size_factor * self.voxel_size[1] + self.pc_range[1]
        dim = dim.exp()
        height = height - dim[:, 2:3] * 0.5
        yaw = torch.atan2(rot[:, 0:1], rot[:, 1:2])
        parts = [cx, cy, height, dim, yaw] + ([vel] if vel is not None else [])
        return torch.cat(parts, 1).permute(0, 2, 1)

    def encode(self, dst_boxes):
        """transfusion_bbox_coder.py:24-38: (n, 7|9) boxes (x, y, z_bottom, dx, dy, dz, yaw[, vx, vy]) -> (n, code_size)."""
        t = torch.zeros(dst_boxes.shape[0], self.code_size)
        t[:, 0] = (dst_boxes[:, 0] - self.pc_range[0]) / (self.out_size_factor * self.voxel_size[0])
        t[:, 1] = (dst_boxes[:, 1] - self.pc_range[1]) / (self.out_size_factor * self.voxel_size[1])
        t[:, 3:6] = dst_boxes[:, 3:6].log()
        t[:, 2] = dst_boxes[:, 2] + dst_boxes[:, 5] * 0.5
        t[:, 6], t[:, 7] = torch.sin(dst_boxes[:, 6]), torch.cos(dst_boxes[:, 6])
        if self.code_size == 10:
            t[:, 8:10] = dst_boxes[:, 7:]
        return t

    def decode(self, heatmap, rot, dim, center, height, vel, filter=False):
        """transfusion_bbox_coder.py:40-126 (without the reference's in-place writes into `center` / `dim`).
        -> list over samples of dict(bboxes (n, 7|9), scores (n,), labels (n,) int64); filter=True keeps
        score > score_threshold (when it is truthy, :110) and centres inside post_center_range (:101-105)."""
        scores, labels = heatmap.max(1)
        boxes = self.decode_boxes(rot, dim, center, height, vel)
        if not filter:
            return [dict(bboxes=boxes[i], scores=scores[i], labels=labels[i]) for i in range(heatmap.shape[0])]
        if self.post_center_range is None:
            raise NotImplementedError('Need to reorganize output as a batch, only support post_center_range is not None for now!')
        rng = torch.tensor(self.post_center_range, dtype=boxes.dtype)
        mask = (boxes[..., :3] >= rng[:3]).all(2) & (boxes[..., :3] <= rng[3:]).all(2)
        out = []
        for i in range(heatmap.shape[0]):
            m = mask[i]
            if self.score_threshold:
                m = m & (scores[i] > self.score_threshold)
            out.append(dict(bboxes=boxes[i, m], scores=scores[i, m], labels=labels[i, m]))
        return out


def circle_nms(dets, thresh, post_max_size=83):
    """mmdet3d v0.17.1 mmdet3d/core/post_processing/box3d_nms.py::circle_nms (third party, NOT in the reference
    tree: parity unpinned).  dets (n, 3) = x, y, score; `thresh` is compared with the SQUARED centre distance;
    returns the kept indices in descending score order, at most post_max_size."""
    x, y, sc = dets[:, 0], dets[:, 1], dets[:, 2]
    order = np.argsort(-sc, kind='stable')
    suppressed = np.zeros(len(sc), bool)
    keep = []
    for a in range(len(order)):
        i = order[a]
        if suppressed[i]:
            continue
        keep.append(int(i))
        for b in range(a + 1, len(order)):
            j = order[b]
            if not suppressed[j] and (x[i] - x[j]) ** 2 + (y[i] - y[j]) ** 2 <= thresh:
                suppressed[j] = True
    return keep[:post_max_size]


NMS_TASKS = {   # deepinteraction_decoder.py:575-586
    'nuScenes': [dict(indices=[0, 1, 2, 3, 4, 5, 6, 7], radius=-1), dict(indices=[8], radius=0.175),
                 dict(indices=[9], radius=0.175)],
    'Waymo': [dict(indices=[0], radius=0.7), dict(indices=[1], radius=0.7), dict(indices=[2], radius=0.7)],
}


class DynamicConv(nn.Module):
    """decoder_utils.py:584-629."""

    def __init__(self):
        super().__init__()
        self.hidden_dim, self.dim_dynamic, self.num_dynamic = 128, 128, 2
        self.num_params = self.hidden_dim * self.dim_dynamic
        self.dynamic_layer = nn.Linear(self.hidden_dim, self.num_dynamic * self.num_params)
        self.norm1 = nn.LayerNorm(self.dim_dynamic)
        self.norm2 = nn.LayerNorm(self.hidden_dim)
        self.out_layer = nn.Linear(self.hidden_dim * 49, self.hidden_dim)
        self.norm3 = nn.LayerNorm(self.hidden_dim)

    def forward(self, pro_features, roi_features):
        # pro_features (1, n, C), roi_features (49, n, C)
        feats = roi_features.permute(1, 0, 2)
        params = self.dynamic_layer(pro_features).permute(1, 0, 2)
        p1 = params[:, :, :self.num_params].reshape(-1, self.hidden_dim, self.dim_dynamic)
        p2 = params[:, :, self.num_params:].reshape(-1, self.dim_dynamic, self.hidden_dim)
        feats = F.relu(self.norm1(torch.bmm(feats, p1)))
        feats = F.relu(self.norm2(torch.bmm(feats, p2)))
        feats = self.out_layer(feats.flatten(1))
        return F.relu(self.norm3(feats))


def _rcnn_tail(blk, sfx, q_view, roi):
    """Shared tail of both RCNN blocks (decoder_utils.py:743-756 / :824-837):
    q_view (n,1,C) sequence-first, roi (n,C,7,7)."""
    g = lambda name: getattr(blk, name + sfx)
    roi = roi.flatten(2).permute(2, 0, 1)
    q2 = g('dyconv_pre_self_attn')(q_view, q_view, value=q_view)[0]
    q_view = g('norm1')(q_view + q2)
    q_view = q_view.permute(1, 0, 2)
    q2 = g('dyconv')(q_view, roi)
    q_view = g('norm2')(q_view + q2)
    q2 = g('linear2')(F.gelu(g('linear1')(q_view)))
    q_view = g('norm3')(q_view + q2)
    return q_view[0]                      # (n, C)


def _rcnn_params(blk, sfx, c, heads, dropout):
    setattr(blk, 'dyconv' + sfx, DynamicConv())
    setattr(blk, 'dyconv_pre_self_attn' + sfx, nn.MultiheadAttention(c, heads, dropout=dropout))
    for i in (1, 2, 3):
        setattr(blk, f'norm{i}' + sfx, nn.LayerNorm(c))
    setattr(blk, 'linear1' + sfx, nn.Linear(c, 4 * c))
    setattr(blk, 'linear2' + sfx, nn.Linear(4 * c, c))


class ImageRCNNBlock(nn.Module):
    """decoder_utils.py:632-761."""

    def __init__(self, num_views, num_proposals, out_size_factor_img, test_cfg, bbox_coder, c, heads, dropout):
        super().__init__()
        self.num_views, self.num_proposals = num_views, num_proposals
        self.out_size_factor_img, self.test_cfg, self.bbox_coder = out_size_factor_img, test_cfg, bbox_coder
        _rcnn_params(self, '', c, heads, dropout)

    def _tail(self, q_view, roi):
        return _rcnn_tail(self, '', q_view, roi)

    def forward(self, query_feat, res_layer, new_lidar_feat, img_feat_flatten, img_metas, img_h, img_w):
        B = query_feat.shape[0]
        P = self.num_proposals
        prev = query_feat
        out = torch.zeros_like(query_feat)
        center = res_layer['center'].detach()
        real = center * self.test_cfg['out_size_factor'] * self.test_cfg['voxel_size'][0] + self.test_cfg['pc_range'][0]
        pos3d = torch.cat([real, res_layer['height'].detach()], 1)                         # (B,3,P)
        boxes = self.bbox_coder.decode_boxes(res_layer['rot'].detach(), res_layer['dim'].detach(), center,
                                             res_layer['height'].detach(), res_layer.get('vel'))
        on_mask = torch.ones(B, P) * -1
        rects_dbg = []
        for b in range(B):
            meta = img_metas[b]
            l2i = pos3d.new_tensor(np.asarray(meta['lidar2img']))
            corners = lidar_box_corners(boxes[b][:, :7])                                   # (P,8,3)
            pts = torch.cat([pos3d[b], corners.permute(2, 0, 1).reshape(3, -1)], -1).T     # (P+8P,3)
            pts = apply_3d_transformation(pts, meta, reverse=True)
            h, w = meta['input_shape'][:2]
            crop = pos3d.new_tensor(meta['img_crop_offset']) if 'img_crop_offset' in meta else 0
            flip = meta.get('flip', False)
            for v in range(self.num_views):
                p4 = torch.cat([pts, pts.new_ones(pts.shape[0], 1)], -1)
                p2 = p4 @ l2i[v].t()
                z = torch.clamp(p2[:, 2], min=1e-5)
                xy = torch.stack([p2[:, 0] / z, p2[:, 1] / z], -1) - crop
                cx, cy = xy[:, 0], xy[:, 1]
                if flip:
                    cx = meta['img_shape'][0][1] - cx
                ctr_x, ctr_y = cx[:P], cy[:P]
                cor_x, cor_y = cx[P:].reshape(P, 8), cy[P:].reshape(P, 8)
                on = (ctr_x > 0) & (ctr_x < w) & (ctr_y > 0) & (ctr_y < h)
                if on.sum() <= 1:
                    continue
                on_mask[b, on] = v
                rect = torch.stack([cor_x[on].min(1).values, cor_y[on].min(1).values,
                                    cor_x[on].max(1).values, cor_y[on].max(1).values], 1)
                fmap = img_feat_flatten[b, v].reshape(-1, img_h, img_w)
                roi = roi_align(fmap, rect, 7, 1.0 / self.out_size_factor_img, 2)
                q_view = prev[b][:, on].t().unsqueeze(1)                                   # (n,1,C)
                out[b][:, on] = self._tail(q_view, roi).t()
                rects_dbg.append((b, v, on.nonzero().squeeze(1), rect))
        self._dbg = rects_dbg
        return out, on_mask


class PointRCNNBlock(nn.Module):
    """decoder_utils.py:765-841."""

    def __init__(self, c, heads, dropout, bbox_coder):
        super().__init__()
        self.bbox_coder = bbox_coder
        _rcnn_params(self, '_pts', c, heads, dropout)

    def _tail(self, q_view, roi):
        return _rcnn_tail(self, '_pts', q_view, roi)

    def forward(self, query_feat, res_layer, new_lidar_feat, img_feat_flatten, img_metas, img_h, img_w):
        B = query_feat.shape[0]
        out = torch.zeros_like(query_feat)
        boxes = self.bbox_coder.decode_boxes(res_layer['rot'].detach(), res_layer['dim'].detach(),
                                             res_layer['center'].detach(), res_layer['height'].detach(),
                                             res_layer.get('vel'))
        bc = self.bbox_coder
        for b in range(B):
            box = boxes[b][:, :7].clone()
            box[:, 3:6] *= 2
            cc = (lidar_box_corners(box)[..., :2] - bc.pc_range[0]) / (bc.voxel_size[0] * bc.out_size_factor)
            rect = torch.stack([cc[..., 0].min(-1).values, cc[..., 1].min(-1).values,
                                cc[..., 0].max(-1).values, cc[..., 1].max(-1).values], -1)
            roi = roi_align(new_lidar_feat[b], rect, 7, 1.0, 2)
            q_view = query_feat[b].t().unsqueeze(1)
            out[b] = self._tail(q_view, roi).t()
        return out, None


class DeepInteractionDecoder(nn.Module):
    """models/dense_heads/deepinteraction_decoder.py:19-313 (forward only; the config
    of projects/configs/nuscenes/Fusion_0075_refactor.py:194-224 is the default)."""

    def __init__(self, num_views=6, out_size_factor_img=4, num_proposals=200, auxiliary=True, hidden_channel=128,
                 num_classes=10, num_mmpi=4, num_decoder_layers=1, num_heads=8, nms_kernel_size=3, ffn_channel=256,
                 dropout=0.1, bn_momentum=0.1, common_heads=None, num_heatmap_convs=2, bbox_coder=None,
                 test_cfg=None, **unused):
        super().__init__()
        if common_heads is None:
            common_heads = dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2))
        self.num_classes, self.num_proposals, self.auxiliary = num_classes, num_proposals, auxiliary
        self.num_views, self.num_mmpi, self.nms_kernel_size = num_views, num_mmpi, nms_kernel_size
        self.num_decoder_layers = num_decoder_layers
        self.test_cfg = test_cfg
        self.bbox_coder = TransFusionBBoxCoder(**{k: v for k, v in bbox_coder.items() if k != 'type'})
        c = hidden_channel
        self.heatmap_head = nn.Sequential(ConvModule2d(c, c, 3), nn.Conv2d(c, num_classes, 3, padding=1))
        self.heatmap_head_img = copy.deepcopy(self.heatmap_head)
        self.class_encoding = nn.Conv1d(num_classes, c, 1)
        self.decoder = nn.ModuleList(
            TransformerDecoderLayer(c, num_heads, ffn_channel, PositionEmbeddingLearned(2, c),
                                    PositionEmbeddingLearned(2, c)) for _ in range(num_decoder_layers))
        heads = dict(common_heads)
        heads['heatmap'] = (num_classes, num_heatmap_convs)
        self.prediction_heads = nn.ModuleList(PredictionFFN(c, heads) for _ in range(num_decoder_layers))
        self.decode_head = nn.ModuleList()
        self.pred_head = nn.ModuleList()
        for _ in range(num_mmpi // 2):
            self.decode_head.append(ImageRCNNBlock(num_views, num_proposals, out_size_factor_img, test_cfg,
                                                   self.bbox_coder, c, num_heads, dropout))
            self.pred_head.append(PredictionFFN(2 * c, heads))
            self.decode_head.append(PointRCNNBlock(c, num_heads, dropout, self.bbox_coder))
            self.pred_head.append(PredictionFFN(2 * c, heads))
        xs = test_cfg['grid_size'][0] // test_cfg['out_size_factor']
        ys = test_cfg['grid_size'][1] // test_cfg['out_size_factor']
        self.bev_pos = self.create_2D_grid(xs, ys)
        for p in self.decoder.parameters():          # :171-175
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm1d)):
                m.momentum = bn_momentum

    @staticmethod
    def create_2D_grid(x_size, y_size):
        """:162-169 -> (1, x*y, 2) with (x+0.5, y+0.5), flatten index = y*x_size + x."""
        by, bx = torch.meshgrid(torch.linspace(0, x_size - 1, x_size), torch.linspace(0, y_size - 1, y_size),
                                indexing='ij')
        return torch.stack([bx + 0.5, by + 0.5], 0).view(1, 2, -1).permute(0, 2, 1)

    def select_proposals(self, dense_heatmap, dense_heatmap_img):
        """:225-244 -> masked heatmap (B,K,HW), top-k flat indices (B,P)."""
        heatmap = (dense_heatmap.sigmoid() + dense_heatmap_img.sigmoid()) / 2
        pad = self.nms_kernel_size // 2
        local_max = torch.zeros_like(heatmap)
        inner = F.max_pool2d(heatmap, self.nms_kernel_size, stride=1, padding=0)
        local_max[:, :, pad:-pad, pad:-pad] = inner
        if self.test_cfg['dataset'] == 'nuScenes':
            local_max[:, 8] = heatmap[:, 8]
            local_max[:, 9] = heatmap[:, 9]
        elif self.test_cfg['dataset'] == 'Waymo':
            local_max[:, 1] = heatmap[:, 1]
            local_max[:, 2] = heatmap[:, 2]
        heatmap = heatmap * (heatmap == local_max)
        B = heatmap.shape[0]
        heatmap = heatmap.view(B, heatmap.shape[1], -1)
        top = heatmap.view(B, -1).argsort(dim=-1, descending=True)[..., :self.num_proposals]
        return heatmap, top

    def _mmpi(self, query_feat, res, first_res, new_lidar_feat, img_flat, img_metas, ih, iw, aux):
        """:279-296: the alternating image / point RCNN layers (overridden by the ++ decoder)."""
        self.on_the_image_mask = []
        rets = []
        for l in range(self.num_mmpi):
            prev = query_feat.clone()
            query_pos = res['center'].detach().clone().permute(0, 2, 1)
            query_feat, on = self.decode_head[l](prev, res, new_lidar_feat, img_flat, img_metas, ih, iw)
            res = self.pred_head[l](torch.cat([query_feat, prev], 1))
            res['center'] = res['center'] + query_pos.permute(0, 2, 1)
            if l % 2 == 0:
                keep = on != -1
                self.on_the_image_mask.append(keep)
                for key in res:
                    m = (~keep).unsqueeze(1).expand_as(res[key])
                    res[key] = torch.where(m, first_res[key], res[key])
                aux.setdefault('on_view', []).append(on.clone())
            aux['layer_query'].append(query_feat.clone())
            rets.append(res)
        return rets

    def forward(self, pts_inputs, img_inputs, img_metas, return_aux=False):
        lidar_feat, new_lidar_feat = pts_inputs
        B, C = lidar_feat.shape[:2]
        flat = lidar_feat.view(B, C, -1)
        bev_pos = self.bev_pos.repeat(B, 1, 1)
        BN, _, ih, iw = img_inputs.shape
        dense_heatmap = self.heatmap_head(lidar_feat)
        dense_heatmap_img = self.heatmap_head_img(new_lidar_feat)
        heatmap, top = self.select_proposals(dense_heatmap.detach(), dense_heatmap_img.detach())
        hw = heatmap.shape[-1]
        top_class, top_index = top // hw, top % hw
        query_feat = flat.gather(-1, top_index[:, None, :].expand(-1, C, -1))
        self.query_labels = top_class
        one_hot = F.one_hot(top_class, self.num_classes).permute(0, 2, 1).float()
        query_feat = query_feat + self.class_encoding(one_hot)
        query_pos = bev_pos.gather(1, top_index[:, :, None].expand(-1, -1, 2))
        aux = dict(top=top, heatmap=heatmap, query_feat0=query_feat.clone(), query_pos0=query_pos.clone())
        for i in range(self.num_decoder_layers):
            query_feat = self.decoder[i](query_feat, flat, query_pos, bev_pos)
            res = self.prediction_heads[i](query_feat)
            res['center'] = res['center'] + query_pos.permute(0, 2, 1)
            first_res = res
            query_pos = res['center'].detach().clone().permute(0, 2, 1)
        aux['query_feat1'] = query_feat.clone()
        aux['first_res'] = {k: v.clone() for k, v in first_res.items()}
        img_flat = img_inputs.view(B, self.num_views, C, -1)
        aux['layer_query'] = []
        rets = self._mmpi(query_feat, res, first_res, new_lidar_feat, img_flat, img_metas, ih, iw, aux)
        rets[0]['query_heatmap_score'] = heatmap.gather(-1, top_index[:, None, :].expand(-1, self.num_classes, -1))
        rets[0]['dense_heatmap'] = dense_heatmap_img
        if not self.auxiliary:
            out = [rets[-1]]
        else:
            merged = {}
            for key in rets[0]:
                if key in ('dense_heatmap', 'dense_heatmap_old', 'query_heatmap_score'):
                    merged[key] = rets[0][key]
                else:
                    merged[key] = torch.cat([r[key] for r in rets], -1)
            out = [[merged]]
        return (out, aux) if return_aux else out

    def get_bboxes(self, preds_dicts, img_metas, img=None, rescale=False, for_roi=False):
        """deepinteraction_decoder.py:549-638: last-layer scores = sigmoid(heatmap) * query_heatmap_score * one-hot
        of the query label, decode(filter=True), optional per-task circle NMS; one layer, batch size 1 (:631-632).
        -> [[boxes, scores, labels.int()]] (boxes wrapped by img_metas[0]['box_type_3d'] when that key exists)."""
        rets = []
        for preds in preds_dicts:
            p0, P = preds[0], self.num_proposals
            score = p0['heatmap'][..., -P:].sigmoid()
            one_hot = F.one_hot(self.query_labels, num_classes=self.num_classes).permute(0, 2, 1)
            score = score * p0['query_heatmap_score'] * one_hot
            vel = p0['vel'][..., -P:] if 'vel' in p0 else None
            temp = self.bbox_coder.decode(score, p0['rot'][..., -P:], p0['dim'][..., -P:], p0['center'][..., -P:],
                                          p0['height'][..., -P:], vel, filter=True)
            layer = []
            for t in temp:
                boxes, scores, labels = t['bboxes'], t['scores'], t['labels']
                if self.test_cfg['nms_type'] is not None:
                    if self.test_cfg['nms_type'] != 'circle':
                        raise NotImplementedError('only nms_type None / circle are restated (rotated NMS = mmdet3d nms_gpu)')
                    keep = torch.zeros_like(scores, dtype=torch.bool)
                    for task in NMS_TASKS[self.test_cfg['dataset']]:
                        tm = torch.zeros_like(keep)
                        for c in task['indices']:
                            tm |= labels == c
                        idx = torch.where(tm)[0]
                        if task['radius'] > 0:
                            dets = torch.cat([boxes[tm][:, :2], scores[tm][:, None]], 1).detach().numpy()
                            idx = idx[torch.tensor(circle_nms(dets, task['radius']), dtype=torch.long)]
                        keep[idx] = True
                    boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
                layer.append(dict(bboxes=boxes, scores=scores, labels=labels))
            rets.append(layer)
        assert len(rets) == 1 and len(rets[0]) == 1
        r = rets[0][0]
        wrap = img_metas[0].get('box_type_3d') if isinstance(img_metas[0], dict) else None
        boxes = wrap(r['bboxes'], box_dim=r['bboxes'].shape[-1]) if wrap is not None else r['bboxes']
        return [[boxes, r['scores'], r['labels'].int()]]

"""Oracle: the ++ MMPI decoder (DeepInteraction++, BASELINE.json config 4), forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain-PyTorch restatement, fp32, CPU-runnable, of

* models/utils/decoder_utils.py   ImageRCNNBlockV2 :844-991, PointRCNNBlockV2 :995-1089
* models/dense_heads/deepinteractionplusplus_decoder.py   __init__ :21-160 (V2 blocks :135-147, prediction heads on C
  rather than 2C channels), forward :200-319 (look-forward centre update :281-294, cumulative on-image mask and the
  first-layer fallback applied at EVERY layer :295-302)

on top of oracle.mmpi (everything the two decoders share).  mmcv 1.3.18's FFN (`TransFFN`, not under /root/reference)
is oracle.mmri_pp.FFN: Linear-ReLU-Linear + identity, parameters `layers.0.0`, `layers.1`.

One behaviour of the V2 blocks is restated LITERALLY rather than "as intended": after norm1 the reference keeps
`self_feat_view` in the sequence-first layout (n, 1, C) while the main path has been permuted to (1, n, C), so
`query_feat_view * scale + self_feat_view * self_scale` (decoder_utils.py:987, :1085) broadcasts to (n, n, C) and the
`[0]` that follows keeps row 0: EVERY query of the view (sample, for the point block) receives the self-branch
feature of the view's FIRST query.  The same tensor expression is used here, so the broadcast happens by itself; the
golden generated from the reference's own file (tools/make_goldens.py G7) pins it.
"""
import torch
import torch.nn as nn

from . import mmpi as base
from .geometry import lidar_box_corners, roi_align
from .mmri_pp import FFN as TransFFN


def _v2_params(blk, sfx, c, heads, dropout):
    """decoder_utils.py:859-882 / :1004-1026: note that `ffn`, `self_ffn`, `scale`, `self_scale` carry no `_pts` suffix."""
    setattr(blk, 'dyconv' + sfx, base.DynamicConv())
    setattr(blk, 'dyconv_pre_self_attn' + sfx, nn.MultiheadAttention(c, heads, dropout=dropout))
    for i in (1, 2, 3):
        setattr(blk, f'norm{i}' + sfx, nn.LayerNorm(c))
    setattr(blk, 'self_norm' + sfx, nn.LayerNorm(c))
    blk.ffn = TransFFN(c, 4 * c, 2, dict(type='ReLU', inplace=True), dropout)
    blk.self_ffn = TransFFN(c, 4 * c, 2, dict(type='ReLU', inplace=True), dropout)
    blk.scale = nn.Parameter(torch.ones(1) * 0.5)
    blk.self_scale = nn.Parameter(torch.ones(1) * 0.5)


def _v2_tail(blk, sfx, q_view, roi):
    """decoder_utils.py:971-988 / :1071-1085.  q_view (n,1,C) sequence-first, roi (n,C,7,7) -> (n?, n, C) as the
    reference leaves it; the callers index it exactly as the reference does."""
    g = lambda name: getattr(blk, name + sfx)
    roi = roi.flatten(2).permute(2, 0, 1)
    q2 = g('dyconv_pre_self_attn')(q_view, q_view, value=q_view)[0]
    q_view = g('norm1')(q_view + q2)
    self_view = q_view.clone()                                  # (n, 1, C)
    q_view = q_view.permute(1, 0, 2)                            # (1, n, C)
    q2 = g('dyconv')(q_view, roi)
    q_view = g('norm2')(q_view + q2)
    q_view = g('norm3')(blk.ffn(q_view))
    self_view = g('self_norm')(blk.self_ffn(self_view))
    return q_view * blk.scale + self_view * blk.self_scale      # (1,n,C) + (n,1,C) -> (n, n, C)


class ImageRCNNBlockV2(base.ImageRCNNBlock):
    """decoder_utils.py:844-991: geometry identical to ImageRCNNBlock (:632-761); the per-view tail differs."""

    def __init__(self, num_views, num_proposals, out_size_factor_img, test_cfg, bbox_coder, c, heads, dropout):
        nn.Module.__init__(self)
        self.num_views, self.num_proposals = num_views, num_proposals
        self.out_size_factor_img, self.test_cfg, self.bbox_coder = out_size_factor_img, test_cfg, bbox_coder
        _v2_params(self, '', c, heads, dropout)

    def _tail(self, q_view, roi):
        return _v2_tail(self, '', q_view, roi)[0]               # :990  `query_feat_view[0].permute(1, 0)` -> (n, C)


class PointRCNNBlockV2(base.PointRCNNBlock):
    """decoder_utils.py:995-1089."""

    def __init__(self, c, heads, dropout, bbox_coder):
        nn.Module.__init__(self)
        self.bbox_coder = bbox_coder
        _v2_params(self, '_pts', c, heads, dropout)

    def _tail(self, q_view, roi):
        return _v2_tail(self, '_pts', q_view, roi).permute(0, 2, 1)[0].t()   # :1086 `.permute(0,2,1)[0]` is (C, n)


class DeepInteractionPlusPlusDecoder(base.DeepInteractionDecoder):
    """models/dense_heads/deepinteractionplusplus_decoder.py:19-319 (forward only)."""

    def __init__(self, num_views=6, out_size_factor_img=4, num_proposals=200, hidden_channel=128, num_mmpi=4,
                 num_heads=8, dropout=0.1, **kw):
        super().__init__(num_views=num_views, out_size_factor_img=out_size_factor_img, num_proposals=num_proposals,
                         hidden_channel=hidden_channel, num_mmpi=num_mmpi, num_heads=num_heads, dropout=dropout, **kw)
        c = hidden_channel
        heads = dict(self.prediction_heads[0].heads)
        self.decode_head, self.pred_head = nn.ModuleList(), nn.ModuleList()
        for _ in range(num_mmpi // 2):
            self.decode_head.append(ImageRCNNBlockV2(num_views, num_proposals, out_size_factor_img, self.test_cfg,
                                                     self.bbox_coder, c, num_heads, dropout))
            self.pred_head.append(base.PredictionFFN(c, heads))                    # :140
            self.decode_head.append(PointRCNNBlockV2(c, num_heads, dropout, self.bbox_coder))
            self.pred_head.append(base.PredictionFFN(c, heads))                    # :147

    def _mmpi(self, query_feat, res, first_res, new_lidar_feat, img_flat, img_metas, ih, iw, aux):
        """:279-303."""
        self.on_the_image_mask = []
        rets = []
        look = res['center'].detach().clone()                                      # :281
        for l in range(self.num_mmpi):
            prev = query_feat.clone()
            query_pos = res['center'].detach().clone()                             # :285 (B,2,P), not permuted
            query_feat, on = self.decode_head[l](prev, res, new_lidar_feat, img_flat, img_metas, ih, iw)
            res = self.pred_head[l](query_feat)                                    # :291
            delta = res['center'].clone()
            res['center'] = delta + look                                           # :293
            look = delta + query_pos                                               # :294
            if l % 2 == 0:
                keep = on != -1
                if l > 0:
                    keep = keep & self.on_the_image_mask[-1]
                self.on_the_image_mask.append(keep)                                # :297
                aux.setdefault('on_view', []).append(on.clone())
            else:
                self.on_the_image_mask.append(self.on_the_image_mask[-1])          # :299
            keep = self.on_the_image_mask[-1]
            for key in res:                                                        # :300-302, every layer
                m = (~keep).unsqueeze(1).expand_as(res[key])
                res[key] = torch.where(m, first_res[key], res[key])
            aux['layer_query'].append(query_feat.clone())
            rets.append(res)
        return rets

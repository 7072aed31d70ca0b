"""``DeepInteractionEncoder`` (NECKS) -- the MMRI encoder, backed by libdi_b200.

Interface of the reference class (projects/mmdet3d_plugin/models/necks/deepinteraction_encoder.py:35-85):
``DeepInteractionEncoder(num_layers, in_channels_img, in_channels_pts, hidden_channel, bn_momentum, bias)``,
``forward(img_feats, pts_feats, img_metas, pts_metas) -> (new_img_feat, [pts_feat_conv, new_pts_feat])``.
"""
from deepinteraction_b200.mmri import DeepInteractionEncoder as _Engine, DeepInteractionEncoderLayer  # noqa: F401
from ...registry import NECKS


@NECKS.register_module()
class DeepInteractionEncoder(_Engine):
    pass

"""``FusionTransformerv4`` (NECKS), ``DeepInteractionLayer`` (TRANSFORMER_LAYER), ``MMRI_P2I`` / ``MMRI_I2P`` /
``MMRI_I2P_Polar`` (ATTENTION) -- the ++ ("deformable") MMRI encoder of DeepInteraction++, backed by libdi_b200.

Interface of the reference classes (projects/mmdet3d_plugin/models/necks/fusion_transformerv4.py:25-138,142-218,
220-240,242-364,487-640): ``FusionTransformerv4(num_layers, num_lidar_maps, in_channels_img, in_channels_pts,
hidden_channel, bn_momentum, bias, img_transformerlayers, pts_transformerlayers)``,
``forward(img_feats: list, pts_feats: list, img_metas, pts_metas) -> (new_img_feat, [pts_feat_conv, new_pts_feat])``.
The transformer-layer / attention classes are parameter holders with the reference's state_dict layout; the kernel
schedule lives in deepinteraction_b200/mmri_pp.py and polar.py.
"""
from deepinteraction_b200 import mmri_pp as _pp
from ...registry import NECKS, TRANSFORMER_LAYER, ATTENTION


@NECKS.register_module()
class FusionTransformerv4(_pp.FusionTransformerv4):
    pass


@TRANSFORMER_LAYER.register_module()
class DeepInteractionLayer(_pp.DeepInteractionLayer):
    pass


@ATTENTION.register_module()
class MMRI_P2I(_pp.MMRI_P2I):
    pass


@ATTENTION.register_module()
class MMRI_I2P(_pp.MMRI_I2P):
    pass


@ATTENTION.register_module()
class MMRI_I2P_Polar(_pp.MMRI_I2P_Polar):
    pass

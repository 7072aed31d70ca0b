"""``DeepInteractionPlusPlusDecoder`` (HEADS) -- the ++ MMPI decoder forward, backed by libdi_b200.

Interface of the reference class (projects/mmdet3d_plugin/models/dense_heads/deepinteractionplusplus_decoder.py:19-319):
constructor kwargs of Fusion_0075_plusplus.py's ``pts_bbox_head``, ``forward(pts_inputs, img_inputs, img_metas) ->
[[dict]]``, side attributes ``query_labels`` / ``on_the_image_mask`` (one cumulative mask per MMPI layer).  ``get_bboxes``,
``get_targets`` and ``loss`` (forward values) as for the base decoder, with the ++ per-layer mask rule (:513-514).
"""
from deepinteraction_b200.mmpi import DeepInteractionPlusPlusDecoder as _Engine
from ...registry import HEADS


@HEADS.register_module()
class DeepInteractionPlusPlusDecoder(_Engine):
    pass

"""``DeepInteractionDecoder`` (HEADS) -- the MMPI decoder forward, backed by libdi_b200.

Interface of the reference class (projects/mmdet3d_plugin/models/dense_heads/deepinteraction_decoder.py:19-313):
constructor kwargs of Fusion_0075_refactor.py:194-224 (+ train_cfg / test_cfg), ``forward(pts_inputs,
img_inputs, img_metas) -> [[dict]]``, side attributes ``query_labels`` / ``on_the_image_mask``.
``get_bboxes`` (:549-638) is provided (nms_type None / 'circle'); ``get_targets`` / ``loss`` (:315-547) return the
forward VALUES (Hungarian assignment, targets and the three losses on the GPU, deepinteraction_b200/loss.py); gradients
are outside this repository's scope.
"""
from deepinteraction_b200.mmpi import DeepInteractionDecoder as _Engine
from ...registry import HEADS


@HEADS.register_module()
class DeepInteractionDecoder(_Engine):
    pass

"""``TransFusionBBoxCoder`` (BBOX_CODERS): box parametrisation constants used by the decoder kernels
(reference core/bbox/coders/transfusion_bbox_coder.py:7-22)."""
from deepinteraction_b200.mmpi import TransFusionBBoxCoder as _Coder
from ....registry import BBOX_CODERS


@BBOX_CODERS.register_module()
class TransFusionBBoxCoder(_Coder):
    pass

"""``HungarianAssigner3D`` / ``HeuristicAssigner3D`` (BBOX_ASSIGNERS) and ``BBox3DL1Cost`` / ``BBoxBEVL1Cost`` /
``IoU3DCost`` (MATCH_COST), the names `train_cfg.pts.assigner` of Fusion_0075_*.py refers to (reference
core/bbox/assigners/hungarian_assigner.py:14-153).  The matching runs on the GPU (deepinteraction_b200/loss.py:
di_match_cost_f32 + di_hungarian_f32) instead of scipy on the host."""
from deepinteraction_b200 import loss as _loss
from ....registry import BBOX_ASSIGNERS, MATCH_COST


@MATCH_COST.register_module()
class BBox3DL1Cost(_loss.BBox3DL1Cost):
    pass


@MATCH_COST.register_module()
class BBoxBEVL1Cost(_loss.BBoxBEVL1Cost):
    pass


@MATCH_COST.register_module()
class IoU3DCost(_loss.IoU3DCost):
    pass


@BBOX_ASSIGNERS.register_module()
class HeuristicAssigner3D(_loss.HeuristicAssigner3D):
    pass


@BBOX_ASSIGNERS.register_module()
class HungarianAssigner3D(_loss.HungarianAssigner3D):
    pass

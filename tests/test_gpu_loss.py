"""GPU parity of the MMPI loss path (libdi_b200: di_match_cost_f32, di_hungarian_f32, di_loss_targets_f32,
di_gaussian_heatmap_f32, di_mmpi_losses_f32) against the reference's outputs (goldens) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')


def dev():
    return torch.device('cuda:0')


def _random_boxes(n, g, spread=20.0):
    xy = torch.rand(n, 2, generator=g) * 2 * spread - spread
    z = torch.rand(n, 1, generator=g) * 2 - 2.5
    dims = torch.stack([torch.rand(n, generator=g) * 2 + 1, torch.rand(n, generator=g) * 4 + 2, torch.rand(n, generator=g) + 1.2], 1)
    yaw = torch.rand(n, 1, generator=g) * 6.2 - 3.1
    return torch.cat([xy, z, dims, yaw, torch.randn(n, 2, generator=g)], 1)


def _head(plusplus=False, proposals=24, train_cfg=None):
    from deepinteraction_b200 import mmpi
    from tools import make_goldens as mg
    cls = mmpi.DeepInteractionPlusPlusDecoder if plusplus else mmpi.DeepInteractionDecoder
    return mg.make_decoder(cls, proposals=proposals, train_cfg=dict(train_cfg or mg.DEC_TRAIN_CFG),
                           loss_bbox=mg.DEC_LOSSES['loss_bbox'], loss_heatmap=mg.DEC_LOSSES['loss_heatmap']).to(dev()).eval()


def test_match_cost_and_iou_match_oracle():
    import oracle.loss as ol
    from deepinteraction_b200 import loss as L
    from tools.make_goldens import DEC_TRAIN_CFG
    g = torch.Generator().manual_seed(11)
    P, K = 150, 10
    pred = _random_boxes(P, g, 12.0)
    gt = torch.cat([_random_boxes(20, g, 12.0), pred[:10] + 0.2 * torch.randn(10, 9, generator=g)], 0)
    gt[:, 3:6] = gt[:, 3:6].abs() + 0.5
    gl = torch.randint(0, K, (gt.shape[0],), generator=g)
    score = torch.randn(1, K, P, generator=g) * 2
    a = ol.HungarianAssigner3D(**{k: v for k, v in DEC_TRAIN_CFG['assigner'].items() if k != 'type'})
    iou = a.iou_calculator(pred, gt)
    cost = a.cls_cost(score[0].T, gl) + a.reg_cost(pred, gt, DEC_TRAIN_CFG) + a.iou_cost(iou)
    A = L.HungarianAssigner3D(**{k: v for k, v in DEC_TRAIN_CFG['assigner'].items() if k != 'type'})
    gtp, glp, n = L.pad_gt([gt], [gl], dev())
    gt_inds, ov, c2, i2 = A.assign_batch(pred[None].to(dev()), score.to(dev()), gtp, glp, n, 1, DEC_TRAIN_CFG)
    assert float((i2[0].cpu() - iou).abs().max()) < 2e-6 and float(iou.max()) > 0.3
    assert float((c2[0].cpu() - cost).abs().max()) < 1e-5
    r = a.assign(pred, gt, gl, score, DEC_TRAIN_CFG)
    assert torch.equal(gt_inds[0].cpu(), r.gt_inds) and rel_err(ov[0].cpu(), r.max_overlaps) < 1e-5
    # the single-problem entry point with the reference signature
    gi, mo, lab = A.assign(pred.to(dev()), gt.to(dev()), gl.to(dev()), score.to(dev()), DEC_TRAIN_CFG)
    assert torch.equal(gi.cpu(), r.gt_inds) and torch.equal(lab.cpu(), r.labels)


@pytest.mark.parametrize('with_labels', [False, True])
def test_heuristic_assigner_matches_oracle(with_labels):
    import oracle.loss as ol
    from deepinteraction_b200 import loss as L
    g = torch.Generator().manual_seed(13)
    P, G = 120, 25
    pred = _random_boxes(P, g, 30.0)
    gt = torch.cat([_random_boxes(G - 6, g, 30.0), pred[:6] + 0.3 * torch.randn(6, 9, generator=g)], 0)
    gt[:, 3:6] = gt[:, 3:6].abs() + 0.5
    gt[3, :2] = gt[2, :2] + 0.01                       # two boxes compete for one prediction
    gl = torch.randint(0, 4, (G,), generator=g)
    ql = torch.randint(0, 4, (P,), generator=g) if with_labels else None
    ref = ol.HeuristicAssigner3D(dist_thre=20).assign(pred, gt, None, gl, ql)
    gi, ov, lab = L.HeuristicAssigner3D(dist_thre=20).assign(pred.to(dev()), gt.to(dev()), None, gl.to(dev()),
                                                              None if ql is None else ql.to(dev()))
    assert torch.equal(gi.cpu(), ref.gt_inds) and torch.equal(lab.cpu(), ref.labels.float())
    assert float((ov.cpu() - ref.max_overlaps).abs().max()) < 2e-6 and int((gi > 0).sum()) > 5


@pytest.mark.parametrize('P,Gs', [(200, (37, 0, 1)), (24, (24, 30, 5)), (300, (120, 299, 300)), (7, (3, 7, 12))])
def test_hungarian_matches_scipy(P, Gs):
    """Rectangular assignment problems of every orientation (fewer / as many / more ground-truth boxes than proposals, one,
    none), 2 layers per sample: the warp solver returns scipy.optimize.linear_sum_assignment's matching."""
    from scipy.optimize import linear_sum_assignment
    from deepinteraction_b200 import ops
    g = torch.Generator().manual_seed(P)
    B, L = len(Gs), 2
    Gmax = max(1, max(Gs))
    cost = torch.randn(B, L * P, Gmax, generator=g)
    iou = torch.rand(B, L * P, Gmax, generator=g)
    n = torch.tensor(Gs, dtype=torch.int32)
    gi = torch.empty(B, L * P, dtype=torch.int64, device=dev())
    mo = torch.empty(B, L * P, dtype=torch.float32, device=dev())
    cd, idv, nd = cost.to(dev()), iou.to(dev()), n.to(dev())          # keep the device copies alive across the call
    ops._call('di_hungarian_f32', ops._ptr(cd), ops._ptr(idv), ops._ptr(nd), B, L, P, Gmax, ops._ptr(gi), ops._ptr(mo),
              ops._stream())
    gi, mo = gi.cpu(), mo.cpu()
    for b in range(B):
        for l in range(L):
            c = cost[b, l * P:(l + 1) * P, :Gs[b]]
            want = torch.zeros(P, dtype=torch.int64)
            wov = torch.zeros(P)
            if Gs[b]:
                r, cc = linear_sum_assignment(c.double().numpy())
                want[torch.from_numpy(r)] = torch.from_numpy(cc) + 1
                wov[torch.from_numpy(r)] = iou[b, l * P:(l + 1) * P][torch.from_numpy(r), torch.from_numpy(cc)]
            got = gi[b, l * P:(l + 1) * P]
            if not torch.equal(got, want):       # equal total cost would also be a valid optimum; report the costs
                tot = lambda m: float(sum(c[i, int(m[i]) - 1] for i in range(P) if m[i] > 0))
                raise AssertionError((b, l, tot(got), tot(want)))
            assert torch.equal(mo[b, l * P:(l + 1) * P], wov)


@pytest.mark.parametrize('tag', ['decoder_loss', 'decoder_pp_loss'])
def test_loss_matches_reference_golden(tag):
    """get_targets and loss on the GPU, fed with the reference's own predictions, vs the reference's outputs."""
    from deepinteraction_b200 import loss as L
    gold = torch.load(os.path.join(G, tag + '.pt'), weights_only=False)
    m = _head(plusplus=tag == 'decoder_pp_loss')
    d = dev()
    preds = {k: v.to(d) for k, v in gold['preds'].items()}
    m.query_labels = gold['query_labels'].to(d)
    m.on_the_image_mask = [k.to(d) for k in gold['on_the_image_mask']]

    class Boxes:
        def __init__(self, t):
            self.tensor = t
    gt = [Boxes(b.to(d)) for b in gold['gt_boxes']]
    gl = [l.to(d) for l in gold['gt_labels']]
    tg = m.get_targets(gt, gl, [preds])
    t = gold['targets']
    assert torch.equal(tg[0].cpu(), t['labels']) and torch.equal(tg[1].cpu(), t['label_weights'])
    assert torch.equal(tg[3].cpu(), t['bbox_weights']) and int(tg[5]) == t['num_pos']
    assert rel_err(tg[2].cpu(), t['bbox_targets']) < 1e-6 and rel_err(tg[4].cpu(), t['ious']) < 1e-5
    assert abs(float(tg[6]) - t['matched_ious']) < 1e-6
    assert torch.equal(tg[7].cpu() > 0, t['heatmap'] > 0) and rel_err(tg[7].cpu(), t['heatmap']) < 1e-6
    assert torch.equal(tg[7].cpu() == 1, t['heatmap'] == 1)
    before = {k: v.clone() for k, v in preds.items()}
    losses = m.loss(gt, gl, [[preds]])
    assert set(losses) == set(gold['losses'])
    for k, ref in gold['losses'].items():
        assert abs(float(losses[k]) - float(ref)) <= 2e-5 * abs(float(ref)) + 1e-7, (k, float(losses[k]), float(ref))
    for k in preds:                                   # unlike the reference, the predictions are left untouched
        assert torch.equal(preds[k], before[k])


def test_loss_base_shape_matches_oracle():
    """200 proposals x 4 layers, 45 / 0-ish crowded ground-truth boxes, 180x180 heat maps, batch 2, vs oracle/loss.py."""
    import oracle.loss as ol
    import oracle.mmpi as om
    from tools import make_goldens as mg
    g = torch.Generator().manual_seed(21)
    B, P, L, K = 2, 200, 4, 10
    tc = dict(mg.DEC_TRAIN_CFG, grid_size=[1440, 1440, 40], voxel_size=[0.075, 0.075, 0.2])
    coder_cfg = dict(type='TransFusionBBoxCoder', pc_range=[-54.0, -54.0], voxel_size=[0.075, 0.075], out_size_factor=8,
                     post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.0, code_size=10)
    from deepinteraction_b200 import mmpi
    test_cfg = dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=8, pc_range=[-54.0, -54.0],
                    voxel_size=[0.075, 0.075], nms_type=None)
    kw = dict(num_views=6, out_size_factor_img=4, num_proposals=P, auxiliary=True, hidden_channel=128, num_classes=K, num_mmpi=L,
              num_heads=8, initialize_by_heatmap=True, nms_kernel_size=3, ffn_channel=256,
              common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)), bbox_coder=coder_cfg,
              test_cfg=test_cfg, train_cfg=tc, **mg.DEC_LOSSES)
    m = mmpi.DeepInteractionDecoder(**kw).to(dev()).eval()
    gts = [_random_boxes(45, g, 40.0), _random_boxes(3, g, 40.0)]
    gls = [torch.randint(0, K, (45,), generator=g), torch.randint(0, K, (3,), generator=g)]
    coder = om.TransFusionBBoxCoder(**{k: v for k, v in coder_cfg.items() if k != 'type'})
    LP = L * P
    # predictions scattered around the scene, a few of them close to ground-truth boxes
    centre = torch.rand(B, 2, LP, generator=g) * 170 + 5
    enc0 = coder.encode(gts[0])
    for l in range(L):
        centre[0, :, l * P:l * P + 45] = enc0[:, :2].T + 0.3 * torch.randn(2, 45, generator=g)
    preds = dict(center=centre, height=torch.randn(B, 1, LP, generator=g) - 1, dim=torch.randn(B, 3, LP, generator=g) * 0.3 + 0.8,
                 rot=torch.randn(B, 2, LP, generator=g), vel=torch.randn(B, 2, LP, generator=g),
                 heatmap=torch.randn(B, K, LP, generator=g) * 2, dense_heatmap=torch.randn(B, K, 180, 180, generator=g) * 2 - 2,
                 query_heatmap_score=torch.rand(B, K, P, generator=g))
    masks = [torch.rand(B, P, generator=g) > 0.2 for _ in range(2)]
    lh = ol.LossHead(K, P, L, coder, tc, **mg.DEC_LOSSES)
    lh.on_the_image_mask = masks
    ref = lh.loss([ol.LiDARBoxes(b) for b in gts], gls, [[{k: v.clone() for k, v in preds.items()}]])
    d = dev()
    m.on_the_image_mask = [k.to(d) for k in masks]
    out = m.loss([b.to(d) for b in gts], [l.to(d) for l in gls], [[{k: v.to(d) for k, v in preds.items()}]])
    t, rt = m._last_targets, ref['_targets']
    assert torch.equal(t['labels'].cpu(), rt['labels']) and torch.equal(t['label_weights'].cpu(), rt['label_weights'])
    assert torch.equal(t['bbox_weights'].cpu(), rt['bbox_weights']) and torch.equal(t['num_pos'].cpu(), rt['num_pos'])
    assert rel_err(t['heatmap'].cpu(), rt['heatmap']) < 1e-6 and rel_err(t['bbox_targets'].cpu(), rt['bbox_targets']) < 1e-6
    for k in ref:
        if k != '_targets':
            assert abs(float(out[k]) - float(ref[k])) <= 2e-5 * abs(float(ref[k])) + 1e-7, (k, float(out[k]), float(ref[k]))

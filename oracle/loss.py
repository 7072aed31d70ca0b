"""Oracle: the MMPI loss path: target assignment, Hungarian matching, gaussian heat-map
targets, focal / L1 / gaussian-focal losses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain PyTorch / numpy / scipy, CPU.

Part 1 restates the REFERENCE's own code (paths relative to projects/mmdet3d_plugin/):
* core/bbox/assigners/hungarian_assigner.py   BBox3DL1Cost :14-21, BBoxBEVL1Cost :24-37, IoU3DCost :40-47,
  HeuristicAssigner3D :50-91, HungarianAssigner3D :94-153
* models/dense_heads/deepinteraction_decoder.py   get_targets :315-353, get_targets_single :355-482, loss :484-547
  (deepinteractionplusplus_decoder.py :513-514: the ++ loss multiplies EVERY layer's weights by that layer's mask)

Part 2 restates the THIRD-PARTY pieces those functions call, which are not under /root/reference ("parity unpinned",
SURVEY.md 8(c)): mmdet 2.14 FocalLossCost, FocalLoss (sigmoid), L1Loss, GaussianFocalLoss, AssignResult, PseudoSampler;
mmdet3d 0.17.1 BboxOverlaps3D(coordinate='lidar') (rotated BEV intersection x height overlap / union of volumes),
gaussian_radius, draw_heatmap_gaussian, clip_sigmoid, LiDARInstance3DBoxes.gravity_center; mmcv ConfigDict.
tools/make_goldens.py (G8) runs the reference's UNMODIFIED files on top of part 2 and stores the outputs; part 1 is
checked against them (tests/test_oracle_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

try:
    from scipy.optimize import linear_sum_assignment
except ImportError:                                   # pragma: no cover
    linear_sum_assignment = None


# =====================================================================================================================
# Part 2: third party
# =====================================================================================================================
class ConfigDict(dict):
    """mmcv ConfigDict: dict with attribute access (nested dicts converted)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for key, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, ConfigDict):
                self[key] = ConfigDict(v)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


class AssignResult:
    def __init__(self, num_gts, gt_inds, max_overlaps, labels=None):
        self.num_gts, self.gt_inds, self.max_overlaps, self.labels = num_gts, gt_inds, max_overlaps, labels


class SamplingResult:
    def __init__(self, pos_inds, neg_inds, bboxes, gt_bboxes, assign_result):
        self.pos_inds, self.neg_inds = pos_inds, neg_inds
        self.pos_bboxes, self.neg_bboxes = bboxes[pos_inds], bboxes[neg_inds]
        self.num_gts = gt_bboxes.shape[0]
        self.pos_assigned_gt_inds = assign_result.gt_inds[pos_inds] - 1
        if gt_bboxes.numel() == 0:
            self.pos_gt_bboxes = torch.empty_like(gt_bboxes).view(-1, gt_bboxes.shape[-1] if gt_bboxes.dim() > 1 else 4)
        else:
            self.pos_gt_bboxes = gt_bboxes[self.pos_assigned_gt_inds, :]


class PseudoSampler:
    """mmdet PseudoSampler.sample: every assigned box is a positive, every gt_inds == 0 box a negative."""

    def sample(self, assign_result, bboxes, gt_bboxes, **kw):
        pos = torch.nonzero(assign_result.gt_inds > 0, as_tuple=False).squeeze(-1).unique()
        neg = torch.nonzero(assign_result.gt_inds == 0, as_tuple=False).squeeze(-1).unique()
        return SamplingResult(pos, neg, bboxes, gt_bboxes, assign_result)


class FocalLossCost:
    """mmdet 2.14 core/bbox/match_costs/match_cost.py."""

    def __init__(self, weight=1., alpha=0.25, gamma=2, eps=1e-12):
        self.weight, self.alpha, self.gamma, self.eps = weight, alpha, gamma, eps

    def __call__(self, cls_pred, gt_labels):
        p = cls_pred.sigmoid()
        neg = -(1 - p + self.eps).log() * (1 - self.alpha) * p.pow(self.gamma)
        pos = -(p + self.eps).log() * self.alpha * (1 - p).pow(self.gamma)
        return (pos[:, gt_labels] - neg[:, gt_labels]) * self.weight


def _rect_corners(b):
    """(n, 5) x, y, dx, dy, yaw -> (n, 4, 2) corners, mmdet3d 0.17 yaw convention (x' = x cos + y sin, y' = -x sin + y cos)."""
    x, y, dx, dy, r = (b[:, i] for i in range(5))
    c, s = np.cos(r), np.sin(r)
    out = np.empty((b.shape[0], 4, 2))
    for k, (sx, sy) in enumerate(((-1, -1), (1, -1), (1, 1), (-1, 1))):
        lx, ly = sx * dx / 2, sy * dy / 2
        out[:, k, 0] = lx * c + ly * s + x
        out[:, k, 1] = -lx * s + ly * c + y
    return out


def rotated_intersection_area(b1, b2):
    """Area of intersection of rotated rectangles b1 (n, 5) and b2 (m, 5) -> (n, m), float64.  Exact convex clipping:
    the corners of A are clipped by the four half planes of B (Sutherland-Hodgman), evaluated pair by pair in numpy."""
    b1, b2 = np.asarray(b1, np.float64), np.asarray(b2, np.float64)
    n, m = b1.shape[0], b2.shape[0]
    out = np.zeros((n, m))
    if n == 0 or m == 0:
        return out
    ca, cb = _rect_corners(b1), _rect_corners(b2)
    # orientation of B's corner loop decides the sign of "inside"
    for j in range(m):
        q = cb[j]
        e0, e1 = q[1] - q[0], q[2] - q[1]
        sgn = np.sign(e0[0] * e1[1] - e0[1] * e1[0]) or 1.0
        # cheap rejection on circumscribed circles
        ra = 0.5 * np.hypot(b1[:, 2], b1[:, 3])
        rb = 0.5 * np.hypot(b2[j, 2], b2[j, 3])
        near = np.hypot(b1[:, 0] - b2[j, 0], b1[:, 1] - b2[j, 1]) <= ra + rb
        for i in np.nonzero(near)[0]:
            poly = [tuple(p) for p in ca[i]]
            for e in range(4):
                p0, p1 = q[e], q[(e + 1) % 4]
                ex, ey = p1[0] - p0[0], p1[1] - p0[1]
                side = lambda pt: sgn * (ex * (pt[1] - p0[1]) - ey * (pt[0] - p0[0]))
                nxt = []
                for k in range(len(poly)):
                    a, b = poly[k], poly[(k + 1) % len(poly)]
                    sa, sb = side(a), side(b)
                    if sa >= 0:
                        nxt.append(a)
                    if (sa > 0 and sb < 0) or (sa < 0 and sb > 0):
                        t = sa / (sa - sb)
                        nxt.append((a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1])))
                poly = nxt
                if not poly:
                    break
            if len(poly) >= 3:
                xs, ys = np.array([p[0] for p in poly]), np.array([p[1] for p in poly])
                out[i, j] = 0.5 * abs(np.dot(xs, np.roll(ys, -1)) - np.dot(ys, np.roll(xs, -1)))
    return out


class BboxOverlaps3D:
    """mmdet3d 0.17.1 core/bbox/iou_calculators/iou3d_calculator.py, coordinate='lidar', mode='iou':
    boxes (x, y, z_bottom, dx, dy, dz, yaw, ...); 3-D IoU = BEV intersection * height overlap / union of volumes."""

    def __init__(self, coordinate='lidar'):
        assert coordinate == 'lidar'

    def __call__(self, b1, b2, mode='iou'):
        assert mode == 'iou' and b1.shape[-1] >= 7 and b2.shape[-1] >= 7
        rows, cols = b1.shape[0], b2.shape[0]
        if rows * cols == 0:
            return b1.new_zeros(rows, cols)
        top1, bot1 = (b1[:, 2] + b1[:, 5]).view(-1, 1), b1[:, 2].view(-1, 1)
        top2, bot2 = (b2[:, 2] + b2[:, 5]).view(1, -1), b2[:, 2].view(1, -1)
        oh = torch.clamp(torch.min(top1, top2) - torch.max(bot1, bot2), min=0)
        bev = torch.from_numpy(rotated_intersection_area(b1[:, [0, 1, 3, 4, 6]].detach().numpy(),
                                                         b2[:, [0, 1, 3, 4, 6]].detach().numpy())).to(b1.dtype)
        o3 = bev * oh
        v1, v2 = (b1[:, 3] * b1[:, 4] * b1[:, 5]).view(-1, 1), (b2[:, 3] * b2[:, 4] * b2[:, 5]).view(1, -1)
        return o3 / torch.clamp(v1 + v2 - o3, min=1e-8)


def _reduce(loss, weight, reduction, avg_factor):
    """mmdet weight_reduce_loss."""
    if weight is not None:
        loss = loss * weight
    if avg_factor is None:
        return loss.mean() if reduction == 'mean' else (loss.sum() if reduction == 'sum' else loss)
    assert reduction == 'mean'
    return loss.sum() / avg_factor


class FocalLoss:
    """mmdet FocalLoss(use_sigmoid=True): mmcv sigmoid_focal_loss (target == num_classes is background), weight per
    row, mean with avg_factor."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction='mean', loss_weight=1.0, **kw):
        assert use_sigmoid
        self.gamma, self.alpha, self.reduction, self.loss_weight = gamma, alpha, reduction, loss_weight

    def __call__(self, pred, target, weight=None, avg_factor=None):
        C = pred.shape[1]
        t = F.one_hot(target.clamp(max=C), C + 1)[:, :C].to(pred.dtype)
        p = pred.sigmoid()
        tiny = torch.finfo(torch.float32).tiny
        term_p = (1 - p).pow(self.gamma) * torch.log(p.clamp(min=tiny))
        term_n = p.pow(self.gamma) * torch.log((1 - p).clamp(min=tiny))
        loss = -t * self.alpha * term_p - (1 - t) * (1 - self.alpha) * term_n
        if weight is not None:
            weight = weight.view(-1, 1).to(pred.dtype)
        return self.loss_weight * _reduce(loss, weight, self.reduction, avg_factor)


class L1Loss:
    def __init__(self, reduction='mean', loss_weight=1.0, **kw):
        self.reduction, self.loss_weight = reduction, loss_weight

    def __call__(self, pred, target, weight=None, avg_factor=None):
        return self.loss_weight * _reduce((pred - target).abs(), weight, self.reduction, avg_factor)


class GaussianFocalLoss:
    def __init__(self, alpha=2.0, gamma=4.0, reduction='mean', loss_weight=1.0, **kw):
        self.alpha, self.gamma, self.reduction, self.loss_weight = alpha, gamma, reduction, loss_weight

    def __call__(self, pred, target, weight=None, avg_factor=None):
        eps = 1e-12
        pos_w, neg_w = target.eq(1), (1 - target).pow(self.gamma)
        loss = -(pred + eps).log() * (1 - pred).pow(self.alpha) * pos_w - (1 - pred + eps).log() * pred.pow(self.alpha) * neg_w
        return self.loss_weight * _reduce(loss, weight, self.reduction, avg_factor)


LOSSES = dict(FocalLoss=FocalLoss, L1Loss=L1Loss, GaussianFocalLoss=GaussianFocalLoss)


def build_loss(cfg):
    cfg = dict(cfg)
    return LOSSES[cfg.pop('type')](**cfg)


def clip_sigmoid(x, eps=1e-4):
    return torch.clamp(x.sigmoid_(), min=eps, max=1 - eps)


def gaussian_radius(det_size, min_overlap=0.5):
    height, width = det_size
    a1, b1, c1 = 1, height + width, width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + torch.sqrt(b1 ** 2 - 4 * a1 * c1)) / 2
    a2, b2, c2 = 4, 2 * (height + width), (1 - min_overlap) * width * height
    r2 = (b2 + torch.sqrt(b2 ** 2 - 4 * a2 * c2)) / 2
    a3, b3, c3 = 4 * min_overlap, -2 * min_overlap * (height + width), (min_overlap - 1) * width * height
    r3 = (b3 + torch.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def gaussian_2d(shape, sigma=1):
    m, n = [(ss - 1.) / 2. for ss in shape]
    y, x = np.ogrid[-m:m + 1, -n:n + 1]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def draw_heatmap_gaussian(heatmap, center, radius, k=1):
    diameter = 2 * radius + 1
    g = gaussian_2d((diameter, diameter), sigma=diameter / 6)
    x, y = int(center[0]), int(center[1])
    height, width = heatmap.shape[0:2]
    left, right = min(x, radius), min(width - x, radius + 1)
    top, bottom = min(y, radius), min(height - y, radius + 1)
    mh = heatmap[y - top:y + bottom, x - left:x + right]
    mg = torch.from_numpy(g[radius - top:radius + bottom, radius - left:radius + right]).to(heatmap.device, torch.float32)
    if min(mg.shape) > 0 and min(mh.shape) > 0:
        torch.max(mh, mg * k, out=mh)
    return heatmap


def multi_apply(func, *args, **kwargs):
    return tuple(map(list, zip(*map(lambda *a: func(*a, **kwargs), *args))))


class LiDARBoxes:
    """The three members of mmdet3d 0.17.1 LiDARInstance3DBoxes the loss path reads."""

    def __init__(self, tensor, box_dim=9):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def gravity_center(self):
        t = self.tensor
        return torch.cat([t[:, :2], (t[:, 2] + t[:, 5] * 0.5)[:, None]], 1)


# =====================================================================================================================
# Part 1: the reference's own code
# =====================================================================================================================
class BBox3DL1Cost:
    def __init__(self, weight):
        self.weight = weight

    def __call__(self, bboxes, gt_bboxes, train_cfg):
        return torch.cdist(bboxes, gt_bboxes, p=1) * self.weight


class BBoxBEVL1Cost:
    """hungarian_assigner.py:24-37: L1 distance of the BEV centres normalised by the point-cloud range."""

    def __init__(self, weight):
        self.weight = weight

    def __call__(self, bboxes, gt_bboxes, train_cfg):
        r = train_cfg['point_cloud_range']
        start = bboxes.new_tensor(r[0:2])
        size = bboxes.new_tensor(r[3:5]) - start
        return torch.cdist((bboxes[:, :2] - start) / size, (gt_bboxes[:, :2] - start) / size, p=1) * self.weight


class IoU3DCost:
    def __init__(self, weight):
        self.weight = weight

    def __call__(self, iou):
        return -iou * self.weight


MATCH_COSTS = dict(FocalLossCost=FocalLossCost, BBox3DL1Cost=BBox3DL1Cost, BBoxBEVL1Cost=BBoxBEVL1Cost, IoU3DCost=IoU3DCost)


def _build(table, cfg):
    cfg = dict(cfg)
    return table[cfg.pop('type')](**cfg)


class HungarianAssigner3D:
    """hungarian_assigner.py:94-153."""

    def __init__(self, cls_cost, reg_cost, iou_cost, iou_calculator=dict(type='BboxOverlaps3D', coordinate='lidar')):
        self.cls_cost, self.reg_cost, self.iou_cost = (_build(MATCH_COSTS, c) for c in (cls_cost, reg_cost, iou_cost))
        self.iou_calculator = _build(dict(BboxOverlaps3D=BboxOverlaps3D), iou_calculator)

    def assign(self, bboxes, gt_bboxes, gt_labels, cls_pred, train_cfg):
        num_gts, num_bboxes = gt_bboxes.size(0), bboxes.size(0)
        gt_inds = bboxes.new_full((num_bboxes,), -1, dtype=torch.long)
        labels = bboxes.new_full((num_bboxes,), -1, dtype=torch.long)
        if num_gts == 0 or num_bboxes == 0:
            if num_gts == 0:
                gt_inds[:] = 0
            return AssignResult(num_gts, gt_inds, None, labels=labels)
        iou = self.iou_calculator(bboxes, gt_bboxes)
        cost = self.cls_cost(cls_pred[0].T, gt_labels) + self.reg_cost(bboxes, gt_bboxes, train_cfg) + self.iou_cost(iou)
        rows, cols = linear_sum_assignment(cost.detach().cpu())
        rows, cols = torch.from_numpy(rows), torch.from_numpy(cols)
        gt_inds[:] = 0
        gt_inds[rows] = cols + 1
        labels[rows] = gt_labels[cols]
        max_overlaps = torch.zeros_like(iou.max(1).values)
        max_overlaps[rows] = iou[rows, cols]
        return AssignResult(num_gts, gt_inds, max_overlaps, labels=labels)


class HeuristicAssigner3D:
    """hungarian_assigner.py:50-91: every ground-truth box takes its nearest prediction (BEV distance, optional
    same-class constraint); a prediction claimed twice keeps the nearer box."""

    def __init__(self, dist_thre=100, iou_calculator=dict(type='BboxOverlaps3D')):
        self.dist_thre = dist_thre
        self.iou_calculator = BboxOverlaps3D()

    def assign(self, bboxes, gt_bboxes, gt_bboxes_ignore=None, gt_labels=None, query_labels=None):
        num_gts, num_bboxes = len(gt_bboxes), len(bboxes)
        dist = torch.norm(bboxes[:, 0:2][None] - gt_bboxes[:, 0:2][:, None], dim=-1)
        if query_labels is not None:
            dist = dist + (query_labels[None] != gt_labels[:, None]) * self.dist_thre
        nearest = dist.min(1).indices
        inds = torch.zeros(num_bboxes)
        vals = torch.ones(num_bboxes) * 10000
        labels = torch.ones(num_bboxes) * -1
        for g in range(num_gts):
            p = nearest[g]
            if dist[g, p] <= self.dist_thre and dist[g, p] < vals[p]:
                vals[p], inds[p], labels[p] = dist[g, p], g + 1, gt_labels[g]
        overlaps = torch.zeros(num_bboxes)
        hit = torch.where(inds > 0)
        overlaps[hit] = self.iou_calculator(gt_bboxes[inds[hit].long() - 1], bboxes[hit]).diag()
        return AssignResult(num_gts, inds.long(), overlaps, labels=labels)


ASSIGNERS = dict(HungarianAssigner3D=HungarianAssigner3D, HeuristicAssigner3D=HeuristicAssigner3D)


class LossHead:
    """Loss-side state and methods of DeepInteractionDecoder / DeepInteractionPlusPlusDecoder
    (deepinteraction_decoder.py:70-78, :186-199, :315-547).  `plusplus` selects the ++ mask rule (:513-514 of
    deepinteractionplusplus_decoder.py).  query_labels / on_the_image_mask are the forward's side outputs."""

    def __init__(self, num_classes, num_proposals, num_mmpi, bbox_coder, train_cfg, loss_cls, loss_bbox, loss_heatmap,
                 auxiliary=True, plusplus=False):
        self.num_classes, self.num_proposals, self.num_mmpi = num_classes, num_proposals, num_mmpi
        self.bbox_coder, self.auxiliary, self.plusplus = bbox_coder, auxiliary, plusplus
        self.train_cfg = ConfigDict(train_cfg)
        self.loss_cls, self.loss_bbox, self.loss_heatmap = build_loss(loss_cls), build_loss(loss_bbox), build_loss(loss_heatmap)
        self.bbox_sampler = PseudoSampler()
        self.bbox_assigner = _build(ASSIGNERS, self.train_cfg.assigner)
        self.query_labels, self.on_the_image_mask = None, None

    def get_targets(self, gt_bboxes_3d, gt_labels_3d, preds_dict):
        """:315-353."""
        per = []
        for b in range(len(gt_bboxes_3d)):
            per.append({k: v[b:b + 1] for k, v in preds_dict[0].items()})
        res = multi_apply(self.get_targets_single, gt_bboxes_3d, gt_labels_3d, per, np.arange(len(gt_labels_3d)))
        cat = lambda i: torch.cat(res[i], 0)
        return cat(0), cat(1), cat(2), cat(3), cat(4), np.sum(res[5]), np.mean(res[6]), cat(7)

    def get_targets_single(self, gt_bboxes_3d, gt_labels_3d, preds_dict, batch_idx):
        """:355-482."""
        P = preds_dict['center'].shape[-1]
        score = preds_dict['heatmap'].detach().clone()
        dec = self.bbox_coder.decode(score, preds_dict['rot'].detach().clone(), preds_dict['dim'].detach().clone(),
                                     preds_dict['center'].detach().clone(), preds_dict['height'].detach().clone(),
                                     preds_dict['vel'].detach().clone() if 'vel' in preds_dict else None)
        boxes = dec[0]['bboxes']
        gt = gt_bboxes_3d.tensor
        n_layer = self.num_mmpi if self.auxiliary else 1
        results = []
        for l in range(n_layer):
            sl = slice(self.num_proposals * l, self.num_proposals * (l + 1))
            if self.train_cfg.assigner.type == 'HungarianAssigner3D':
                results.append(self.bbox_assigner.assign(boxes[sl], gt, gt_labels_3d, score[..., sl], self.train_cfg))
            elif self.train_cfg.assigner.type == 'HeuristicAssigner':
                results.append(self.bbox_assigner.assign(boxes[sl], gt, None, gt_labels_3d, self.query_labels[batch_idx]))
            else:
                raise NotImplementedError
        ens = AssignResult(sum(r.num_gts for r in results), torch.cat([r.gt_inds for r in results]),
                           torch.cat([r.max_overlaps for r in results]), torch.cat([r.labels for r in results]))
        samp = self.bbox_sampler.sample(ens, boxes, gt)
        pos, neg = samp.pos_inds, samp.neg_inds
        assert len(pos) + len(neg) == P
        code = self.bbox_coder.code_size
        bbox_targets, bbox_weights = torch.zeros(P, code), torch.zeros(P, code)
        ious = torch.clamp(ens.max_overlaps, min=0.0, max=1.0)
        labels = boxes.new_zeros(P, dtype=torch.long)
        label_weights = boxes.new_zeros(P, dtype=torch.long)
        if gt_labels_3d is not None:
            labels += self.num_classes
        if len(pos) > 0:
            bbox_targets[pos] = self.bbox_coder.encode(samp.pos_gt_bboxes)
            bbox_weights[pos] = 1.0
            labels[pos] = 1 if gt_labels_3d is None else gt_labels_3d[samp.pos_assigned_gt_inds]
            label_weights[pos] = 1.0 if self.train_cfg.pos_weight <= 0 else self.train_cfg.pos_weight
        if len(neg) > 0:
            label_weights[neg] = 1.0
        # dense heat-map targets (:443-476)
        tc = self.train_cfg
        g = torch.cat([gt_bboxes_3d.gravity_center, gt[:, 3:]], 1)
        grid, rng, vox = torch.tensor(tc['grid_size']), torch.tensor(tc['point_cloud_range']), torch.tensor(tc['voxel_size'])
        fmap = grid[:2] // tc['out_size_factor']
        heatmap = g.new_zeros(self.num_classes, int(fmap[1]), int(fmap[0]))
        for i in range(len(g)):
            width = g[i][3] / vox[0] / tc['out_size_factor']
            length = g[i][4] / vox[1] / tc['out_size_factor']
            if width > 0 and length > 0:
                radius = gaussian_radius((length, width), min_overlap=tc['gaussian_overlap'])
                radius = max(tc['min_radius'], int(radius))
                cx = (g[i][0] - rng[0]) / vox[0] / tc['out_size_factor']
                cy = (g[i][1] - rng[1]) / vox[1] / tc['out_size_factor']
                ci = torch.tensor([cx, cy], dtype=torch.float32).to(torch.int32)
                draw_heatmap_gaussian(heatmap[gt_labels_3d[i]], ci, radius)
        mean_iou = ious[pos].sum() / max(len(pos), 1)
        return (labels[None], label_weights[None], bbox_targets[None], bbox_weights[None], ious[None], int(pos.shape[0]),
                float(mean_iou), heatmap[None])

    def loss(self, gt_bboxes_3d, gt_labels_3d, preds_dicts):
        """:484-547 -> dict of loss tensors (+ the targets under '_targets' for the tests)."""
        labels, label_weights, bbox_targets, bbox_weights, ious, _, matched_ious, heatmap = \
            self.get_targets(gt_bboxes_3d, gt_labels_3d, preds_dicts[0])
        P = self.num_proposals
        num_pos = []
        for l in range(self.num_mmpi):
            sl = slice(l * P, (l + 1) * P)
            mask = None
            if self.plusplus:
                mask = self.on_the_image_mask[l]
            elif l % 2 == 0:
                mask = self.on_the_image_mask[l // 2]
            if mask is not None:
                label_weights[..., sl] = label_weights[..., sl] * mask
                bbox_weights[:, sl, :] = bbox_weights[:, sl, :] * mask[:, :, None]
            num_pos.append(bbox_weights.max(-1).values[..., sl].sum())
        pd = preds_dicts[0][0]
        out = dict()
        out['loss_heatmap'] = self.loss_heatmap(clip_sigmoid(pd['dense_heatmap'].clone()), heatmap,
                                                avg_factor=max(heatmap.eq(1).float().sum().item(), 1))
        cw = self.train_cfg.get('code_weights', None)
        for l in range(self.num_mmpi):
            sl = slice(l * P, (l + 1) * P)
            cls_score = pd['heatmap'][..., sl].permute(0, 2, 1).reshape(-1, self.num_classes)
            out[f'layer_{l}_loss_cls'] = self.loss_cls(cls_score, labels[..., sl].reshape(-1), label_weights[..., sl].reshape(-1),
                                                       avg_factor=max(num_pos[l], 1))
            parts = [pd['center'][..., sl], pd['height'][..., sl], pd['dim'][..., sl], pd['rot'][..., sl]]
            if 'vel' in pd:
                parts.append(pd['vel'][..., sl])
            preds = torch.cat(parts, 1).permute(0, 2, 1)
            reg_w = bbox_weights[:, sl, :] * bbox_weights.new_tensor(cw)
            out[f'layer_{l}_loss_bbox'] = self.loss_bbox(preds, bbox_targets[:, sl, :], reg_w, avg_factor=max(num_pos[l], 1))
        out['matched_ious'] = pd['heatmap'].new_tensor(matched_ious)
        out['_targets'] = dict(labels=labels, label_weights=label_weights, bbox_targets=bbox_targets, bbox_weights=bbox_weights,
                               ious=ious, heatmap=heatmap, num_pos=torch.stack(num_pos))
        return out

"""Oracle restatements of third-party geometry used by the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  None of these live under the
reference tree; they are pinned only by the versions in the reference's
install.md (mmdet3d 0.17.1, detectron2 for torch 1.9).  The semantics below
follow SURVEY.md Appendix C.
"""
import numpy as np
import torch


def apply_3d_transformation(pcd, img_meta, reverse=False):
    """mmdet3d 0.17.1 ``fusion_layers/coord_transform.py::apply_3d_transformation``
    for LiDAR coordinates (call sites: reference
    models/utils/encoder_utils.py:156,189,280; models/utils/decoder_utils.py:692).

    Points are row vectors; ops are applied in ``transformation_3d_flow`` order
    (reversed, with inverse parameters, when ``reverse``)."""
    dtype, dev = pcd.dtype, pcd.device
    rot = (torch.as_tensor(np.asarray(img_meta['pcd_rotation']), dtype=dtype, device=dev)
           if 'pcd_rotation' in img_meta else torch.eye(3, dtype=dtype, device=dev))
    scale = img_meta.get('pcd_scale_factor', 1.0)
    trans = (torch.as_tensor(np.asarray(img_meta['pcd_trans']), dtype=dtype, device=dev)
             if 'pcd_trans' in img_meta else torch.zeros(3, dtype=dtype, device=dev))
    hflip = img_meta.get('pcd_horizontal_flip', False)
    vflip = img_meta.get('pcd_vertical_flip', False)
    flow = list(img_meta.get('transformation_3d_flow', []))
    p = pcd.clone()
    if reverse:
        scale = 1.0 / scale
        trans = -trans
        rot = rot.inverse()
        flow = flow[::-1]
    for op in flow:
        if op == 'T':
            p[:, :3] = p[:, :3] + trans
        elif op == 'S':
            p[:, :3] = p[:, :3] * scale
        elif op == 'R':
            p[:, :3] = p[:, :3] @ rot
        elif op == 'HF':
            if hflip:
                p[:, 1] = -p[:, 1]
        elif op == 'VF':
            if vflip:
                p[:, 0] = -p[:, 0]
        else:
            raise AssertionError(op)
    return p


def affine_of_transformation(img_meta, reverse=False):
    """4x4 float64 matrix A (column-vector convention, ``p' = A @ [p;1]``) equal to
    ``apply_3d_transformation``; used only to check the host-side folding of the
    product path."""
    eye = torch.eye(3, dtype=torch.float64)
    pts = torch.cat([torch.zeros(1, 3, dtype=torch.float64), eye], 0)
    meta = dict(img_meta)
    out = apply_3d_transformation(pts, meta, reverse=reverse)
    A = torch.eye(4, dtype=torch.float64)
    A[:3, 3] = out[0]
    A[:3, :3] = (out[1:] - out[0]).T
    return A


_CORNER_ORDER = [0, 1, 3, 2, 4, 5, 7, 6]


def lidar_box_corners(boxes):
    """mmdet3d 0.17.1 ``LiDARInstance3DBoxes(boxes[:, :7]).corners`` -> (N, 8, 3).

    Box = (x, y, z_bottom, dx, dy, dz, yaw); relative origin (0.5, 0.5, 0);
    rotation about z is ``corners @ [[c,-s,0],[s,c,0],[0,0,1]]`` (the pre-1.0
    convention).  Call sites: reference decoder_utils.py:690-691, 808."""
    dims = boxes[:, 3:6]
    idx = np.stack(np.unravel_index(np.arange(8), [2] * 3), axis=1)[_CORNER_ORDER]
    norm = torch.as_tensor(idx, dtype=boxes.dtype, device=boxes.device)
    norm = norm - boxes.new_tensor([0.5, 0.5, 0.0])
    corners = dims.view(-1, 1, 3) * norm.view(1, 8, 3)
    yaw = boxes[:, 6]
    c, s = torch.cos(yaw), torch.sin(yaw)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    rot_t = torch.stack([torch.stack([c, -s, z]), torch.stack([s, c, z]), torch.stack([z, z, o])])
    corners = torch.einsum('aij,jka->aik', corners, rot_t)
    return corners + boxes[:, :3].view(-1, 1, 3)


def _bilinear_roi(feat, y, x):
    """detectron2/torchvision ROIAlign ``bilinear_interpolate`` on feat (C,H,W) at
    points y,x (any shape) -> (C, *shape)."""
    C, H, W = feat.shape
    dead = (y < -1.0) | (y > H) | (x < -1.0) | (x > W)
    y = y.clamp(min=0.0)
    x = x.clamp(min=0.0)
    y0 = y.floor().long()
    x0 = x.floor().long()
    ytop = y0 >= H - 1
    xtop = x0 >= W - 1
    y0 = torch.where(ytop, torch.full_like(y0, H - 1), y0)
    x0 = torch.where(xtop, torch.full_like(x0, W - 1), x0)
    y1 = torch.where(ytop, y0, y0 + 1)
    x1 = torch.where(xtop, x0, x0 + 1)
    y = torch.where(ytop, y0.to(y.dtype), y)
    x = torch.where(xtop, x0.to(x.dtype), x)
    ly, lx = y - y0, x - x0
    hy, hx = 1.0 - ly, 1.0 - lx
    flat = feat.reshape(C, -1)

    def g(yy, xx):
        return flat[:, (yy * W + xx).reshape(-1)].reshape(C, *y.shape)
    val = (hy * hx) * g(y0, x0) + (hy * lx) * g(y0, x1) + (ly * hx) * g(y1, x0) + (ly * lx) * g(y1, x1)
    return torch.where(dead, torch.zeros_like(val), val)


def roi_align(feat, boxes, out_size=7, spatial_scale=1.0, sampling_ratio=2):
    """detectron2 ``ROIPooler(output_size=7, scales=[s], sampling_ratio=2,
    pooler_type='ROIAlignV2')`` on ONE feature map (C,H,W) with boxes (R,4)
    (x0,y0,x1,y1) -> (R, C, 7, 7).  aligned=True: continuous coordinate
    ``x*s - 0.5``, no minimum RoI size.  Call sites: reference
    decoder_utils.py:641-646,739-741 (scale 1/4) and :769-774,822-823 (scale 1)."""
    R = boxes.shape[0]
    C = feat.shape[0]
    if R == 0:
        return feat.new_zeros(0, C, out_size, out_size)
    x0 = boxes[:, 0] * spatial_scale - 0.5
    y0 = boxes[:, 1] * spatial_scale - 0.5
    x1 = boxes[:, 2] * spatial_scale - 0.5
    y1 = boxes[:, 3] * spatial_scale - 0.5
    bw = (x1 - x0) / out_size
    bh = (y1 - y0) / out_size
    g = sampling_ratio
    p = torch.arange(out_size, dtype=feat.dtype, device=feat.device)
    i = torch.arange(g, dtype=feat.dtype, device=feat.device)
    # sample coordinate for bin p, sub-sample i:  start + p*bin + (i+.5)*bin/g
    ys = y0[:, None, None] + p[None, :, None] * bh[:, None, None] + (i[None, None, :] + 0.5) * bh[:, None, None] / g
    xs = x0[:, None, None] + p[None, :, None] * bw[:, None, None] + (i[None, None, :] + 0.5) * bw[:, None, None] / g
    Y = ys[:, :, :, None, None].expand(R, out_size, g, out_size, g)
    X = xs[:, None, None, :, :].expand(R, out_size, g, out_size, g)
    vals = _bilinear_roi(feat, Y, X)          # (C,R,7,g,7,g)
    out = vals.sum(dim=(3, 5)) / float(g * g)  # (C,R,7,7)
    return out.permute(1, 0, 2, 3).contiguous()

"""Host-side geometry folding for the MMRI/MMPI kernels (tiny 4x4 algebra, float64).

The reference undoes / re-applies the recorded 3-D augmentation with mmdet3d 0.17.1
``apply_3d_transformation`` on every point (reference models/utils/encoder_utils.py:156,189,280;
models/utils/decoder_utils.py:692).  That transform is affine, so it is folded ONCE per sample
into the camera matrices that the kernels consume:

    proj[b,v] = lidar2img[b,v] @ A_reverse(b)          (3x4 rows, fp32)
    i2l[b,v]  = A_forward(b) @ inverse(lidar2img[b,v])  (3x4 rows, fp32)
"""
import numpy as np
import torch


def aug_affine(img_meta, reverse):
    """4x4 float64 matrix of apply_3d_transformation(., 'LIDAR', img_meta, reverse) in column-vector
    convention.  Points are row vectors in mmdet3d (p @ rot); flow ops: R,S,T,HF (y -> -y), VF (x -> -x)."""
    rot = np.asarray(img_meta['pcd_rotation'], np.float64) if 'pcd_rotation' in img_meta else np.eye(3)
    scale = float(img_meta.get('pcd_scale_factor', 1.0))
    trans = np.asarray(img_meta['pcd_trans'], np.float64) if 'pcd_trans' in img_meta else np.zeros(3)
    hflip = bool(img_meta.get('pcd_horizontal_flip', False))
    vflip = bool(img_meta.get('pcd_vertical_flip', False))
    flow = list(img_meta.get('transformation_3d_flow', []))
    if reverse:
        rot, scale, trans, flow = np.linalg.inv(rot), 1.0 / scale, -trans, flow[::-1]
    A = np.eye(4)
    for op in flow:
        S = np.eye(4)
        if op == 'R':
            S[:3, :3] = rot.T            # p_row @ rot  ==  rot.T @ p_col
        elif op == 'S':
            S[:3, :3] *= scale
        elif op == 'T':
            S[:3, 3] = trans
        elif op == 'HF':
            if hflip:
                S[1, 1] = -1.0
        elif op == 'VF':
            if vflip:
                S[0, 0] = -1.0
        else:
            raise ValueError(f'unknown transformation_3d_flow op {op!r}')
        A = S @ A
    return A


def lidar2img_array(img_metas):
    return np.asarray([np.asarray(m['lidar2img'], np.float32) for m in img_metas], np.float32)   # (B,V,4,4)


def camera_rows_host(img_metas):
    """-> proj (B,V,12) fp32, i2l (B*V,12) fp32 as CPU tensors."""
    l2i = lidar2img_array(img_metas)
    B, V = l2i.shape[:2]
    inv = torch.inverse(torch.from_numpy(l2i)).numpy().astype(np.float64)    # fp32 inverse, as the reference (:149)
    proj = np.empty((B, V, 12), np.float32)
    i2l = np.empty((B, V, 12), np.float32)
    for b, meta in enumerate(img_metas):
        Ar, Af = aug_affine(meta, True), aug_affine(meta, False)
        for v in range(V):
            proj[b, v] = (l2i[b, v].astype(np.float64) @ Ar)[:3].reshape(12)
            i2l[b, v] = (Af @ inv[b, v])[:3].reshape(12)
    return torch.from_numpy(proj), torch.from_numpy(i2l.reshape(B * V, 12))


def camera_rows(img_metas, device):
    """-> proj (B,V,12) fp32, i2l (B*V,12) fp32 on `device`."""
    proj, i2l = camera_rows_host(img_metas)
    return proj.to(device, non_blocking=True), i2l.to(device, non_blocking=True)


def input_hw(img_metas):
    hw = tuple(int(v) for v in img_metas[0]['input_shape'][:2])
    for m in img_metas:
        assert tuple(int(v) for v in m['input_shape'][:2]) == hw, 'all samples must share input_shape'
    return hw

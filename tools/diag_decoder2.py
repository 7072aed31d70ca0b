"""Every dense call of one base-shape decoder forward: tensor-core result vs FFMA result on the same inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
from deepinteraction_b200 import ops, synth, mmpi, fold
from test_gpu_decoder import _build

dev = torch.device('cuda:0')
test_cfg = dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=8, pc_range=[-54.0, -54.0],
                voxel_size=[0.075, 0.075], nms_type=None)
coder = dict(type='TransFusionBBoxCoder', pc_range=[-54.0, -54.0], voxel_size=[0.075, 0.075], out_size_factor=8,
             post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.0, code_size=10)
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7
o, m = _build(seed, 6, 200, test_cfg, coder)
gen = torch.Generator().manual_seed(seed)
rig = synth.camera_rig(6, (448, 800))
metas = [dict(lidar2img=[r.astype(np.float32) for r in rig], input_shape=(448, 800), img_shape=[(448, 800, 3)] * 6)]
pts_in = [torch.randn(1, 128, 180, 180, generator=gen), torch.randn(1, 128, 180, 180, generator=gen)]
img_in = torch.randn(6, 128, 112, 200, generator=gen)

real_linear = ops.linear
count = [0]


def checked_linear(srcs, W, *a, **kw):
    out = real_linear(srcs, W, *a, **kw)
    if isinstance(W, fold.Weight) and kw.get('splits', 1) == 1:
        ops.USE_TC[0] = False
        try:
            ref = real_linear([s.clone() for s in srcs], W, *a, **{k: v for k, v in kw.items() if k != 'out'})
        finally:
            ops.USE_TC[0] = True
        torch.cuda.synchronize()
        d = float((out - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
        bad = int(((out - ref).abs() > 1e-3 * ref.abs().max()).sum())
        count[0] += 1
        flag = '  <<<<<< MISMATCH' if d > 1e-4 else ''
        print('linear #%d M=%d N=%d K=%s act=%s res=%s: rel diff %.2e bad=%d nan=%d%s' % (
            count[0], out.shape[0], out.shape[1], [s.shape[1] for s in srcs], a[1] if len(a) > 1 else kw.get('act', 0),
            kw.get('res') is not None, d, bad, int(torch.isnan(out).sum()), flag))
        if flag:
            idx = ((out - ref).abs() > 1e-3 * ref.abs().max()).nonzero()
            print('   first bad idx', idx[:6].tolist(), ' last', idx[-3:].tolist())
    return out


ops.linear = checked_linear
mmpi.ops.linear = checked_linear
out = m([p.to(dev) for p in pts_in], img_in.to(dev), metas)[0][0]

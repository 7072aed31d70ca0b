"""Base-shape decoder, stage tensors of the tensor-core run vs the FFMA run (no syncs between kernels)."""
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
pts_in = [torch.randn(1, 128, 180, 180, generator=gen).to(dev), torch.randn(1, 128, 180, 180, generator=gen).to(dev)]
img_in = torch.randn(6, 128, 112, 200, generator=gen).to(dev)
nh = lambda t: t.permute(0, 2, 3, 1).contiguous()
a, b, c = nh(pts_in[0]), nh(pts_in[1]), nh(img_in)


def run(use_tc):
    ops.USE_TC[0] = use_tc
    dbg = {}
    try:
        r = m.forward_rows(a, b, c, metas, dbg)
    finally:
        ops.USE_TC[0] = True
    torch.cuda.synchronize()
    dbg['preds'] = r['preds']
    return dbg


for rep in range(2):
    ref = run(False)
    tc = run(True)
    tc2 = run(True)
    for k in ('heat', 'query_feat0', 'query_pos0', 'query_feat1', 'first_res', 'layer_query', 'preds'):
        va, vb, vc = ref[k], tc[k], tc2[k]
        if not isinstance(va, list):
            va, vb, vc = [va], [vb], [vc]
        for i, (x, y, z) in enumerate(zip(va, vb, vc)):
            x, y, z = x.float(), y.float(), z.float()
            print(rep, k, i, 'tc vs ffma %.2e' % float((x - y).abs().max() / x.abs().max().clamp_min(1e-20)),
                  ' tc vs tc(rerun) %.2e' % float((z - y).abs().max() / x.abs().max().clamp_min(1e-20)))
    print(rep, 'top equal', bool(torch.equal(ref['top'], tc['top'])))

"""Base-shape decoder vs oracle under the FFMA / 3xTF32 / bf16-split dense paths (error per output tensor)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch
from deepinteraction_b200 import ops, synth, mmpi
from test_gpu_decoder import _build
from conftest import rel_err

dev = torch.device('cuda:0')
test_cfg = dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=8, pc_range=[-54.0, -54.0],
                voxel_size=[0.075, 0.075], nms_type=None)
coder = dict(type='TransFusionBBoxCoder', pc_range=[-54.0, -54.0], voxel_size=[0.075, 0.075], out_size_factor=8,
             post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.0, code_size=10)
for seed in (1601, 7):
    o, m = _build(seed, 6, 200, test_cfg, coder)
    gen = torch.Generator().manual_seed(seed)
    rig = synth.camera_rig(6, (448, 800))
    metas = [dict(lidar2img=[r.astype(np.float32) for r in rig], input_shape=(448, 800), img_shape=[(448, 800, 3)] * 6)]
    pts_in = [torch.randn(1, 128, 180, 180, generator=gen), torch.randn(1, 128, 180, 180, generator=gen)]
    img_in = torch.randn(6, 128, 112, 200, generator=gen)
    with torch.no_grad():
        ref = o(pts_in, img_in, metas)[0][0]
        ref64 = None
    outs = {}
    for name, use_tc, force_bf in (('ffma', False, False), ('tc-3xtf32', True, False), ('tc-bf16', True, True)):
        ops.USE_TC[0] = use_tc
        impl = mmpi.DeepInteractionDecoder._schedule
        if force_bf:
            mmpi.DeepInteractionDecoder._schedule = mmpi.DeepInteractionDecoder._schedule_impl
        try:
            out = m([p.to(dev) for p in pts_in], img_in.to(dev), metas)[0][0]
        finally:
            mmpi.DeepInteractionDecoder._schedule = impl
            ops.USE_TC[0] = True
        outs[name] = {k: v.cpu() for k, v in out.items()}
        print(seed, name, {k: '%.1e' % rel_err(outs[name][k], ref[k]) for k in ref},
              'labels_equal', bool(torch.equal(m.query_labels.cpu(), o.query_labels)))
    print(seed, 'tc-3xtf32 vs ffma', {k: '%.1e' % rel_err(outs['tc-3xtf32'][k], outs['ffma'][k]) for k in ref})

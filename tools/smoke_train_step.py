#!/usr/bin/env python
"""Smallest end-to-end run of train.encoder_train_step with every option on (attention dropout, gradient reporting into
GradBuckets): checks that the glue runs and that all outputs are finite.  Parity is the job of tests/test_gpu_backward.py."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from deepinteraction_b200 import mmri, synth, train
    from deepinteraction_b200.shard import GradBuckets
    from tools.make_goldens import small_frame
    d = torch.device('cuda:0')
    torch.manual_seed(3)
    enc = mmri.DeepInteractionEncoder(2, 64, 64, 128)
    synth.randomize_norm_stats(enc, 3)
    enc = enc.to(d).train()
    fr = small_frame(3, aug=True, views=2, c_img=64, c_pts=64, bev=36, batch=1)
    pm = {k: (v.to(d) if torch.is_tensor(v) else [p.to(d) for p in v]) for k, v in fr['pts_metas'].items()}
    res = {}
    for seed in (None, 7):
        buckets, order = GradBuckets(bucket_bytes=1 << 16), []
        r = train.encoder_train_step(enc, fr['img_feats'].to(d), fr['pts_feats'].to(d), fr['img_metas'], pm,
                                     lambda a, b, c: (torch.ones_like(a), None, torch.ones_like(c) * 0.5), dropout_seed=seed,
                                     on_grad=lambda n, t: (order.append(n), buckets.add(t)))
        buckets.finish()
        torch.cuda.synchronize()
        fin = all(bool(torch.isfinite(t).all()) for t in list(r['grads'].values()) + list(r['outputs']) + [r['d_img_feats'], r['d_pts_feats']])
        res[seed] = r
        print('dropout_seed=%s: %d gradients (first reported: %s), %d buckets, finite=%s' % (seed, len(r['grads']), order[0], buckets.launched, fin))
        assert fin and len(r['grads']) == len(list(enc.parameters()))
    a, b = res[None]['outputs'][2], res[7]['outputs'][2]
    print('BEV output changes with dropout: %.3e (relative)' % float((a - b).abs().max() / a.abs().max()))


if __name__ == '__main__':
    main()

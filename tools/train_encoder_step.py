#!/usr/bin/env python
"""Partial BASELINE config 3: forward + backward of the base MMRI ENCODER at the base shapes (6 x (256,112,200) camera maps,
(512,180,180) BEV map, ~250k points), synthetic loss = sum of squared differences to fixed random targets on the three
encoder outputs, parameter gradients all-reduced over the ranks with shard.GradBuckets (NCCL).
--bn eval: BatchNorm folded (backward.encoder_backward; gradients of the folded weights).
--bn train: BatchNorm with batch statistics (train.encoder_train_step; gradients of the module's own parameters), followed by
a torch.optim.AdamW step on the encoder's parameters (the reference's optimiser, configs/nuscenes/Fusion_0075_refactor.py).
--dropout adds the MMRI_I2P attention dropout.  The decoder's backward is not built, so this is NOT the config-3 metric;
it measures what exists: the encoder's training-side kernels and the path's one collective.

    python tools/train_encoder_step.py [--steps 5]                                    # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/train_encoder_step.py
Prints one JSON line: frames/s of forward + backward + all-reduce (bs = 1 per GPU), ms per step, gradient bytes.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def flat_grads(r):
    out = []
    for name in ('img', 'pts'):
        out += list(r['shared_conv'][name])
    for lg in r['layers']:
        out += list(lg['i2p'])
        for blk in ('p_iml', 'p2i', 'i_iml'):
            for nm in ('q1', 'q2', 'k1', 'k2', 'v'):
                out += list(lg[blk][nm])
        out += list(lg['p_fuse']) + list(lg['i_fuse'])
    return [t.contiguous() for t in out]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--bn', default='eval', choices=['eval', 'train'])
    ap.add_argument('--dropout', action='store_true', help='--bn train: MMRI_I2P attention dropout (module rate, seed = step number)')
    ap.add_argument('--profile', action='store_true', help='after the timed steps: one more step with per-entry-point device times')
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import bench
    from deepinteraction_b200 import backward, ops, synth
    from deepinteraction_b200.shard import GradBuckets
    torch.set_grad_enabled(False)
    neck, _ = bench.build_models(dev)
    fr = synth.to_device(synth.make_frame_batch(bench.SEED + 7 + rank, batch=1, cloud='lidar'), dev)
    img, pts_conv, pts = neck.forward_nhwc(fr['img_feats'], fr['pts_feats'], fr['img_metas'], fr['pts_metas'])
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    targets = [torch.randn(t.shape, device=dev, generator=g) for t in (img, pts_conv, pts)]
    two, mtwo = torch.full((1,), 2.0, device=dev), torch.full((1,), -2.0, device=dev)

    def step():
        o = neck.forward_nhwc(fr['img_feats'], fr['pts_feats'], fr['img_metas'], fr['pts_metas'])
        # d/d out of sum (out - target)^2 = 2 out - 2 target
        zero = [torch.zeros_like(t) for t in o]
        grads_out = [ops.axpy(ops.axpy(z, t.contiguous(), two), tg, mtwo) for z, t, tg in zip(zero, o, targets)]
        r = backward.encoder_backward(neck, fr['img_feats'], fr['pts_feats'], fr['img_metas'], fr['pts_metas'], *grads_out)
        buckets = GradBuckets()
        gl = flat_grads(r)
        for t in reversed(gl):
            buckets.add(t)
        buckets.finish()
        return gl, buckets.launched
    if args.bn == 'train':
        from deepinteraction_b200 import train
        neck.train()
        params = dict(neck.named_parameters())
        opt = torch.optim.AdamW(neck.parameters(), lr=1e-4, weight_decay=0.01)

        def grad_fn(*outs):
            zero = [torch.zeros_like(t) for t in outs]
            return [ops.axpy(ops.axpy(z, t.contiguous(), two), tg, mtwo) for z, t, tg in zip(zero, outs, targets)]

        it = [0]

        def step():                                                      # noqa: F811
            it[0] += 1
            buckets = GradBuckets()                 # buckets fill in backward order and are all-reduced while the backward continues
            r = train.encoder_train_step(neck, fr['img_feats'], fr['pts_feats'], fr['img_metas'], fr['pts_metas'], grad_fn,
                                         dropout_seed=it[0] if args.dropout else None, on_grad=lambda n, t: buckets.add(t))
            buckets.finish()
            gl = list(r['grads'].values())
            for n, t in r['grads'].items():
                params[n].grad = t.view_as(params[n])
            opt.step()
            return gl, buckets.launched
    gl, launched = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        gl, launched = step()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        nbytes = sum(t.numel() * 4 for t in gl)
        print(json.dumps(dict(metric='frames/sec MMRI encoder forward + backward + gradient all-reduce%s (partial config 3: %s-mode '
                                     'BatchNorm, no decoder backward)' % (' + AdamW step' if args.bn == 'train' else '', args.bn), value=world * 1000.0 / ms, unit='frames/s',
                              n_gpus=world, steps=args.steps, ms_per_step=ms, gradient_tensors=len(gl), gradient_bytes=nbytes,
                              allreduce_buckets=launched, finite=bool(all(torch.isfinite(t).all() for t in gl)))), flush=True)
    if args.profile and rank == 0:
        import collections
        import time
        ops.PROFILE[0] = []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        rec, ops.PROFILE[0] = ops.PROFILE[0], None
        agg = collections.defaultdict(lambda: [0, 0.0])
        for name, e0, e1, *_ in rec:
            a_ = agg[name.split(' ')[0]]
            a_[0] += 1
            a_[1] += e0.elapsed_time(e1)
        tot = sum(v[1] for v in agg.values())
        print('profiled step: wall %.1f ms, %d libdi_b200 calls, sum of their device times %.1f ms' % (wall, len(rec), tot))
        for name, (n, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
            print('  %-28s %4d calls %8.2f ms  %5.1f%%' % (name, n, ms_, 100 * ms_ / tot))
        tags = collections.defaultdict(lambda: [0, 0.0])
        for name, e0, e1, *_ in rec:
            if name.startswith('di_linear_f32'):
                tags[name][0] += 1
                tags[name][1] += e0.elapsed_time(e1)
        for name, (n, ms_) in sorted(tags.items(), key=lambda kv: -kv[1][1])[:8]:
            print('    %-44s %4d calls %8.2f ms' % (name, n, ms_))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

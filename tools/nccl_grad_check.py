#!/usr/bin/env python
"""Data-parallel backward of a LocalContextAttentionBlock on N GPUs: every rank runs the forward / backward kernels on its
frames, the parameter gradients go through shard.GradBuckets (bucketed NCCL all-reduce launched while the backward is still
producing gradients), and the averaged result is compared with the single-process gradient of the whole batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_grad_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import oracle.mmri as om
    from deepinteraction_b200 import backward, mmri, synth
    from deepinteraction_b200.shard import GradBuckets, frame_slice
    torch.manual_seed(3)
    blk = om.LocalContextAttentionBlock(128, 128, 9).eval()
    synth.randomize_norm_stats(blk, 3)
    pk = mmri._pack_lcab(blk, dev)
    g = torch.Generator().manual_seed(4)
    N, H, W, C = 2 * world, 112, 200, 128                      # two camera maps per rank
    x = torch.randn(N * H * W, C, generator=g)
    G = torch.randn(N * H * W, C, generator=g)
    sl = frame_slice(N, world, rank)
    rows = slice(sl.start * H * W, sl.stop * H * W)
    xl, Gl = x[rows].to(dev), G[rows].to(dev)
    n_loc = sl.stop - sl.start

    def step():
        r = backward.lcab_backward(pk, xl, xl, n_loc, H, W, Gl)
        buckets = GradBuckets(bucket_bytes=128 << 10)
        grads = []
        for name in ('v', 'k1', 'q1', 'k2', 'q2'):
            for t in r[name]:
                grads.append(t)
                buckets.add(t)
        buckets.finish()
        return grads, buckets.launched
    step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    grads, launched = step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # reference: the whole batch on this rank alone, divided by the world size (what averaging gives)
    full = backward.lcab_backward(pk, x.to(dev), x.to(dev), N, H, W, G.to(dev))
    want = [t / world for name in ('v', 'k1', 'q1', 'k2', 'q2') for t in full[name]]
    err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)) for a, b in zip(grads, want))
    errs = [None] * world
    dist.all_gather_object(errs, err)
    if rank == 0:
        print('world %d: %d buckets all-reduced over NCCL, backward + all-reduce %.2f ms per step, max rel err vs the '
              'whole-batch gradient %.2e' % (world, launched, dt * 1e3, max(errs)), flush=True)
        assert max(errs) < 1e-4
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Multi-GPU partitioning of the MMRI+MMPI forward: frames are independent (eval-mode BatchNorm uses
running statistics and no op mixes samples; reference deepinteraction_encoder.py:79-85,
deepinteraction_decoder.py:201-313), so the batch is sharded into contiguous per-rank slices with NO
data-path collective.  torch.distributed is used only for the timing barrier / max-over-ranks and, when
asked, to gather the per-frame results."""
import torch
import torch.distributed as dist


def frame_slice(total, world, rank):
    """Contiguous slice of `total` frames owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def max_over_ranks(value, device='cpu'):
    """Max of a python float over all ranks (device times are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frames(local, total, dim=0):
    """All-gather per-frame results (tensor whose `dim` indexes this rank's frames) into frame order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [frame_slice(total, world, r) for r in range(world)]
    sizes = [s.stop - s.start for s in sizes]
    mx = max(sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = mx
    buf = local.new_zeros(pad_shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o.narrow(dim, 0, n) for o, n in zip(outs, sizes)], dim)


class GradBuckets:
    """Bucketed gradient all-reduce for data-parallel training ("one bucketed NCCL all-reduce of gradients,
    overlapped with backward"; the reference gets it from MMDistributedDataParallel, tools/train.py:157).

    Gradients are appended in the order the backward produces them; a bucket is launched (async all_reduce on the default
    process group: NCCL on GPUs, gloo in the CPU tests) as soon as it holds `bucket_bytes`, so the transfer of early
    buckets overlaps the rest of the backward.  ``finish()`` launches the last bucket, waits for all of them and writes
    the rank-AVERAGED gradients back into the tensors that were added (DDP semantics)."""

    def __init__(self, bucket_bytes=25 << 20):
        self.bucket_bytes = bucket_bytes
        self._pending, self._pending_bytes, self._inflight = [], 0, []
        self.launched = 0

    def _active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def add(self, grad):
        self._pending.append(grad)
        self._pending_bytes += grad.numel() * grad.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self._launch()

    def _launch(self):
        if not self._pending:
            return
        grads, self._pending, self._pending_bytes = self._pending, [], 0
        flat = torch.cat([g.reshape(-1) for g in grads])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True) if self._active() else None
        self._inflight.append((grads, flat, work))
        self.launched += 1

    def finish(self):
        self._launch()
        world = dist.get_world_size() if self._active() else 1
        for grads, flat, work in self._inflight:
            if work is not None:
                work.wait()
            off = 0
            for g in grads:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g) / world)
                off += n
        self._inflight = []

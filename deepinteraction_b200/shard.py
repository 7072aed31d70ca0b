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

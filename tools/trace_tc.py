"""Pipeline trace of the tcgen05 GEMM (CTA 0): prints per-chunk timings of each role."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepinteraction_b200 import ops, fold, _lib

L = _lib.lib()
dev = torch.device('cuda:0')
for (M, N, K) in [(134400, 128, 128)]:
    A = torch.randn(M, K, device=dev)
    W = fold.Weight(torch.randn(N, K) / 11, dev)
    b = torch.randn(N, device=dev)
    for _ in range(3):
        ops.linear([A], W, b, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear([A], W, b, 1)
    e1.record()
    torch.cuda.synchronize()
    print(f'== M={M} N={N} K={K}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch')
    for mode in (3, 5):
        L.di_tc_set_debug(mode)
        for _ in range(3):
            ops.linear([A], W, b, 1)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.linear([A], W, b, 1)
        e1.record()
        torch.cuda.synchronize()
        print(f'   debug mode {mode} (2=skip LDTM, 4=skip store): {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per launch')
    L.di_tc_set_debug(1)
    ops.linear([A], W, b, 1)
    torch.cuda.synchronize()
    L.di_tc_set_debug(0)
    buf = (ctypes.c_longlong * (8 * 512))()
    _lib.check(L.di_tc_debug_read(buf), 'dbg')
    t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
    t0 = t[0, 0]
    nk = K // 32
    ntile = 8
    names = ['issue', 'landed', 'split', 'mma_rdy', 'mma_iss']
    for it in range(min(nk * ntile, 40)):
        print(it, ' '.join(f'{names[r]}={int(t[r, it] - t0):7d}' for r in range(5)))
    for ch in range(12):
        print('epi chunk', ch, 'pre_ld', int(t[7, ch * 4] - t0), 'post_ld', int(t[7, ch * 4 + 1] - t0), 'post_sts',
              int(t[7, ch * 4 + 2] - t0), 'post_fence', int(t[7, ch * 4 + 3] - t0))
    for tl in range(ntile):
        print('tile', tl, 'acc_ready', int(t[5, tl] - t0), 'stored', int(t[6, tl] - t0))

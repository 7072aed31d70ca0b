"""Timing of the tcgen05 GEMM variants (3 = A via TMEM + resident W, 4 = A via TMEM, 2 = A via smem)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepinteraction_b200 import ops, fold, _lib

L = _lib.lib()
dev = torch.device('cuda:0')
for (M, N, K) in [(134400, 128, 128), (134400, 384, 128), (134400, 128, 384), (32400, 128, 128)]:
    A = torch.randn(M, K, device=dev)
    A2 = torch.randn(M, K, device=dev)
    W = fold.Weight(torch.randn(N, K) / 11, dev)
    b = torch.randn(N, device=dev)
    line = f'M={M:6d} N={N:5d} K={K:3d}:'
    for mode in (3, 19):
        L.di_tc_set_mode(3)
        L.di_tc_set_debug(16 if mode == 19 else 0)
        for _ in range(3):
            ops.linear([A], W, b, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            ops.linear([A if i % 2 else A2], W, b, 1)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        gb = 4 * (M * K + M * N + N * K) / us / 1e3
        line += f'  mode{mode} {us:7.1f} us ({gb:6.0f} GB/s)'
    L.di_tc_set_debug(0)
    print(line, flush=True)

"""clock64 pipeline trace of CTA 0 of the tcgen05 GEMM (v3), normal and skip-store."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepinteraction_b200 import ops, fold, _lib

L = _lib.lib()
dev = torch.device('cuda:0')
M, N, K = 134400, 128, 128
A = torch.randn(M, K, device=dev)
W = fold.Weight(torch.randn(N, K) / 11, dev)
b = torch.randn(N, device=dev)
for dbg in (1,):
    for _ in range(2):
        ops.linear([A], W, b, 1)
    torch.cuda.synchronize()
    L.di_tc_set_debug(dbg)
    ops.linear([A], W, b, 1)
    torch.cuda.synchronize()
    L.di_tc_set_debug(0)
    buf = (ctypes.c_longlong * (8 * 512))()
    _lib.check(L.di_tc_debug_read(buf), 'dbg')
    t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
    t0 = t[0, 0]
    print('==== dbg', dbg)
    names = ['issue', 'landed', 'split', 'mma_rdy', 'mma_iss']
    for it in range(0):
        print(it, ' '.join(f'{names[r]}={int(t[r, it] - t0):7d}' for r in range(5)))
    for ch in range(8):
        print('chunk', ch, ' '.join(f'{nm}={int(t[7, ch * 5 + i] - t0)}' for i, nm in enumerate(['pre_ld', 'ld', 'act', 'sts', 'stg'])))
    for tl in range(8):
        print('tile', tl, 'acc_ready', int(t[5, tl] - t0), 'stored', int(t[6, tl] - t0))

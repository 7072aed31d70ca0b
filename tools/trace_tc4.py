"""MMA pace experiments on the v3 pipeline: dbg bits 32 (MMA ignores the splitter) and 64 (1 of 3 products)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepinteraction_b200 import ops, fold, _lib

L = _lib.lib()
dev = torch.device('cuda:0')


def run(label, fn, nk):
    for dbg in (1, 1 | 64, 1 | 4, 1 | 4 | 64):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        L.di_tc_set_debug(dbg)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        L.di_tc_set_debug(0)
        buf = (ctypes.c_longlong * (8 * 512))()
        _lib.check(L.di_tc_debug_read(buf), 'dbg')
        t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
        n = min(512, nk * 6)
        d = t[:, 8:n]
        names = ['issue', 'landed', 'split', 'mma_rdy', 'mma_iss']
        print('%s dbg=%3d: %.1f us  pace/chunk: ' % (label, dbg, e0.elapsed_time(e1) * 1e3) +
              ' '.join('%s=%.0f' % (names[r], (d[r, -1] - d[r, 0]) / (d.shape[1] - 1)) for r in range(5)) +
              '  rdy->iss=%.0f' % (d[4] - d[3]).mean())


M, N, K = 134400, 128, 128
A = torch.randn(M, K, device=dev)
W = fold.Weight(torch.randn(N, K) / 11, dev)
b = torch.randn(N, device=dev)
run('linear 134400x128x128', lambda: ops.linear([A], W, b, 1), K // 32)
x = torch.randn(6, 112, 200, 256, device=dev)
Wc = fold.Weight(torch.randn(128, 9 * 256) / 30, dev)
run('conv 6x112x200x256', lambda: ops.conv3x3(x, Wc, b, 128, True), 72)

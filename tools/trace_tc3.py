"""clock64 trace of CTA 0 of the tcgen05 GEMM / conv (v3): per-role pace and epilogue phases."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from deepinteraction_b200 import ops, fold, _lib

L = _lib.lib()
dev = torch.device('cuda:0')


def trace(label, fn, nk, flops, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    junk = torch.empty(64 << 20, device=dev)
    ts = []
    for _ in range(5):
        junk.zero_()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = float(np.median(ts))
    print('%s: %.1f us  (%.0f GB/s algorithmic, %.1f TFLOP/s)' % (label, us, nbytes / us / 1e3, flops / us / 1e6))
    L.di_tc_set_debug(1)
    fn()
    torch.cuda.synchronize()
    L.di_tc_set_debug(0)
    buf = (ctypes.c_longlong * (8 * 512))()
    _lib.check(L.di_tc_debug_read(buf), 'dbg')
    t = np.array(buf[:], dtype=np.int64).reshape(8, 512)
    t0 = t[7, 100]
    print('  setup_done=%d end=%d' % (t[7, 101] - t0, t[7, 102] - t0))
    names = ['issue', 'landed', 'split', 'mma_rdy', 'mma_iss']
    n = min(512, nk * 7)
    for it in list(range(min(6, n))) + list(range(max(6, n - 3), n)):
        print('   %3d ' % it + ' '.join(f'{names[r]}={int(t[r, it] - t0):7d}' for r in range(5)))
    if n > 16:
        d = t[:, 8:n]
        print('   mean pace/chunk: ' + ' '.join('%s=%.0f' % (names[r], (d[r, -1] - d[r, 0]) / (d.shape[1] - 1)) for r in range(5)),
              ' load latency=%.0f' % (d[1] - d[0]).mean())
    for tl in range(3):
        print('   tile', tl, 'acc_ready', int(t[5, tl] - t0), 'stored', int(t[6, tl] - t0),
              ' halves:', [int(t[7, (tl * 2 + h) * 3 + i] - t0) for h in range(2) for i in range(3)])


for (M, N, K) in ((200, 128, 128), (200, 32768, 128), (134400, 128, 128)):
    A = torch.randn(M, K, device=dev)
    W = fold.Weight(torch.randn(N, K) / 11, dev)
    b = torch.randn(N, device=dev)
    trace('linear M=%d N=%d K=%d' % (M, N, K), lambda: ops.linear([A], W, b, 1), K // 32, 2.0 * M * N * K,
          4.0 * (M * K + M * N))
for (n, h, w, cin, cout) in ((6, 112, 200, 256, 128), (1, 180, 180, 128, 128)):
    x = torch.randn(n, h, w, cin, device=dev)
    W = fold.Weight(torch.randn(cout, 9 * cin) / 30, dev)
    b = torch.randn(cout, device=dev)
    trace('conv %dx%dx%d Cin=%d Cout=%d' % (n, h, w, cin, cout), lambda: ops.conv3x3(x, W, b, cout, True), 9 * cin // 32,
          2.0 * n * h * w * cout * 9 * cin, 4.0 * n * h * w * (cin + cout))

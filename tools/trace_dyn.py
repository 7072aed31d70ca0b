import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepinteraction_b200 import ops, fold
dev = torch.device('cuda:0')
M, N, K = 200, 32768, 128
A = torch.randn(M, K, device=dev); W = fold.Weight(torch.randn(N, K) / 11, dev); b = torch.randn(N, device=dev)
junk = torch.empty(64 << 20, device=dev)
for bf in (True, False):
    ops.TC_BF16[0] = bf
    for _ in range(3): ops.linear([A], W, b, 0)
    ts = []
    for _ in range(7):
        junk.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear([A], W, b, 0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print('M=200 N=32768 K=128 bf16=%s: %.1f us (min %.1f)' % (bf, float(np.median(ts)), min(ts)))

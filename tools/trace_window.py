#!/usr/bin/env python
"""clock64 pipeline trace of CTA 0 of the tcgen05 window kernel (lcab_tc.cu WSTAMP): prints, per tile, the cycle
offsets of the MMA issuer, two softmax warps, one epilogue warp and the TMA producer.  Needs a B200 (gpurun)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from deepinteraction_b200 import ops, fold, _lib
    dev = torch.device('cuda:0')
    N, H, W, C = 6, 112, 200, 128
    g = torch.Generator().manual_seed(1)
    mk = lambda: fold.split_rows(torch.randn(N * H * W, C, generator=g).to(dev), 3)
    q, k, v = mk(), mk(), mk()
    for _ in range(3):
        ops.lcab_window_tc(q, k, v, N, H, W, C)
    torch.cuda.synchronize()
    L = _lib.lib()
    L.di_lcab_window_tc_set_debug(1)
    ops.lcab_window_tc(q, k, v, N, H, W, C)
    torch.cuda.synchronize()
    L.di_lcab_window_tc_set_debug(0)
    buf = (ctypes.c_longlong * (6 * 256))()
    _lib.check(L.di_lcab_window_tc_debug_read(buf), 'debug_read')
    st = [[buf[s * 256 + i] for i in range(256)] for s in range(6)]
    t0 = st[0][0]
    names = {0: 'mma : start qfull K0 K1 K2 K3 Sdone oempty P0 P1 P2 P3 end Vfull0', 1: 'smx2: start p1done p2start pf0 pf1 pf2 pf3',
             2: 'smx13: start p1done p2start pf0 pf1 pf2 pf3', 3: 'epi : start sumfull ofull oempty', 4: 'prod: K0 K1 K2 K3 V0 V1 V2 V3 Qnext'}
    for s in range(5):
        print(names[s])
        for t in range(8):
            row = st[s][t * 16:(t + 1) * 16]
            print('  tile %d: ' % t + ' '.join('%7d' % (x - t0) if x else '      .' for x in row))


if __name__ == '__main__':
    main()

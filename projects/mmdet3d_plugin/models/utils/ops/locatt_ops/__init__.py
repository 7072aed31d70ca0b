"""Drop-in for the reference's JIT-built ``locatt_ops.localattention`` pybind module
(models/utils/ops/locatt_ops/__init__.py:22-26, localAttention.cpp:61-73): the same five functions with
the same tensor contracts (NCHW fp32 CUDA tensors in, fresh tensors out, current stream, asynchronous),
served by the prebuilt libdi_b200.so instead of a per-import nvcc JIT."""
import ctypes

import torch

from deepinteraction_b200 import _lib
from deepinteraction_b200.ops import _call, _ptr, _stream


def _chk(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError('x must be a CUDA tensor')          # utils.cuh:26-28
        if t.dtype != torch.float32:
            raise RuntimeError('localattention is fp32-only')


def _chk_same(a, b, what):
    if a.dim() != 4 or a.shape != b.shape:
        raise RuntimeError(f'{what}: expected two [N,C,H,W] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}')


def _chk_weight(x, w, kH, kW, what):
    if x.dim() != 4 or w.dim() != 4 or tuple(w.shape) != (x.shape[0], x.shape[2], x.shape[3], kH * kW):
        raise RuntimeError(f'{what}: weight must be [N,H,W,kH*kW] = {(x.shape[0], x.shape[2], x.shape[3], kH * kW)}, '
                           f'got {tuple(w.shape)}')


class localattention:
    @staticmethod
    def similar_forward(x_ori, x_loc, kH, kW):
        _chk(x_ori, x_loc)
        _chk_same(x_ori, x_loc, 'similar_forward')
        x_ori, x_loc = x_ori.contiguous(), x_loc.contiguous()
        N, C, H, W = x_ori.shape
        y = torch.empty(N, H, W, kH * kW, device=x_ori.device, dtype=torch.float32)
        _call('di_locatt_cc2k_f32', _ptr(x_ori), _ptr(x_loc), _ptr(y), N, C, H, W, kH, kW, _stream())
        return y

    @staticmethod
    def similar_backward(x, grad_out, kH, kW, is_ori):
        _chk(x, grad_out)
        _chk_weight(x, grad_out, kH, kW, 'similar_backward')
        x, grad_out = x.contiguous(), grad_out.contiguous()
        N, C, H, W = x.shape
        y = torch.empty_like(x)
        fn = 'di_locatt_ck2c_ori_f32' if is_ori else 'di_locatt_ck2c_loc_f32'
        _call(fn, _ptr(x), _ptr(grad_out), _ptr(y), N, C, H, W, kH, kW, _stream())
        return y

    @staticmethod
    def weighting_forward(x_ori, x_weight, kH, kW):
        _chk(x_ori, x_weight)
        _chk_weight(x_ori, x_weight, kH, kW, 'weighting_forward')
        x_ori, x_weight = x_ori.contiguous(), x_weight.contiguous()
        N, C, H, W = x_ori.shape
        y = torch.empty_like(x_ori)
        _call('di_locatt_ck2c_ori_f32', _ptr(x_ori), _ptr(x_weight), _ptr(y), N, C, H, W, kH, kW, _stream())
        return y

    @staticmethod
    def weighting_backward_ori(x_weight, grad_out, kH, kW):
        _chk(x_weight, grad_out)
        _chk_weight(grad_out, x_weight, kH, kW, 'weighting_backward_ori')
        x_weight, grad_out = x_weight.contiguous(), grad_out.contiguous()
        N, C, H, W = grad_out.shape
        y = torch.empty_like(grad_out)
        _call('di_locatt_ck2c_loc_f32', _ptr(grad_out), _ptr(x_weight), _ptr(y), N, C, H, W, kH, kW, _stream())
        return y

    @staticmethod
    def weighting_backward_weight(x_ori, grad_out, kH, kW):
        _chk(x_ori, grad_out)
        _chk_same(x_ori, grad_out, 'weighting_backward_weight')
        x_ori, grad_out = x_ori.contiguous(), grad_out.contiguous()
        N, C, H, W = x_ori.shape
        y = torch.empty(N, H, W, kH * kW, device=x_ori.device, dtype=torch.float32)
        _call('di_locatt_cc2k_f32', _ptr(grad_out), _ptr(x_ori), _ptr(y), N, C, H, W, kH, kW, _stream())
        return y

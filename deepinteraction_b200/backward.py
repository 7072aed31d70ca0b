"""Backward of LocalContextAttentionBlock on libdi_b200 (SURVEY.md 8(b) `di_lcab_backward`, the training side of the
hot module; forward: mmri.lcab_forward / di_lcab_forward_f32).

Reference: models/utils/encoder_utils.py:36-81 (autograd Functions around the window ops: similar_backward x 2,
weighting_backward_ori, weighting_backward_weight), :119-135 (the block), ops/locatt_ops kernels.cuh:44-119.

Scope: gradients of the block with BatchNorm in EVAL mode (running statistics, folded into the 1x1 convolutions -- the
form the forward kernels use): d/d target map, d/d source map, d/d folded weight and bias of the five Conv+BN layers.
The gradient of a folded weight maps to the convolution weight by the BN scale (dW_conv = dW_folded * gamma / sqrt(var +
eps)), and d beta = d bias_folded.  Train-mode BatchNorm (batch statistics with their own gradient) and dropout are not
covered.  The intermediates are recomputed in fp32 (q1, k1, q, k, v and the [P, 81] attention weights), the window
products run in the unfused kernels of csrc/lcab_bwd.cu, dense products on the tcgen05 / FFMA GEMMs.
"""
import torch

from . import fold, ops


def _bw_pack(pk):
    """Transposed weights for the dX products, built once per pack."""
    if '_bw' in pk:
        return pk['_bw']
    C = pk['C']
    dev = pk['b_self'].device
    w1, w2 = pk['w_self'].w.detach().cpu().double(), pk['w_2'].w.detach().cpu().double()
    wq1, wk1, wv = w1[:C], w1[C:2 * C], w1[2 * C:]
    wq2, wk2 = w2[:C], w2[C:]
    T = lambda m: fold.Weight(m.t().contiguous(), dev)
    pk['_bw'] = dict(q2=T(wq2), k2=T(wk2), q1=T(wq1), kv1=T(torch.cat([wk1, wv], 0)), self1=T(torch.cat([wq1, wk1, wv], 0)))
    return pk['_bw']


def _wgrad(dy, x):
    """dy [M, N], x [M, K] (contiguous rows) -> dy^T x [N, K]: both operands transposed to channel-major, then a split-K
    product over the M pixels with a fixed-order sum of the partials."""
    M = dy.shape[0]
    dyt = ops.nhwc_to_nchw(dy.view(1, M, 1, dy.shape[1])).view(dy.shape[1], M)
    xt = ops.nhwc_to_nchw(x.view(1, M, 1, x.shape[1])).view(x.shape[1], M)
    splits = max(1, min(256, M // 512))
    part = ops.linear([dyt], xt, splits=splits)
    return ops.rows_finish(part) if part.shape[0] > 1 else part[0]


def lcab_backward(pk, target, source, N, H, W, grad_out):
    """target / source / grad_out: [N*H*W, C] fp32 pixel-major rows (same tensor object for target and source = self
    attention).  -> dict(d_target, d_source (None for self attention: summed into d_target), and per layer name in
    (q1, q2, k1, k2, v) a pair (dW_folded [C, C], db_folded [C]))."""
    C, ks = pk['C'], pk['ks']
    M = N * H * W
    bw = _bw_pack(pk)
    self_attn = target is source
    x_t, x_s, dO = target.contiguous(), source.contiguous(), grad_out.contiguous()
    R = ops.ACT_RELU
    # forward intermediates in fp32 rows
    if self_attn:
        t = ops.linear([x_s], pk['w_self'], pk['b_self'], R)
        q1, k1, v = t[:, :C].contiguous(), t[:, C:2 * C].contiguous(), t[:, 2 * C:].contiguous()
    else:
        q1 = ops.linear([x_t], pk['w_q1'], pk['b_q1'], R)
        t = ops.linear([x_s], pk['w_kv1'], pk['b_kv1'], R)
        k1, v = t[:, :C].contiguous(), t[:, C:].contiguous()
    q = ops.linear([q1], pk['w_q2'], pk['b_q2'], R)
    k = ops.linear([k1], pk['w_k2'], pk['b_k2'], R)
    scale = 1.0 / float(C) ** 0.5
    A = ops.win_softmax(ops.win_dot(q, k, N, H, W, ks), scale)
    # window attention
    dA = ops.win_dot(dO, v, N, H, W, ks)
    dv = ops.win_scatter(A, dO, N, H, W, ks)
    dS = ops.win_softmax_bwd(A, dA, scale)
    dq = ops.win_gather(dS, k, N, H, W, ks)
    dk = ops.win_scatter(dS, q, N, H, W, ks)
    # projections: ReLU masks from the saved outputs, dX on the GEMM kernels, dW as split-K products, db as column sums
    g = {}
    dq = ops.relu_bwd(dq, q)
    dk = ops.relu_bwd(dk, k)
    dv = ops.relu_bwd(dv, v)
    g['q2'] = (_wgrad(dq, q1), ops.col_sum(dq))
    g['k2'] = (_wgrad(dk, k1), ops.col_sum(dk))
    dq1 = ops.relu_bwd(ops.linear([dq], bw['q2']), q1)
    dk1 = ops.relu_bwd(ops.linear([dk], bw['k2']), k1)
    g['q1'] = (_wgrad(dq1, x_t), ops.col_sum(dq1))
    g['k1'] = (_wgrad(dk1, x_s), ops.col_sum(dk1))
    g['v'] = (_wgrad(dv, x_s), ops.col_sum(dv))
    if self_attn:
        d_t, d_s = ops.linear([dq1, dk1, dv], bw['self1']), None
    else:
        d_t, d_s = ops.linear([dq1], bw['q1']), ops.linear([dk1, dv], bw['kv1'])
    return dict(d_target=d_t, d_source=d_s, **g)


def i2p_backward(i2p_pack, pts_nhwc, img_nhwc, pts_metas, proj, V, in_hw, grad_out):
    """Backward of the MMRI_I2P block as the product evaluates it (mmri.DeepInteractionEncoder.i2p; reference
    encoder_utils.py:216-320): rows = pts[coors]; qk = M1 rows + c1; s = attend(qk, image samples); o = M2 s + c2;
    out[coors] = o where a pillar saw >= 1 sample.  i2p_pack = (M1, c1, M2, c2) folded from nn.MultiheadAttention
    (fold.i2p_fold); proj: [B, V, 12] device camera rows (mmri.Geometry.proj).  grad_out [B, Y, X, C].
    -> dict(d_pts [B,Y,X,C], d_img [B*V,h,w,C], dM1, dc1, dM2, dc2); fold.i2p_unfold_grads maps the last four to the
    attention module's own parameters."""
    M1, c1, M2, c2 = i2p_pack
    coors = pts_metas['pillar_coors']
    pillars, npts = pts_metas['pillars'], pts_metas['pillars_num_points']
    d_pts = torch.zeros_like(pts_nhwc)
    d_img = torch.zeros_like(img_nhwc)
    if coors.shape[0] == 0:
        z = lambda t: torch.zeros_like(t.w if hasattr(t, 'w') else t)
        return dict(d_pts=d_pts, d_img=d_img, dM1=z(M1), dc1=z(c1), dM2=z(M2), dc2=z(c2))
    dev = pts_nhwc.device
    T_ = lambda Wt: fold.Weight(Wt.w.detach().cpu().double().t().contiguous(), dev)
    # forward intermediates
    rows = ops.gather_rows(pts_nhwc, coors)
    qk = ops.linear([rows], M1, c1)
    s, cnt = ops.i2p_attend(qk, pillars, npts, coors, proj, img_nhwc, V, in_hw)
    # out = scatter(M2 s + c2) at pillars with cnt > 0
    do = ops.gather_rows_masked(grad_out.contiguous(), cnt, coors)
    dM2, dc2 = _wgrad(do, s), ops.col_sum(do)
    ds = ops.linear([do], T_(M2))
    dqk = ops.i2p_attend_bwd(qk, ds, pillars, npts, coors, proj, img_nhwc, d_img, V, in_hw)
    dM1, dc1 = _wgrad(dqk, rows), ops.col_sum(dqk)
    drows = ops.linear([dqk], T_(M1))
    ones = torch.ones(coors.shape[0], device=dev, dtype=torch.int32)
    ops.scatter_rows(drows, ones, coors, d_pts)             # one pillar per BEV cell: a plain store
    return dict(d_pts=d_pts, d_img=d_img, dM1=dM1, dc1=dc1, dM2=dM2, dc2=dc2)


class LCABFunction(torch.autograd.Function):
    """autograd wrapper: out = LCAB(target, source) with the inference kernels, input gradients through lcab_backward.
    (Parameter gradients are returned by lcab_backward for the caller's optimiser on the folded weights.)"""

    @staticmethod
    def forward(ctx, pk, target, source, N, H, W):
        from .mmri import lcab_forward
        ctx.pk, ctx.dims, ctx.same = pk, (N, H, W), target is source
        ctx.save_for_backward(target, source)
        return lcab_forward(pk, target, source, N, H, W)

    @staticmethod
    def backward(ctx, grad_out):
        target, source = ctx.saved_tensors
        r = lcab_backward(ctx.pk, target, target if ctx.same else source, *ctx.dims, grad_out)
        return None, r['d_target'], r['d_source'], None, None, None

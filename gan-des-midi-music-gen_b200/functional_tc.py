"""torch.autograd glue for the GAN-DES layers on the tensor cores (GAN_DES/SIMNN.py:70-84, :123-142): every contraction -- forward, data
gradient and weight gradient of ``nn.Linear``, ``nn.Conv2d`` and ``nn.ConvTranspose2d`` -- is one ``mmg_gemm_tc`` launch (tcgen05.mma with
TMEM accumulators, TMA-fed, csrc/gemm_tc.cu) on bf16 operands with fp32 accumulation; im2col / col2im / packing are small kernels of the
same library.  Same call signatures as ``functional.py`` (the fp32 SIMT path); the modules pick one of the two.

Bias handling of the convolutions: the im2col rows carry a column of ones, the packed weight matrix a column with the bias, so the bias is
added inside the forward GEMM and its gradient is the extra row of the weight-gradient GEMM.
"""
import torch

from . import _native as N
from .functional import ACT_NONE, ACT_RELU, _f32c

_BF = torch.bfloat16
_epoch = 0
_wcache = {}


def invalidate_weight_cache():
    """The packed bf16 weights are cached per parameter version; optimisers that update parameters through raw pointers (optim.FusedAdam) call this."""
    global _epoch
    _epoch += 1


def _cached(t, tag, make):
    """bf16 copy of parameter ``t``, rebuilt when the parameter changes.  The entry keeps ``t`` alive, so its address cannot be handed to
    another tensor while the entry exists (a bare data_ptr key would alias a freed weight with a new one)."""
    key = (t.data_ptr(), tuple(t.shape), tag)
    ver = (t._version, _epoch)
    hit = _wcache.get(key)
    if hit is None or hit[0] != ver:
        if len(_wcache) >= 256:
            _wcache.clear()
        hit = (ver, make(), t)
        _wcache[key] = hit
    return hit[1]


def _r8(n):
    return (n + 7) // 8 * 8


def _r16(n):
    """im2col row pitch: a multiple of 16 elements (the im2col kernel then stores whole 32-byte sectors), except for tiny K"""
    return (n + 15) // 16 * 16 if n > 8 else 8


def _pack(src, dims, perm, pitch, out=None, a_stride=0):
    """fp32 tensor viewed as (d0,d1,d2) -> bf16 [n_pa * n_pb][pitch] (zero padded), or into ``out`` with ``a_stride`` elements between the pa blocks."""
    if out is None:
        out = torch.empty(dims[perm[0]] * dims[perm[1]], pitch, device=src.device, dtype=_BF)
    N.call("mmg_pack_bf16", N.ptr(src), N.ptr(out), dims[0], dims[1], dims[2], perm[0], perm[1], perm[2], pitch, a_stride, N.stream())
    return out


def _gemm(A, a_mn, lda, B, b_mn, ldb, C, ldc, M, Nn, K, split_k=1, trans_out=False, inner=0, atomic=False, bias=None, bias_on_m=False, act=ACT_NONE):
    N.call("mmg_gemm_tc", N.ptr(A), int(a_mn), lda, N.ptr(B), int(b_mn), ldb, N.ptr(C), ldc, M, Nn, K, 0, split_k, int(trans_out), inner, int(atomic),
           N.ptr(bias), int(bias_on_m), act, N.stream())


def _split(K, tiles):
    """split-K factor: enough CTAs for the 148 SMs, at least two 64-element chunks each"""
    chunks = (K + 63) // 64
    return max(1, min(chunks // 2, (2 * 148 + tiles - 1) // tiles))


def _act_bwd(y, dy, act):
    if act == ACT_NONE:
        return dy
    dz = torch.empty_like(dy)
    N.call("mmg_act_bwd_f32", N.ptr(y), N.ptr(dy), N.ptr(dz), dy.numel(), act, N.stream())
    return dz


class LinearActTC(torch.autograd.Function):
    """y = act(x @ w.T + b).  The weight is the M operand of every GEMM (the batch is small): forward y^T = W x^T (split-K over long rows),
    dx^T = W^T dz^T with W read MN-major, dW = dz^T x with both operands MN-major (K = batch)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        x, w = _f32c(x), _f32c(w)
        b = _f32c(b) if b is not None else None
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        M, K, Nf = x2.shape[0], x2.shape[1], w.shape[0]
        Kp = _r8(K)
        wp = _cached(w, "lin", lambda: _pack(w, (1, Nf, K), (0, 1, 2), Kp))
        xp = _pack(x2, (1, M, K), (0, 1, 2), Kp)
        split = _split(K, (Nf + 127) // 128) if K >= 4096 else 1
        if split > 1:
            y = torch.zeros(M, Nf, device=x.device)
            _gemm(wp, 0, Kp, xp, 0, Kp, y, Nf, Nf, M, K, split_k=split, trans_out=True, inner=Nf, atomic=True)
            N.call("mmg_bias_act_inplace_f32", N.ptr(y), N.ptr(b), M, Nf, act, N.stream())
        else:
            y = torch.empty(M, Nf, device=x.device)
            _gemm(wp, 0, Kp, xp, 0, Kp, y, Nf, Nf, M, K, trans_out=True, inner=Nf, bias=b, bias_on_m=True, act=act)
        ctx.save_for_backward(xp, w, y if act != ACT_NONE else None)
        ctx.act, ctx.lead, ctx.has_b, ctx.dims = act, lead, b is not None, (M, K, Nf, Kp)
        return y.reshape(*lead, Nf)

    @staticmethod
    def backward(ctx, dy):
        xp, w, y = ctx.saved_tensors
        M, K, Nf, Kp = ctx.dims
        dz = _act_bwd(y, _f32c(dy).reshape(M, Nf), ctx.act)
        Np = _r8(Nf)
        dzp = _pack(dz, (1, M, Nf), (0, 1, 2), Np)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wp = _cached(w, "lin", lambda: _pack(w, (1, Nf, K), (0, 1, 2), Kp))
            dx = torch.empty(M, K, device=dz.device)
            _gemm(wp, 1, Kp, dzp, 0, Np, dx, K, K, M, Nf, trans_out=True, inner=K)          # C'[f][b] stored as dx[b][f]
        if ctx.needs_input_grad[1]:
            dw = torch.empty(Nf, K, device=dz.device)
            _gemm(dzp, 1, Np, xp, 1, Kp, dw, K, Nf, K, M)
        if ctx.has_b and ctx.needs_input_grad[2]:
            db = torch.empty(Nf, device=dz.device)
            N.call("mmg_colsum_f32", N.ptr(dz), N.ptr(db), M, Nf, N.stream())
        return (dx.reshape(*ctx.lead, K) if dx is not None else None), dw, db, None


class Conv2dActTC(torch.autograd.Function):
    """y = act(conv2d(x, w, b, stride, padding)) as im2col + GEMM; the output leaves the GEMM in NCHW.  The data gradient (stride 1) is the
    convolution of dz with the flipped, transposed weights; the weight gradient reads the forward's im2col rows MN-major."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act):
        x, w = _f32c(x), _f32c(w)
        b = _f32c(b) if b is not None else None
        Nn, Ci, H, W = x.shape
        Co, _, kh, kw = w.shape
        OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        K = Ci * kh * kw
        Kp = _r16(K + 1)

        def make_w():
            ext = torch.zeros(Co, Kp, device=w.device)
            ext[:, :K] = w.reshape(Co, K)
            if b is not None:
                ext[:, K] = b
            return ext.to(_BF)
        wp = _cached(w, ("conv", None if b is None else (b.data_ptr(), b._version)), make_w)
        P = Nn * OH * OW
        col = torch.empty(P, Kp, device=x.device, dtype=_BF)
        N.call("mmg_im2col_bf16", N.ptr(x), N.ptr(col), Nn, Ci, H, W, kh, kw, stride, pad, Kp, 1, N.stream())
        y = torch.empty(Nn, Co, OH, OW, device=x.device)
        _gemm(col, 0, Kp, wp, 0, Kp, y, Co, P, Co, K + 1, trans_out=True, inner=OH * OW, act=act)
        ctx.save_for_backward(col, w, y if act != ACT_NONE else None)
        ctx.cfg = (Nn, Ci, H, W, Co, kh, kw, stride, pad, OH, OW, K, Kp)
        ctx.act, ctx.has_b = act, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        col, w, y = ctx.saved_tensors
        Nn, Ci, H, W, Co, kh, kw, stride, pad, OH, OW, K, Kp = ctx.cfg
        dz = _act_bwd(y, _f32c(dy), ctx.act)
        P = Nn * OH * OW
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if stride != 1:
                dx = torch.empty(Nn, Ci, H, W, device=dz.device)           # (not on the GAN-DES path: every Conv2d there has stride 1)
                N.call("mmg_conv2d_bwd_data_f32", N.ptr(dz), N.ptr(w), None, N.ptr(dx), Nn, Ci, H, W, Co, kh, kw, stride, pad, ACT_NONE, N.stream())
            else:
                K2 = Co * kh * kw
                K2p = _r16(K2)
                wf = _cached(w, "convT", lambda: w.flip(2, 3).permute(1, 0, 2, 3).reshape(Ci, K2).contiguous())
                wfp = _cached(w, "convTp", lambda: _pack(wf, (1, Ci, K2), (0, 1, 2), K2p))
                col2 = torch.empty(Nn * H * W, K2p, device=dz.device, dtype=_BF)
                N.call("mmg_im2col_bf16", N.ptr(dz), N.ptr(col2), Nn, Co, OH, OW, kh, kw, 1, kh - 1 - pad, K2p, 0, N.stream())
                dx = torch.empty(Nn, Ci, H, W, device=dz.device)
                _gemm(col2, 0, K2p, wfp, 0, K2p, dx, Ci, Nn * H * W, Ci, K2, trans_out=True, inner=H * W)
        if ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2]):
            Pp = _r8(P)                                                      # dz as [Co][b * p] rows (K-major B), row pitch padded to 16 bytes
            dzt = _pack(dz, (Nn, Co, OH * OW), (1, 0, 2), OH * OW, out=torch.empty(Co, Pp, device=dz.device, dtype=_BF), a_stride=Pp)
            buf = torch.zeros(Co, K + 1, device=dz.device)
            _gemm(col, 1, Kp, dzt, 0, Pp, buf, K + 1, K + 1, Co, P, split_k=_split(P, (K + 128) // 128), trans_out=True, inner=K + 1, atomic=True)
            dw = buf[:, :K].reshape(w.shape).contiguous()
            db = buf[:, K].contiguous() if ctx.has_b else None
        return dx, dw, db, None, None, None


class Conv2dReluPoolTC(torch.autograd.Function):
    """maxpool2(relu(conv2d(x, w, b, 1, padding))) -- one block of the GAN-DES discriminator (SIMNN.py:138-139).  Forward: im2col + GEMM (bias,
    ReLU, NCHW store in the epilogue) + the pooling kernel; only the pooled output and the argmax codes are kept.  Backward: ONE kernel
    undoes pooling and ReLU and writes the gradient both as fp32 NCHW (for the data-gradient GEMM) and as bf16 [oc][b*p] rows (the K-major B
    operand of the weight-gradient GEMM) -- the pre-pool activation is never read again."""

    @staticmethod
    def forward(ctx, x, w, b, pad):
        x, w = _f32c(x), _f32c(w)
        b = _f32c(b) if b is not None else None
        Nn, Ci, H, W = x.shape
        Co, _, kh, kw = w.shape
        OH, OW = H + 2 * pad - kh + 1, W + 2 * pad - kw + 1
        K = Ci * kh * kw
        Kp = _r16(K + 1)

        def make_w():
            ext = torch.zeros(Co, Kp, device=w.device)
            ext[:, :K] = w.reshape(Co, K)
            if b is not None:
                ext[:, K] = b
            return ext.to(_BF)
        P = Nn * OH * OW
        yp = torch.empty(Nn, Co, OH // 2, OW // 2, device=x.device)
        idx = torch.empty(Nn, Co, OH // 2, OW // 2, device=x.device, dtype=torch.uint8)
        stencil = (Ci, kh, kw) == (1, 2, 2) and Co <= 32
        if stencil:
            # K = 4 is a stencil, not a GEMM: conv + bias + ReLU + pool in one fp32 kernel, the (B,Co,OH,OW) activation never exists; the
            # backward (a reduction over all pixels: a GEMM again) rebuilds the im2col rows of x
            N.call("mmg_conv_small_relu_pool_f32", N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(yp), N.ptr(idx), Nn, Ci, H, W, Co, kh, kw, pad, N.stream())
            ctx.save_for_backward(x, w, yp, idx)
        else:
            wp = _cached(w, ("conv", None if b is None else (b.data_ptr(), b._version)), make_w)
            col = torch.empty(P, Kp, device=x.device, dtype=_BF)
            N.call("mmg_im2col_bf16", N.ptr(x), N.ptr(col), Nn, Ci, H, W, kh, kw, 1, pad, Kp, 1, N.stream())
            y = torch.empty(Nn, Co, OH, OW, device=x.device)
            _gemm(col, 0, Kp, wp, 0, Kp, y, Co, P, Co, K + 1, trans_out=True, inner=OH * OW, act=ACT_RELU)
            N.call("mmg_maxpool2_fwd_f32", N.ptr(y), N.ptr(yp), N.ptr(idx), Nn * Co, OH, OW, N.stream())
            ctx.save_for_backward(col, w, yp, idx)
        ctx.cfg = (Nn, Ci, H, W, Co, kh, kw, pad, OH, OW, K, Kp)
        ctx.has_b, ctx.stencil = b is not None, stencil
        return yp

    @staticmethod
    def backward(ctx, dyp):
        col, w, yp, idx = ctx.saved_tensors
        Nn, Ci, H, W, Co, kh, kw, pad, OH, OW, K, Kp = ctx.cfg
        dyp = _f32c(dyp)
        P = Nn * OH * OW
        Pp = _r8(P)
        if ctx.stencil and (ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2])):
            x, col = col, torch.empty(P, Kp, device=dyp.device, dtype=_BF)
            N.call("mmg_im2col_bf16", N.ptr(x), N.ptr(col), Nn, Ci, H, W, kh, kw, 1, pad, Kp, 1, N.stream())
        need_dx = ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2])
        gather = need_dx and Ci == 16 and Co % 8 == 0          # data gradient as GEMM + gather (below) instead of im2col(dz) + GEMM
        dz = torch.empty(Nn, Co, OH, OW, device=dyp.device) if (need_dx and not gather) else None
        dzt = torch.empty(Co, Pp, device=dyp.device, dtype=_BF) if need_dw else None
        dzn = torch.empty(P, Co, device=dyp.device, dtype=_BF) if gather else None
        N.call("mmg_pool_relu_bwd", N.ptr(dyp), N.ptr(idx), N.ptr(yp), N.ptr(dz), N.ptr(dzt), N.ptr(dzn), Nn, Co, OH, OW, Pp, N.stream())
        dx = dw = db = None
        if gather:
            # dx = sum over taps of (dz x W_tap) shifted back: ONE GEMM dz[p][oc] x W[(ky,kx,ci)][oc] writes the tap columns in bf16 (60 MB for
            # the GAN-DES conv2 instead of the 119 MB im2col of dz written and read again), a gather kernel adds the nine shifted runs
            T = kh * kw * Ci
            wt = _cached(w, "dgradT", lambda: w.permute(2, 3, 1, 0).reshape(T, Co).contiguous())      # [(ky,kx,ci)][oc]
            wtp = _cached(w, "dgradTp", lambda: _pack(wt, (1, T, Co), (0, 1, 2), Co))
            dcol = torch.empty(P, T, device=dyp.device, dtype=_BF)
            N.call("mmg_gemm_tc", N.ptr(dzn), 0, Co, N.ptr(wtp), 0, Co, N.ptr(dcol), T, P, T, Co, 0, 1, 2, 0, 0, None, 0, 0, N.stream())
            dx = torch.empty(Nn, Ci, H, W, device=dyp.device)
            N.call("mmg_conv_dgrad_gather", N.ptr(dcol), N.ptr(dx), Nn, Ci, H, W, kh, kw, pad, T, N.stream())
        elif need_dx:
            K2 = Co * kh * kw
            K2p = _r16(K2)
            wf = _cached(w, "convT", lambda: w.flip(2, 3).permute(1, 0, 2, 3).reshape(Ci, K2).contiguous())
            wfp = _cached(w, "convTp", lambda: _pack(wf, (1, Ci, K2), (0, 1, 2), K2p))
            col2 = torch.empty(Nn * H * W, K2p, device=dz.device, dtype=_BF)
            N.call("mmg_im2col_bf16", N.ptr(dz), N.ptr(col2), Nn, Co, OH, OW, kh, kw, 1, kh - 1 - pad, K2p, 0, N.stream())
            dx = torch.empty(Nn, Ci, H, W, device=dz.device)
            _gemm(col2, 0, K2p, wfp, 0, K2p, dx, Ci, Nn * H * W, Ci, K2, trans_out=True, inner=H * W)
        if need_dw:
            buf = torch.zeros(Co, K + 1, device=dyp.device)
            _gemm(col, 1, Kp, dzt, 0, Pp, buf, K + 1, K + 1, Co, P, split_k=_split(P, (K + 128) // 128), trans_out=True, inner=K + 1, atomic=True)
            dw = buf[:, :K].reshape(w.shape).contiguous()
            db = buf[:, K].contiguous() if ctx.has_b else None
        return dx, dw, db, None


class ConvTranspose2dActTC(torch.autograd.Function):
    """y = act(conv_transpose2d(x, w, None, stride, padding)), w (Cin, Cout, kh, kw), bias-free (SIMNN.py:70-84): GEMM over the input
    channels into per-input-pixel tap columns, then col2im.  Backward: dx = conv2d(dz, w) and dW = x^T col(dz), both GEMMs."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, act):
        x, w = _f32c(x), _f32c(w)
        Nn, Cin, Hin, Win = x.shape
        _, Cout, kh, kw = w.shape
        Hout, Wout = (Hin - 1) * stride - 2 * pad + kh, (Win - 1) * stride - 2 * pad + kw
        P, T = Nn * Hin * Win, Cout * kh * kw
        Cp, Tp = _r8(Cin), _r8(T)
        xp = _pack(x, (Nn, Cin, Hin * Win), (0, 2, 1), Cp)                   # NHWC rows [b*p][Cin]
        wp = _cached(w, "ct", lambda: _pack(w, (1, Cin, T), (0, 1, 2), Tp))  # [Cin][Cout*taps]: MN-major B
        col = torch.empty(P, Tp, device=x.device)
        _gemm(xp, 0, Cp, wp, 1, Tp, col, Tp, P, T, Cin)
        y = torch.empty(Nn, Cout, Hout, Wout, device=x.device)
        N.call("mmg_col2im_f32", N.ptr(col), N.ptr(y), Nn, Cout, Hin, Win, kh, kw, stride, pad, Tp, act, N.stream())
        ctx.save_for_backward(xp, w, y if act != ACT_NONE else None)
        ctx.cfg, ctx.act = (Nn, Cin, Hin, Win, Cout, kh, kw, stride, pad, Hout, Wout, P, T, Cp, Tp), act
        return y

    @staticmethod
    def backward(ctx, dy):
        xp, w, y = ctx.saved_tensors
        Nn, Cin, Hin, Win, Cout, kh, kw, stride, pad, Hout, Wout, P, T, Cp, Tp = ctx.cfg
        dz = _act_bwd(y, _f32c(dy), ctx.act)
        col = torch.empty(P, Tp, device=dz.device, dtype=_BF)                # im2col of dz: rows = input pixels of the transposed convolution
        N.call("mmg_im2col_bf16", N.ptr(dz), N.ptr(col), Nn, Cout, Hout, Wout, kh, kw, stride, pad, Tp, 0, N.stream())
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wk = _cached(w, "ct", lambda: _pack(w, (1, Cin, T), (0, 1, 2), Tp))      # the same matrix, now the K-major B of dx[p][ci] = col[p][t] w[ci][t]
            dx = torch.empty(Nn, Cin, Hin, Win, device=dz.device)
            _gemm(col, 0, Tp, wk, 0, Tp, dx, Cin, P, Cin, T, trans_out=True, inner=Hin * Win)
        if ctx.needs_input_grad[1]:
            dwb = torch.empty(Cin, Tp, device=dz.device)
            _gemm(xp, 1, Cp, col, 1, Tp, dwb, Tp, Cin, T, P)
            dw = dwb[:, :T].reshape(w.shape).contiguous() if Tp != T else dwb.reshape(w.shape)
        return dx, dw, None, None, None


def linear(x, w, b=None, act=ACT_NONE):
    return LinearActTC.apply(x, w, b, act)


def conv2d(x, w, b=None, stride=1, padding=0, act=ACT_NONE):
    return Conv2dActTC.apply(x, w, b, stride, padding, act)


def conv_transpose2d(x, w, stride=1, padding=0, act=ACT_NONE):
    return ConvTranspose2dActTC.apply(x, w, stride, padding, act)


def conv2d_relu_pool(x, w, b=None, padding=0):
    return Conv2dReluPoolTC.apply(x, w, b, padding)

"""torch.autograd glue over the C ABI (fp32 path).  Every Function launches only kernels from
libmmgan_b200.so; PyTorch supplies device memory, streams and the autograd tape.

Activation codes: 0 none, 1 LeakyReLU(0.2), 2 ReLU, 3 sigmoid.
"""
import ctypes

import torch

from . import _native as N

ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3


def _f32c(t):
    N.require_cuda(t)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _Workspace:
    """Per-device scratch for the BatchNorm statistics (2*C doubles)."""
    bufs = {}

    @classmethod
    def get(cls, device, nbytes):
        b = cls.bufs.get(device)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
            cls.bufs[device] = b
        return b


class LinearAct(torch.autograd.Function):
    """y = act(x @ w.T + b)  -- nn.Linear (+ LeakyReLU/ReLU/sigmoid) as one kernel."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        x, w = _f32c(x), _f32c(w)
        b = _f32c(b) if b is not None else None
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        M, K, Nf = x2.shape[0], x2.shape[1], w.shape[0]
        y = torch.empty(M, Nf, device=x.device, dtype=torch.float32)
        if M:
            N.call("mmg_linear_fwd_f32", N.ptr(x2), N.ptr(w), N.ptr(b), N.ptr(y), M, Nf, K, act, N.stream())
        ctx.save_for_backward(x2, w, y if act != ACT_NONE else None)
        ctx.act, ctx.lead, ctx.has_b = act, lead, b is not None
        return y.reshape(*lead, Nf)

    @staticmethod
    def backward(ctx, dy):
        x2, w, y = ctx.saved_tensors
        M, K, Nf = x2.shape[0], x2.shape[1], w.shape[0]
        dy = _f32c(dy).reshape(M, Nf)
        if ctx.act != ACT_NONE and M:
            dz = torch.empty_like(dy)
            N.call("mmg_act_bwd_f32", N.ptr(y), N.ptr(dy), N.ptr(dz), dy.numel(), ctx.act, N.stream())
            dy = dz
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        db = torch.zeros(Nf, device=w.device) if (ctx.has_b and ctx.needs_input_grad[2]) else None
        if M:
            N.call("mmg_linear_bwd_f32", N.ptr(x2), N.ptr(w), N.ptr(dy), N.ptr(dx), N.ptr(dw), N.ptr(db), M, Nf, K, 0, N.stream())
        elif dx is not None:
            dx.zero_()
        return (dx.reshape(*ctx.lead, K) if dx is not None else None), dw, db, None


class BatchNormAct(torch.autograd.Function):
    """y = act(BatchNorm(z)) over (N, C, *spatial); training mode updates the running stats in place."""

    @staticmethod
    def forward(ctx, z, gamma, beta, run_mean, run_var, training, momentum, eps, act):
        z, gamma, beta = _f32c(z), _f32c(gamma), _f32c(beta)
        Nn, C = z.shape[0], z.shape[1]
        HW = z.numel() // max(Nn * C, 1)
        y = torch.empty_like(z)
        if training:
            if Nn * HW <= 1:
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z.shape)}")
            mean = torch.empty(C, device=z.device)
            invstd = torch.empty(C, device=z.device)
            ws = _Workspace.get(z.device, 16 * C)
            N.call("mmg_bn_fwd_train_f32", N.ptr(z), N.ptr(gamma), N.ptr(beta), N.ptr(run_mean), N.ptr(run_var), N.ptr(y), N.ptr(mean),
                   N.ptr(invstd), Nn, C, HW, momentum, eps, act, N.ptr(ws), ws.numel(), N.stream())
            ctx.save_for_backward(z, gamma, beta, mean, invstd)
        else:
            N.call("mmg_bn_fwd_eval_f32", N.ptr(z), N.ptr(gamma), N.ptr(beta), N.ptr(run_mean), N.ptr(run_var), N.ptr(y), Nn, C, HW, eps, act, N.stream())
        ctx.training, ctx.act, ctx.dims = training, act, (Nn, C, HW)
        return y

    @staticmethod
    def backward(ctx, dy):
        if not ctx.training:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the reference path")
        z, gamma, beta, mean, invstd = ctx.saved_tensors
        Nn, C, HW = ctx.dims
        dy = _f32c(dy)
        dz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
        dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
        ws = _Workspace.get(z.device, 16 * C)
        N.call("mmg_bn_bwd_f32", N.ptr(z), N.ptr(dy), N.ptr(gamma), N.ptr(beta), N.ptr(mean), N.ptr(invstd), N.ptr(dz), N.ptr(dgamma),
               N.ptr(dbeta), Nn, C, HW, ctx.act, 0, N.ptr(ws), ws.numel(), N.stream())
        return dz, dgamma, dbeta, None, None, None, None, None, None


class Conv2dAct(torch.autograd.Function):
    """y = act(conv2d(x, w, b, stride, padding))."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, act):
        x, w = _f32c(x), _f32c(w)
        b = _f32c(b) if b is not None else None
        Nn, Ci, H, W = x.shape
        Co, _, kh, kw = w.shape
        OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        y = torch.empty(Nn, Co, OH, OW, device=x.device)
        N.call("mmg_conv2d_fwd_f32", N.ptr(x), N.ptr(w), N.ptr(b), N.ptr(y), Nn, Ci, H, W, Co, kh, kw, stride, pad, act, N.stream())
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        ctx.cfg = (Nn, Ci, H, W, Co, kh, kw, stride, pad)
        ctx.act, ctx.has_b = act, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _f32c(dy)
        if ctx.act != ACT_NONE:
            dz = torch.empty_like(dy)
            N.call("mmg_act_bwd_f32", N.ptr(y), N.ptr(dy), N.ptr(dz), dy.numel(), ctx.act, N.stream())
            dy = dz
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            N.call("mmg_conv2d_bwd_data_f32", N.ptr(dy), N.ptr(w), None, N.ptr(dx), *ctx.cfg, ACT_NONE, N.stream())
        if ctx.needs_input_grad[1] or (ctx.has_b and ctx.needs_input_grad[2]):
            dw = torch.empty_like(w)
            db = torch.empty(w.shape[0], device=w.device) if ctx.has_b else None
            N.call("mmg_conv2d_bwd_weight_f32", N.ptr(x), N.ptr(dy), N.ptr(dw), N.ptr(db), *ctx.cfg, 0, N.stream())
        return dx, dw, db, None, None, None


class ConvTranspose2dAct(torch.autograd.Function):
    """y = act(conv_transpose2d(x, w, None, stride, padding)); w is (Cin, Cout, kh, kw), bias-free
    (SIMNN.py:70-84 use bias=False).  Forward = data gradient of the mirrored Conv2d."""

    @staticmethod
    def forward(ctx, x, w, stride, pad, act):
        x, w = _f32c(x), _f32c(w)
        Nn, Cin, Hin, Win = x.shape
        _, Cout, kh, kw = w.shape
        Hout, Wout = (Hin - 1) * stride - 2 * pad + kh, (Win - 1) * stride - 2 * pad + kw
        y = torch.empty(Nn, Cout, Hout, Wout, device=x.device)
        cfg = (Nn, Cout, Hout, Wout, Cin, kh, kw, stride, pad)       # the mirrored conv: (Cout,Hout,Wout) -> (Cin,Hin,Win)
        N.call("mmg_conv2d_bwd_data_f32", N.ptr(x), N.ptr(w), None, N.ptr(y), *cfg, act, N.stream())
        ctx.save_for_backward(x, w, y if act != ACT_NONE else None)
        ctx.cfg, ctx.act = cfg, act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _f32c(dy)
        if ctx.act != ACT_NONE:
            dz = torch.empty_like(dy)
            N.call("mmg_act_bwd_f32", N.ptr(y), N.ptr(dy), N.ptr(dz), dy.numel(), ctx.act, N.stream())
            dy = dz
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            N.call("mmg_conv2d_fwd_f32", N.ptr(dy), N.ptr(w), None, N.ptr(dx), *ctx.cfg, ACT_NONE, N.stream())
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            N.call("mmg_conv2d_bwd_weight_f32", N.ptr(dy), N.ptr(x), N.ptr(dw), None, *ctx.cfg, 0, N.stream())
        return dx, dw, None, None, None


class MaxPool2(torch.autograd.Function):
    """nn.MaxPool2d(kernel_size=2, stride=2)."""

    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        Nn, C, H, W = x.shape
        y = torch.empty(Nn, C, H // 2, W // 2, device=x.device)
        idx = torch.empty(Nn, C, H // 2, W // 2, device=x.device, dtype=torch.uint8)
        N.call("mmg_maxpool2_fwd_f32", N.ptr(x), N.ptr(y), N.ptr(idx), Nn * C, H, W, N.stream())
        ctx.save_for_backward(idx)
        ctx.shape = (Nn, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        Nn, C, H, W = ctx.shape
        dx = torch.empty(Nn, C, H, W, device=dy.device)
        N.call("mmg_maxpool2_bwd_f32", N.ptr(_f32c(dy)), N.ptr(idx), N.ptr(dx), Nn * C, H, W, N.stream())
        return dx


class BCEWithLogits(torch.autograd.Function):
    """Mean-reduced binary cross entropy on logits, forward and backward in one kernel."""

    @staticmethod
    def forward(ctx, logits, target):
        x = _f32c(logits).reshape(-1)
        if torch.is_tensor(target):
            t = _f32c(target).reshape(-1).expand_as(x).contiguous()
            tc = 0.0
        else:
            t, tc = None, float(target)
        loss = torch.empty(1, device=x.device)
        dx = torch.empty_like(x)
        n = x.numel()
        N.call("mmg_bce_logits_f32", N.ptr(x), N.ptr(t), tc, n, N.ptr(loss), 0, N.ptr(dx), 1.0 / n, None, N.stream())
        ctx.save_for_backward(dx)
        ctx.shape = logits.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return (dx * g).reshape(ctx.shape), None


def linear(x, w, b=None, act=ACT_NONE):
    return LinearAct.apply(x, w, b, act)


def batch_norm(z, gamma, beta, run_mean, run_var, training, momentum=0.1, eps=1e-5, act=ACT_NONE):
    return BatchNormAct.apply(z, gamma, beta, run_mean, run_var, training, momentum, eps, act)


def conv2d(x, w, b=None, stride=1, padding=0, act=ACT_NONE):
    return Conv2dAct.apply(x, w, b, stride, padding, act)


def conv_transpose2d(x, w, stride=1, padding=0, act=ACT_NONE):
    return ConvTranspose2dAct.apply(x, w, stride, padding, act)


def max_pool2(x):
    return MaxPool2.apply(x)


def bce_with_logits(logits, target):
    return BCEWithLogits.apply(logits, target)


def adam_step(params, grads, exp_avgs, exp_avg_sqs, step, lr, beta1, beta2, eps, grad_scale=1.0):
    """One multi-tensor Adam launch over fp32 CUDA tensors (all updated in place)."""
    n = len(params)
    if n == 0:
        return
    for t in (*params, *grads, *exp_avgs, *exp_avg_sqs):
        N.require_cuda(t)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise N.NativeError("adam_step needs contiguous fp32 tensors")
    ptrs = (ctypes.c_void_p * (4 * n))(*[t.data_ptr() for t in (*params, *grads, *exp_avgs, *exp_avg_sqs)])
    sizes = (ctypes.c_int64 * n)(*[p.numel() for p in params])
    N.call("mmg_adam_multi_tensor_f32", n, ptrs, sizes, lr, beta1, beta2, eps, step, grad_scale, N.stream())

"""bf16 tensor-core execution of DiscriminatorCNN (forward + backward) over the C ABI.

``DiscTC`` wraps a ``DiscriminatorCNN`` module (fp32 master parameters, the reference's state-dict keys):
it owns the packed bf16 weights and the activation buffers in the padded space-to-depth layout described
in csrc/disc_tc.cu, and exposes ``forward(x) -> logits`` / ``backward(dlogit)`` which accumulates the six
parameter gradients into ``param.grad`` (fp32), exactly where ``loss.backward()`` would put them.
"""
import torch

from . import _native as N

ROWS = 429         # conv1-activation / conv2 row space: 33 x 13 super pixels per sample
XROWS = 1690       # input / conv1-output row space: 65 x 26 super pixels per sample
_XD = {torch.float32: 0, torch.uint8: 2}


class DiscTC:
    def __init__(self, disc, max_batch, roll_size=(2, 128, 50), fused_backward=True, fused_forward=True):
        self.fused_backward, self.fused_forward = fused_backward, fused_forward
        if tuple(roll_size) != (2, 128, 50) or disc.conv1.weight.shape != (16, 2, 4, 4) or disc.conv2.weight.shape != (32, 16, 4, 4):
            raise ValueError("the tensor-core discriminator path is specialised to roll_size (2,128,50), hidden_dim 16")
        self.d = disc
        dev = disc.conv1.weight.device
        N.require_cuda(disc.conv1.weight)
        self.dev, self.cap = dev, int(max_batch)
        self.packed = torch.zeros(N.lib().mmg_disc_packed_weights_bytes(), dtype=torch.uint8, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.p1 = torch.zeros(self.cap * ROWS, 64, **bf)            # pad cells stay zero forever
        self.a2 = torch.empty(self.cap * ROWS, 32, **bf)
        self.dz2 = torch.empty(self.cap * ROWS, 32, **bf)
        self.xs = torch.empty(self.cap * XROWS, 8, **bf)
        self.dz1c = torch.zeros(self.cap * XROWS, 16, **bf)         # junk rows stay zero forever
        self.logits = torch.empty(self.cap, device=dev)
        self.x, self.B = None, 0
        self.ws = torch.zeros(N.lib().mmg_disc_pass_workspace_bytes(), dtype=torch.uint8, device=dev)     # per-CTA scratch of the one-kernel pass + index flag
        self.pack()

    def pack(self):
        """Re-derive the operand layouts from the fp32 master weights (call after every optimiser step)."""
        d = self.d
        N.call("mmg_disc_pack_weights", N.ptr(d.conv1.weight.data), N.ptr(d.conv2.weight.data), N.ptr(d.fc.weight.data), N.ptr(self.packed), N.stream())

    def forward(self, x, index=None, bufs=None):
        """x: (B,2,128,50) uint8 or float32 CUDA tensor -> logits (B,) fp32 (a view of an internal buffer).
        ``index`` (B,) int64 CUDA tensor: the pass runs on rows ``x[index]`` of a larger resident set, gathered inside the kernel.
        ``bufs`` = (xs, p1, a2, logits): activation buffers of this call instead of the internal ones (fused forward only; p1 zero-initialised)."""
        N.require_cuda(x, index)
        B = x.shape[0] if index is None else index.numel()
        if B > self.cap or tuple(x.shape[1:]) != (2, 128, 50) or x.dtype not in _XD:
            raise ValueError(f"bad discriminator input {tuple(x.shape)} {x.dtype} (capacity {self.cap})")
        if index is not None and (index.dtype != torch.int64 or not self.fused_forward):
            raise ValueError("index must be an int64 tensor (fused forward only)")
        x = x.contiguous()
        d, s = self.d, N.stream()
        logits = self.logits[:B]
        if self.fused_forward:       # one persistent kernel: P1 stays in shared memory between conv1 and conv2 (csrc/disc_tc_fused.cu)
            xs, p1, a2 = (self.xs, self.p1, self.a2) if bufs is None else bufs[:3]
            if bufs is not None:
                logits = bufs[3]
            N.call("mmg_disc_fwd_fused_gather", N.ptr(x), _XD[x.dtype], N.ptr(index), N.ptr(self.packed), N.ptr(d.conv1.bias.data), N.ptr(d.conv2.bias.data),
                   N.ptr(d.fc.bias.data), N.ptr(xs), N.ptr(p1), N.ptr(a2), N.ptr(logits), B, s)
            self.x, self.B = x, B
            return logits
        N.call("mmg_fill_scalar_f32", N.ptr(logits), N.ptr(d.fc.bias.data), B, s)
        N.call("mmg_disc_xs_pack", N.ptr(x), _XD[x.dtype], N.ptr(self.xs), B, s)
        N.call("mmg_disc_conv1_fwd", N.ptr(self.xs), N.ptr(self.packed), N.ptr(d.conv1.bias.data), N.ptr(self.p1), B, s)
        N.call("mmg_disc_conv2_fwd", N.ptr(self.p1), N.ptr(self.packed), N.ptr(d.conv2.bias.data), N.ptr(self.a2), N.ptr(logits), B, s)
        self.x, self.B = x, B
        return logits

    def pass_fused(self, x, target, loss=None, index=None, loss_rows=0, want_logits=True, dbg=None):
        """forward + BCE-with-logits against the constant ``target`` + backward of one batch in ONE kernel (csrc/disc_tc_pass.cu):
        the six gradients are accumulated into ``param.grad``, ``loss[0] +=`` the mean BCE over ``loss_rows`` rows (0 = this batch),
        the logits (B,) are returned (a view of an internal buffer) when ``want_logits``.  Nothing but the rolls is read from HBM."""
        N.require_cuda(x, index, loss)
        B = x.shape[0] if index is None else index.numel()
        if B > self.cap or tuple(x.shape[1:]) != (2, 128, 50) or x.dtype not in _XD:
            raise ValueError(f"bad discriminator input {tuple(x.shape)} {x.dtype} (capacity {self.cap})")
        if index is not None and index.dtype != torch.int64:
            raise ValueError("index must be an int64 tensor")
        x = x.contiguous()
        d = self.d
        g = {k: self._grad(p) for k, p in d.named_parameters()}
        logits = self.logits[:B] if want_logits else None
        args = (N.ptr(x), _XD[x.dtype], N.ptr(index), x.shape[0], N.ptr(self.packed), N.ptr(d.conv1.bias.data), N.ptr(d.conv2.bias.data), N.ptr(d.fc.bias.data),
                float(target), int(loss_rows), N.ptr(logits), N.ptr(loss), N.ptr(g["conv1.weight"]), N.ptr(g["conv1.bias"]), N.ptr(g["conv2.weight"]),
                N.ptr(g["conv2.bias"]), N.ptr(g["fc.weight"]), N.ptr(g["fc.bias"]), N.ptr(self.ws), self.ws.numel(), B, N.stream())
        if dbg is None:
            N.call("mmg_disc_pass_fused", *args)
        else:
            N.call("mmg_disc_pass_fused_dbg", *args, dbg)
        self.x, self.B = x, B
        return logits

    def index_out_of_range(self):
        """True if a gathered pass met a row index outside the resident set since the last call (synchronises; the row was read as row 0)."""
        flag = self.ws[-128:-124].view(torch.int32)
        bad = bool(flag.item())
        flag.zero_()
        return bad

    def _grad(self, p):
        if p.grad is None:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return p.grad

    def backward(self, dlogit, bufs=None):
        """dlogit: (B,) fp32 = dLoss/dlogits of the last forward (or of the forward that filled ``bufs``).  Accumulates into the six ``.grad`` tensors."""
        B, d, s = (self.B if bufs is None else dlogit.numel()), self.d, N.stream()
        dlogit = dlogit.contiguous()
        g = {k: self._grad(p) for k, p in d.named_parameters()}
        if self.fused_backward:      # one persistent kernel: dz2 / dz1c never leave the SM (csrc/disc_tc_fused.cu)
            xs, p1, a2 = (self.xs, self.p1, self.a2) if bufs is None else bufs[:3]
            N.call("mmg_disc_bwd_fused", N.ptr(xs), N.ptr(p1), N.ptr(a2), N.ptr(dlogit), N.ptr(self.packed), N.ptr(g["conv1.weight"]),
                   N.ptr(g["conv1.bias"]), N.ptr(g["conv2.weight"]), N.ptr(g["conv2.bias"]), N.ptr(g["fc.weight"]), N.ptr(g["fc.bias"]), B, s)
            return
        N.call("mmg_sum_f32", N.ptr(dlogit), B, N.ptr(g["fc.bias"]), 1, s)
        N.call("mmg_disc_fc_bwd", N.ptr(self.a2), N.ptr(dlogit), N.ptr(self.packed), N.ptr(self.dz2), N.ptr(g["fc.weight"]), N.ptr(g["conv2.bias"]), B, s)
        N.call("mmg_disc_conv2_wgrad", N.ptr(self.p1), N.ptr(self.dz2), N.ptr(g["conv2.weight"]), B, s)
        N.call("mmg_disc_conv2_dgrad", N.ptr(self.dz2), N.ptr(self.packed), N.ptr(self.p1), N.ptr(self.dz1c), N.ptr(g["conv1.bias"]), B, s)
        N.call("mmg_disc_conv1_wgrad", N.ptr(self.xs), N.ptr(self.dz1c), N.ptr(g["conv1.weight"]), B, s)


class DiscTCFunction(torch.autograd.Function):
    """``DiscriminatorCNN.forward`` on the tensor-core kernels as an autograd node (``DiscriminatorCNN.enable_tensor_cores``): forward = the fused
    forward kernel, backward = the fused backward kernel; every call owns its activation buffers (the reference calls the discriminator on the fake and
    on the real batch before one backward, network_tests.py:294-307), parameter gradients are returned to autograd."""

    @staticmethod
    def forward(ctx, x, tc, w1, b1, w2, b2, wf, bf):
        B = x.shape[0]
        bfk = dict(dtype=torch.bfloat16, device=x.device)
        bufs = (torch.empty(B * XROWS, 8, **bfk), torch.zeros(B * ROWS, 64, **bfk), torch.empty(B * ROWS, 32, **bfk), torch.empty(B, device=x.device))
        with torch.no_grad():
            tc.pack()                                   # the fp32 master weights may have changed since the last call
            logits = tc.forward(x, bufs=bufs)
        ctx.tc, ctx.bufs = tc, bufs
        return logits.view(-1, 1)

    @staticmethod
    def backward(ctx, dlogits):
        tc = ctx.tc
        d = tc.d
        saved = {k: p.grad for k, p in d.named_parameters()}
        with torch.no_grad():
            for p in d.parameters():
                p.grad = None
            tc.backward(dlogits.reshape(-1).float().contiguous(), bufs=ctx.bufs)      # accumulates into fresh zeroed .grad tensors
            grads = [p.grad for p in (d.conv1.weight, d.conv1.bias, d.conv2.weight, d.conv2.bias, d.fc.weight, d.fc.bias)]
            for k, p in d.named_parameters():
                p.grad = saved[k]
        return (None, None, *grads)

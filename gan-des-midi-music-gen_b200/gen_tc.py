"""bf16 tensor-core execution of the MM-GAN generators (forward, training or eval mode) over the C ABI.

``GenTC`` wraps a ``Generator`` / ``BeatGenerator`` module (fp32 master parameters, the reference's state-dict keys,
network_tests.py:58-123): it owns the packed bf16 weights, the fp32 pre-activation buffers of the hidden layers and the
fp64 batch-sum buffers, and runs the 4-block stack as 5 launches of ``mmg_gen_layer_fwd`` (csrc/gen_tc.cu): one per
layer, the last layer twice (batch sums, then normalise + sigmoid + the single write of the output).  Training mode
updates ``running_mean`` / ``running_var`` / ``num_batches_tracked`` in place like ``nn.BatchNorm1d``.
Data parallel: ``sync_bn=True`` with an initialised process group all-reduces every layer's fp64 column sums between the
layer kernels (4 collectives of <= 64 KB per forward) and normalises with the GLOBAL batch count, so that the sharded run
reproduces the single-process global-batch BatchNorm of the reference; default = per-replica statistics.
Forward only: the reference never back-propagates into the generators; the fp32 modules keep the
differentiable path.
"""
import ctypes

import torch

from . import _native as N


class GenTC:
    def __init__(self, gen, max_batch, process_group=None, sync_bn=False, gram_stats=True, sum_views=None, fused_hidden=True):
        self.g = gen
        dist = torch.distributed
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (sync_bn and dist.is_available() and dist.is_initialized()) else 1
        self.sync_bn = bool(sync_bn) and self.world > 1
        self.blocks = [(blk[0], blk[1]) for blk in gen.gen]            # (Linear, BatchNorm1d)
        dev = self.blocks[0][0].weight.device
        N.require_cuda(self.blocks[0][0].weight)
        self.dev, self.cap = dev, int(max_batch)
        self.widths = [lin.out_features for lin, _ in self.blocks]
        if self.blocks[0][0].in_features > 256 or any(w > 256 for w in self.widths[:-1]):
            raise ValueError("the tensor-core generator path supports at most 256 input features per layer")
        self.packed = [torch.empty(N.lib().mmg_gen_packed_weight_bytes(lin.out_features, lin.in_features), dtype=torch.uint8, device=dev)
                       for lin, _ in self.blocks]
        self._versions = None
        self.z = [torch.empty(self.cap, w, device=dev) for w in self.widths[:-1]]
        offs, o = [], 0
        for w in self.widths:
            offs.append(o)
            o += 2 * w
        if sum_views is None:
            self.sums = torch.zeros(o, dtype=torch.float64, device=dev)     # [layer][sum | sumsq][width]
            self.sum_views = [self.sums[a:a + 2 * w] for a, w in zip(offs, self.widths)]
        else:           # storage owned by the caller (trainer.py: the two generators' layer-i sums sit side by side: one all-reduce for both)
            assert len(sum_views) == len(self.widths) and all(v.numel() == 2 * w and v.dtype == torch.float64 for v, w in zip(sum_views, self.widths))
            self.sums, self.sum_views = None, list(sum_views)
        # the wide output layer's batch statistics come from the 64 x 64 Gram matrix of its input instead of a GEMM pass (csrc/gen_tc.cu)
        last = self.blocks[-1][0]
        self.gram_stats = bool(gram_stats) and last.in_features <= 64 and last.in_features % 8 == 0 and last.out_features >= 512
        self.gram_ws = torch.empty(N.lib().mmg_gen_layer_stats_gram_workspace(), dtype=torch.uint8, device=dev) if self.gram_stats else None
        # train mode with local statistics: the three hidden blocks run as ONE cooperative launch (mmg_gen_hidden_fused) when the batch fits
        self.fused_hidden = bool(fused_hidden) and len(self.blocks) == 4 and not self.sync_bn
        self.barrier = torch.zeros(2, dtype=torch.int32, device=dev) if self.fused_hidden else None
        self._fused_ok = {}
        self.pack()

    def _fused_supported(self, B):
        ok = self._fused_ok.get(B)
        if ok is None:
            n3 = (ctypes.c_int * 3)(*self.widths[:3])
            ok = self._fused_ok[B] = bool(N.lib().mmg_gen_hidden_fused_supported(B, self.blocks[0][0].in_features, n3, int(self.gram_stats)))
        # one momentum / eps for the three hidden BatchNorm1d layers (the reference's are the nn defaults); checked per call, they are plain attributes
        bns = [bn for _, bn in self.blocks[:3]]
        return ok and all(bn.momentum == bns[0].momentum and bn.eps == bns[0].eps for bn in bns)

    def pack(self, force=True):
        """Re-derive the bf16 operand copies from the fp32 master weights (cheap; skipped when nothing changed)."""
        vers = tuple(lin.weight._version for lin, _ in self.blocks)
        if not force and vers == self._versions:
            return
        for (lin, _), pk in zip(self.blocks, self.packed):
            N.call("mmg_gen_pack_weight", N.ptr(lin.weight.data), lin.out_features, lin.in_features, N.ptr(pk), N.stream())
        self._versions = vers

    def _sync(self, i):
        """SyncBN: sum layer i's batch sums over the ranks (every rank holds an equal shard)."""
        if self.sync_bn:
            torch.distributed.all_reduce(self.sum_views[i], group=self.pg)

    def forward(self, noise, input_tensor, training=None, out=None):
        """noise (B,z) and input_tensor (B,input_dim) fp32 CUDA tensors -> (B, out_features) fp32 (caller reshapes)."""
        steps = self.forward_steps(noise, input_tensor, training, out)
        while True:
            try:
                i = next(steps)
            except StopIteration as done:
                return done.value
            self._sync(i)

    def forward_steps(self, noise, input_tensor, training=None, out=None):
        """The forward as a Python generator: yields the index of a layer whose batch sums are complete on this rank and must be summed over
        the ranks (SyncBN) before the next kernel is enqueued; returns the output.  ``forward`` all-reduces each one on its own; the trainer
        drives both generators in lock step and reduces their layer-i sums with ONE collective."""
        N.require_cuda(noise, input_tensor)
        training = self.g.training if training is None else training
        B = noise.shape[0]
        if B > self.cap:
            raise ValueError(f"batch {B} exceeds the preallocated capacity {self.cap}")
        if training and B * self.world <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {(B, self.widths[0])}")
        noise, input_tensor = noise.float().contiguous(), input_tensor.float().contiguous()
        self.pack(force=False)
        s = N.stream()
        if training:
            if self.sums is not None:
                N.call("mmg_zero", N.ptr(self.sums), self.sums.numel() * 8, s)
            else:
                for v in self.sum_views:
                    N.call("mmg_zero", N.ptr(v), v.numel() * 8, s)
        y = out if out is not None else torch.empty(B, self.widths[-1], device=self.dev)
        mode = 1 if training else 2
        last = len(self.blocks) - 1
        fused = training and self.fused_hidden and self._fused_supported(B)
        if fused:
            N.call("mmg_zero", N.ptr(self.barrier), 8, s)
            h = N.GenHiddenArgs()
            h.x0, h.k0, h.x1, h.k1 = N.ptr(noise), noise.shape[1], N.ptr(input_tensor), input_tensor.shape[1]
            for l in range(3):
                lin, bn = self.blocks[l]
                h.w_packed[l], h.bias[l], h.N[l] = N.ptr(self.packed[l]), N.ptr(lin.bias.data), lin.out_features
                h.gamma[l], h.beta[l], h.run_mean[l], h.run_var[l] = N.ptr(bn.weight.data), N.ptr(bn.bias.data), N.ptr(bn.running_mean), N.ptr(bn.running_var)
                h.sums[l] = N.ptr(self.sum_views[l])
            bn0 = self.blocks[0][1]             # (the reference's blocks share momentum / eps: nn.BatchNorm1d defaults)
            h.z_out, h.gram_part, h.barrier = N.ptr(self.z[2]), (N.ptr(self.gram_ws) if self.gram_stats else None), N.ptr(self.barrier)
            h.momentum, h.eps, h.update_running, h.M, h.stat_count = (bn0.momentum if bn0.momentum is not None else 0.1), bn0.eps, 1, B, 0
            N.call("mmg_gen_hidden_fused", ctypes.byref(h), s)
        for i, (lin, bn) in enumerate(self.blocks):
            if fused and i < last:
                continue
            a = N.GenLayerArgs()
            if i == 0:
                a.x0, a.k0, a.x1, a.k1, a.in_mode = N.ptr(noise), noise.shape[1], N.ptr(input_tensor), input_tensor.shape[1], 0
            else:
                pbn = self.blocks[i - 1][1]
                a.x0, a.k0, a.x1, a.k1, a.in_mode = N.ptr(self.z[i - 1]), self.widths[i - 1], None, 0, mode
                a.in_sums = N.ptr(self.sum_views[i - 1]) if training else None
                a.in_gamma, a.in_beta = N.ptr(pbn.weight.data), N.ptr(pbn.bias.data)
                a.in_run_mean, a.in_run_var = N.ptr(pbn.running_mean), N.ptr(pbn.running_var)
                if fused:                      # the fused launch has made the one update of the hidden layers' running statistics
                    a.in_run_mean = a.in_run_var = None
            a.w_packed, a.bias, a.N = N.ptr(self.packed[i]), N.ptr(lin.bias.data), lin.out_features
            a.momentum = bn.momentum if bn.momentum is not None else 0.1
            a.eps, a.M = bn.eps, B
            a.stat_count = B * self.world if training else 0
            a.update_running = int(training)
            if i < last:
                a.z_out = N.ptr(self.z[i])
                a.out_sums = N.ptr(self.sum_views[i]) if training else None
                N.call("mmg_gen_layer_fwd", ctypes.byref(a), s)
                if training and self.sync_bn:
                    yield i
            else:
                a.out_gamma, a.out_beta = N.ptr(bn.weight.data), N.ptr(bn.bias.data)
                a.out_run_mean, a.out_run_var = N.ptr(bn.running_mean), N.ptr(bn.running_var)
                if training and self.gram_stats and fused:      # the fused launch left the per-CTA Gram partials: reduce + column statistics
                    N.call("mmg_gen_layer_stats_gram_finish", (B + 127) // 128, N.ptr(lin.weight.data), N.ptr(lin.bias.data), lin.out_features,
                           self.widths[i - 1], B, N.ptr(self.sum_views[i]), N.ptr(self.gram_ws), self.gram_ws.numel(), s)
                elif training and self.gram_stats:   # batch sums of this layer from the Gram matrix of its input: no GEMM pass for the statistics
                    pbn = self.blocks[i - 1][1]
                    N.call("mmg_gen_layer_stats_gram", N.ptr(self.z[i - 1]), B, self.widths[i - 1], N.ptr(self.sum_views[i - 1]), B * self.world,
                           N.ptr(pbn.weight.data), N.ptr(pbn.bias.data), pbn.eps, N.ptr(lin.weight.data), N.ptr(lin.bias.data), lin.out_features,
                           N.ptr(self.sum_views[i]), N.ptr(self.gram_ws), self.gram_ws.numel(), s)
                    if self.sync_bn:
                        yield i                    # (the single pass below also makes the one update of the previous layer's running stats)
                elif training:                     # pass 1: batch sums only (also the one update of the previous layer's running stats)
                    a.out_sums = N.ptr(self.sum_views[i])
                    N.call("mmg_gen_layer_fwd", ctypes.byref(a), s)
                    if self.sync_bn:
                        yield i
                    a.out_sums = None
                    a.in_run_mean = a.in_run_var = None      # already updated by pass 1
                a.y_out, a.out_mode = N.ptr(y), mode
                a.y_sums = N.ptr(self.sum_views[i]) if training else None
                N.call("mmg_gen_layer_fwd", ctypes.byref(a), s)
        if training:
            torch._foreach_add_([bn.num_batches_tracked for _, bn in self.blocks], 1)
        return y

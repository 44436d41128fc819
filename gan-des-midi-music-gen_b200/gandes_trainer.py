"""The GAN-DES loop body (/root/reference/GAN_DES/SIMNN.py:275-334) as replayable segments around the host bridge.

The reference loop is ``real -> disc`` / ``gen(noise) -> matrix_to_wav (host DES + FluidSynth) -> disc(fake)`` / backward / ``disc_opt.step()``,
then ``disc(fake)`` against ones / backward / ``gen_opt.step()`` (a no-op: the fake spectrograms come back from the host, no generator parameter
has a gradient).  ``GANDESTrainer`` keeps that order and cuts it where the host takes over:

    fake_matrices = tr.generate(noise)            # SIMNN.py:291-294   (eval of the generator under no_grad, as the loop detaches it at once)
    fake = matrix_to_wav(fake_matrices ...)       # host, unchanged
    d_loss = tr.d_step(real, fake)                # SIMNN.py:279-316
    g_loss = tr.g_step(fake)                      # SIMNN.py:321-331

Each segment runs eagerly the first time it sees a set of input buffers, is captured into a CUDA graph the second time (forward, autograd
backward and the Adam update: ~60 launches) and replayed afterwards.  Adam's step count and hyper-parameters live in device memory
(``mmg_adam_multi_tensor_dev_f32``) so the replayed update is the right one.  Works with the fp32 kernels and with ``enable_tensor_cores()``.

Data parallel: with an initialised process group every rank runs the segments on its shard of the batch.  The discriminator has no
BatchNorm, so the mean-loss gradient of the global batch is the average of the shard gradients: each parameter's gradient is all-reduced as
soon as autograd has produced it (fc2 / fc1 -- 28 MB -- first, underneath the convolution backward), the 1/world goes into the Adam kernel,
and with NCCL the collectives are captured into the D-step graph.  The G step's discriminator gradients are never applied by the reference and
are not reduced.  ``generate`` uses per-replica BatchNorm statistics; the returned losses are the means over this rank's shard.
"""
import ctypes

import torch

from . import _native as N
from . import functional_tc as FnTC
from .optim import BCEWithLogitsLoss


class GANDESTrainer:
    def __init__(self, gen, disc, lr=2e-5, betas=(0.5, 0.999), eps=1e-8, use_graph=True, process_group=None, data_parallel=None, batch_d_passes=True):
        """``data_parallel``: None = shard over ``process_group`` when torch.distributed is initialised, False = this process alone.
        ``batch_d_passes``: the D step's real and fake batches go through the discriminator as one batch of 2B (same results; False = two passes)."""
        self.gen, self.disc = gen, disc
        self.batch_d_passes = bool(batch_d_passes)
        dist = torch.distributed
        self.pg = process_group
        on = dist.is_available() and dist.is_initialized() and data_parallel is not False
        self.world = dist.get_world_size(process_group) if on else 1
        self._reduce_now, self._forked = False, False
        self.crit = BCEWithLogitsLoss()
        self.d_params = list(disc.parameters())
        dev = self.d_params[0].device
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps], dtype=torch.float32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=dev)       # updates applied so far (device side: graph replays advance it)
        self.exp_avg = [torch.zeros_like(p) for p in self.d_params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.d_params]
        self.use_graph = bool(use_graph)
        self._graphs, self._seen, self._labels = {}, {}, {}
        self.replayed_launches = 0
        if self.world > 1:
            if self.use_graph and dist.get_backend(process_group) != "nccl":
                self.use_graph = False                        # host-side collectives (gloo) cannot be captured
            self._side = torch.cuda.Stream(device=dev)
            for p in self.d_params:
                p.register_post_accumulate_grad_hook(self._grad_ready)

    def _grad_ready(self, p):
        """autograd has finished this parameter's gradient: its all-reduce (D step only) goes to a side stream, so the rest of the backward
        runs underneath it (fork here, join before Adam: the same pattern eagerly and under graph capture; no asynchronous work handles,
        which the NCCL watchdog would keep waiting for after a capture)"""
        if self._reduce_now:
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                torch.distributed.all_reduce(p.grad, group=self.pg)
            self._forked = True

    # ------------------------------------------------------------------ pieces
    def _label(self, B, v):
        t = self._labels.get((B, v))
        if t is None:
            t = self._labels[(B, v)] = torch.full((B,), v, device=self.hyper.device)
        return t

    def _adam(self):
        ps = self.d_params
        ts = [p.data for p in ps] + [p.grad for p in ps] + self.exp_avg + self.exp_avg_sq
        ptrs = (ctypes.c_void_p * (4 * len(ps)))(*[N.ptr(t) for t in ts])
        sizes = (ctypes.c_int64 * len(ps))(*[p.numel() for p in ps])
        N.call("mmg_adam_multi_tensor_dev_f32", len(ps), ptrs, sizes, N.ptr(self.hyper), N.ptr(self.adam_step), 1.0 / self.world, N.stream())
        FnTC.invalidate_weight_cache()

    def _d_body(self, real, fake):
        B = real.shape[0]
        for p in self.d_params:
            p.grad = None
        if self.batch_d_passes:
            # the discriminator has no BatchNorm: D(real) and D(fake) are per-sample functions, so ONE forward / backward over the 2B samples
            # gives the same logits and the same gradients as the reference's two (SIMNN.py:279-313) with half the launches and one read of the
            # 28 MB fc1 weight instead of two.  mean over B of each half, summed = 2 x the mean over 2B.
            key = ("d_targets", B)
            if key not in self._labels:
                self._labels[key] = torch.cat([self._label(B, 0.9), self._label(B, 0.1)])
            loss = 2.0 * self.crit(self.disc(torch.cat([real, fake.detach()])).reshape(-1), self._labels[key])
        else:
            l_real = self.crit(self.disc(real).reshape(-1), self._label(B, 0.9))
            l_fake = self.crit(self.disc(fake.detach()).reshape(-1), self._label(B, 0.1))
            loss = l_fake + l_real
        self._reduce_now = self.world > 1
        try:
            loss.backward()
        finally:
            self._reduce_now = False
        if self._forked:                                       # sums over the ranks are complete before Adam; the 1/world is applied inside its kernel
            torch.cuda.current_stream().wait_stream(self._side)
            self._forked = False
        for p in self.d_params:
            if p.grad is None or not p.grad.is_contiguous():
                raise RuntimeError("GANDESTrainer: every discriminator parameter needs a contiguous gradient")
        self._adam()
        return loss.detach()

    def _g_body(self, fake):
        B = fake.shape[0]
        for p in self.d_params:            # gen_opt.zero_grad() leaves the discriminator's gradients; they are recomputed here and never applied
            p.grad = None
        loss = self.crit(self.disc(fake).squeeze(), self._label(B, 1.0))
        loss.backward()
        return loss.detach()

    def _segment(self, name, fn, tensors):
        """eager on first sight of these buffers, captured on the second call, replayed afterwards"""
        if not self.use_graph:
            return fn()
        key = (name,) + tuple((t.data_ptr(), tuple(t.shape), t.dtype) for t in tensors)
        g = self._graphs.get(key)
        if g is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] == 1:
                return fn()
            torch.cuda.synchronize()
            FnTC.invalidate_weight_cache()                 # every packed weight must be (re)built INSIDE the graph
            l0 = N.lib().mmg_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                out = fn()
            FnTC.invalidate_weight_cache()                 # the cache now points into the graph's private pool: eager code must not trust it
            g = self._graphs[key] = (graph, out, N.lib().mmg_launch_count() - l0)
        g[0].replay()
        self.replayed_launches += g[2]
        return g[1]

    # ------------------------------------------------------------------ public segments
    def generate(self, noise):
        """SIMNN.py:291-297: the generator's forward for the host bridge (train-mode BatchNorm, running statistics advance); no autograd tape."""
        def body():
            with torch.no_grad():
                return self.gen(noise)
        return self._segment("gen", body, (noise,))

    def d_step(self, real, fake):
        """SIMNN.py:279-316.  Returns the discriminator loss (device scalar; under graph replay the same tensor every call)."""
        return self._segment("d", lambda: self._d_body(real, fake), (real, fake))

    def g_step(self, fake):
        """SIMNN.py:321-331."""
        return self._segment("g", lambda: self._g_body(fake), (fake,))

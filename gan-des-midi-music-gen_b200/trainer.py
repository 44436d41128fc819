"""Fused MM-GAN training iteration (the reference loop body, network_tests.py:292-315) behind one call.

``MMGANTrainer.step`` performs, for one batch:
    D step: G1/G2 train-mode forward (BN running stats update), D forward on the fake and the real
            rolls, BCE (fake vs 0, real vs 1), backward, Adam on the 6 discriminator tensors;
    G step: G1/G2 forward again, D forward on the G-step fake rolls, BCE vs 1, backward (its D grads
            accumulate on top of the D-step grads, as in the reference, and are discarded by the next
            step), generator Adam is a no-op because no generator parameter has a grad.
The fake rolls are what the host DES bridge returned for the two generator forwards; here they are
inputs.  Rolls may be float32 or uint8 (B,2,128,W) tensors (piano-roll values are integers 0..127 and
durations < 256, so uint8 is exact and quarters the H2D / HBM traffic).

Public segments (the reference's own dependency G -> host DES -> D, network_tests.py:177-196, expressed through the API): ``generators`` runs one
pair of G1 / G2 train-mode forwards and returns the matrices the host bridge consumes; ``d_step(real, fake_d)`` and ``g_step(fake_g)`` then
take the rolls the bridge made from them.  ``step`` = generators, d_step, generators, g_step on given rolls (bench / parity: the DES is
excluded there, so its rolls are inputs that by construction do not depend on this step's generator outputs).

precision='fp32' : the drop-in nn.Modules + autograd over the fp32 SIMT kernels (reference tolerance).
precision='bf16' : discriminator on the tcgen05 tensor-core kernels (disc_tc.DiscTC): bf16 operands,
                   fp32 accumulation, fp32 master weights / Adam state; fused BCE and multi-tensor Adam.
CUDA graphs (``use_graph=True``, default for bf16): after one eager iteration on the same input buffers the iteration
is captured (one graph, or four around the asynchronous NCCL all-reduce when sharded) and later calls with the same buffers replay
it: one launch per iteration instead of ~100.  Adam's step count and learning rate live in device memory for that
(``mmg_adam_multi_tensor_dev_f32``), so ``StepLR`` keeps working.
``inner_rng``: the reference's Generator draws its second input with ``torch.randn`` on the CPU generator inside forward
(network_tests.py:83-84); 'reference' reproduces that draw (same global RNG consumption), 'device' draws on the GPU.
Data parallel: with torch.distributed initialised (NCCL) the batch is sharded by rank, the D gradients
live in ONE flat fp32 buffer (84 KB) that is all-reduced once per optimiser step, asynchronously, overlapped with the G step's
generator forwards (the 1/world factor is folded into the Adam kernel); the G-step D grads are not reduced (the reference discards them).
``sync_bn=True`` (bf16 path): the generators' train-mode BatchNorm uses GLOBAL-batch statistics (fp64 column sums all-reduced between the
layer kernels, gen_tc.GenTC), reproducing the reference's single-process batch; default False = per-replica statistics (torch-DDP semantics).

Losses under data parallelism: ``step`` / ``d_step`` / ``g_step`` return the mean over THIS RANK's shard (the gradients are the global mean);
the reference's global-batch loss is the average of the per-rank values: ``dist.all_reduce(loss); loss /= world`` (tests/test_gpu_dp_nccl.py).
They are left per-rank so that no collective sits between an iteration and the next one's launch.
"""
import ctypes

import torch

from . import _native as N
from . import functional as Fn
from .optim import FusedAdam


def shard_batch(t, rank, world):
    """Rank's contiguous slice of a global batch tensor (dim 0); the global batch must divide evenly."""
    n = t.shape[0]
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return t[rank * per:(rank + 1) * per]


class MMGANTrainer:
    def __init__(self, mmgan, lr=0.01, betas=(0.9, 0.999), eps=1e-8, precision="fp32", max_batch=None, process_group=None, use_graph=None,
                 inner_rng="reference", sync_bn=False, one_kernel_pass=True, capture_collectives=None):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.m = mmgan
        self.precision = precision
        D = mmgan.discriminator
        self.d_params = list(D.parameters())
        # one flat gradient buffer; every p.grad is a view into it (single memset / all-reduce / Adam launch)
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.d_params), device=self.d_params[0].device)
        o = 0
        for p in self.d_params:
            p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()
        self.disc_opt = FusedAdam(self.d_params, lr=lr, betas=betas, eps=eps)
        for p in self.d_params:                       # state exists from the start (the graph bakes these addresses in)
            self.disc_opt.state[p].update(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
        self.gen_opt = FusedAdam(list(mmgan.generator1.parameters()) + list(mmgan.generator2.parameters()), lr=lr, betas=betas, eps=eps)
        self.pg = process_group
        dist = torch.distributed
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.on_d_grads = None
        self.sync_bn = bool(sync_bn) and self.world > 1
        # SyncBN's statistics all-reduces sit between the generator layer kernels: with NCCL they are captured into the CUDA graphs like any
        # kernel (one replay per segment); gloo collectives are host calls, so those generator forwards stay eager
        self.capture_collectives = (self.world > 1 and dist.get_backend(process_group) == "nccl") if capture_collectives is None else bool(capture_collectives)
        if self.sync_bn and precision != "bf16":
            raise ValueError("sync_bn is implemented on the bf16 tensor-core generator path (precision='bf16')")
        self.tc = None
        self._g_out_owned = True
        self.one_kernel_pass = bool(one_kernel_pass)      # bf16 path: mmg_disc_pass_fused (False: forward kernel + BCE kernel + backward kernel)
        self._g_out = None
        self._side = torch.cuda.Stream(device=self.d_params[0].device) if precision == "bf16" else None
        if precision == "bf16":
            if max_batch is None:
                raise ValueError("precision='bf16' needs max_batch (activation buffers are preallocated)")
            from .disc_tc import DiscTC
            from .gen_tc import GenTC
            self.tc = DiscTC(D, max_batch)
            sv1 = sv2 = None
            if self.sync_bn:
                # SyncBN: the layer-i batch sums of the two generators sit side by side in one buffer, so ONE all-reduce per layer index serves
                # both (5 collectives per pair of forwards instead of 9)
                w1 = [blk[0].out_features for blk in mmgan.generator1.gen]
                w2 = [blk[0].out_features for blk in mmgan.generator2.gen]
                nl = max(len(w1), len(w2))
                tot = sum(2 * (w1[i] if i < len(w1) else 0) + 2 * (w2[i] if i < len(w2) else 0) for i in range(nl))
                self._sync_sums = torch.zeros(tot, dtype=torch.float64, device=self.flat_grad.device)
                sv1, sv2, self._sync_pairs, o = [], [], [], 0
                for i in range(nl):
                    a = o
                    if i < len(w1):
                        sv1.append(self._sync_sums[o:o + 2 * w1[i]]); o += 2 * w1[i]
                    if i < len(w2):
                        sv2.append(self._sync_sums[o:o + 2 * w2[i]]); o += 2 * w2[i]
                    self._sync_pairs.append(self._sync_sums[a:o])
            # the one-launch form of the hidden blocks is a cooperative kernel (grid-wide barriers): used by the single-GPU trainer only, so that it never
            # shares the device with an in-flight NCCL kernel of the gradient all-reduce
            fuse = self.world == 1
            self.gtc1 = GenTC(mmgan.generator1, max_batch, process_group, self.sync_bn, sum_views=sv1, fused_hidden=fuse)
            self.gtc2 = GenTC(mmgan.generator2, max_batch, process_group, self.sync_bn, sum_views=sv2, fused_hidden=fuse)
        dev = self.flat_grad.device
        if inner_rng not in ("reference", "device"):
            raise ValueError("inner_rng must be 'reference' or 'device'")
        self.inner_rng = inner_rng
        self.use_graph = (precision == "bf16") if use_graph is None else bool(use_graph)
        self._graphs, self._seg_graphs, self._seen, self._inner = {}, {}, {}, {}
        self.graph_launches = 0                      # kernels inside one captured iteration (for launch accounting)
        self.replayed_launches = 0                   # kernels launched through graph replays so far
        self.adam_hyper = torch.tensor([lr, betas[0], betas[1], eps], dtype=torch.float32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self._lr_cached = lr
        self.loss_d = torch.zeros(1, device=dev)
        self.loss_g = torch.zeros(1, device=dev)
        self.dlogit = torch.empty(max_batch or 1, device=dev)

    # ------------------------------------------------------------------ helpers
    def _zero_d_grads(self):
        N.call("mmg_zero", N.ptr(self.flat_grad), self.flat_grad.numel() * 4, N.stream())
        o = 0
        for p in self.d_params:       # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
            o += p.numel()

    def _allreduce_d_grads(self, async_op=False):
        if self.world > 1:                 # sum; 1/world is applied inside the Adam kernel
            return torch.distributed.all_reduce(self.flat_grad, group=self.pg, async_op=async_op)
        return None

    def _disc_adam(self):
        """Adam on the six discriminator tensors, one launch, hyper-parameters and step count read from device memory."""
        opt = self.disc_opt
        st = opt.state
        ps = self.d_params
        ts = [p.data for p in ps] + [p.grad for p in ps] + [st[p]["exp_avg"] for p in ps] + [st[p]["exp_avg_sq"] for p in ps]
        n = len(ps)
        ptrs = (ctypes.c_void_p * (4 * n))(*[N.ptr(t) for t in ts])
        sizes = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
        N.call("mmg_adam_multi_tensor_dev_f32", n, ptrs, sizes, N.ptr(self.adam_hyper), N.ptr(self.adam_step), 1.0 / self.world, N.stream())

    def _sync_hyper(self):
        """Host-side bookkeeping that a graph replay cannot do: push a changed lr (StepLR) to the device, count the step."""
        lr = self.disc_opt.param_groups[0]["lr"]
        if lr != self._lr_cached:
            self.adam_hyper[0:1].fill_(lr)
            self._lr_cached = lr
        for p in self.d_params:
            self.disc_opt.state[p]["step"] += 1

    def _draw_inner(self, B, which):
        """The Generator's in-forward ``randn`` (network_tests.py:83-84), into a static buffer (stable address for graph replay)."""
        dim = self.m.generator1.input_tensor_dim
        buf = self._inner.get((which, B))
        if buf is None:
            buf = self._inner[(which, B)] = torch.empty(B, dim, device=self.flat_grad.device)
        if self.inner_rng == "device":
            buf.normal_()
        else:
            buf.copy_(torch.randn(B, dim))
        return buf

    def _generators(self, noise1, noise2, beats, inner, out=None):
        m = self.m
        with torch.no_grad():
            if self.tc is not None:          # bf16 tcgen05 blocks (csrc/gen_tc.cu); outputs land in static buffers
                B = noise1.shape[0]
                if out is not None:
                    self._g_out = out
                elif self._g_out is None or self._g_out[0].shape[0] != B or self._g_out_owned is False:
                    self._g_out = (torch.empty(B, self.gtc1.widths[-1], device=noise1.device), torch.empty(B, self.gtc2.widths[-1], device=noise1.device))
                self._g_out_owned = out is None
                a = m.generator1.adj_size
                if self.sync_bn:
                    # SyncBN: the statistics all-reduces sit between the layer kernels.  Both generators advance in lock step on this stream
                    # and their layer-i sums (adjacent in self._sync_sums) are reduced by one collective.
                    st1 = self.gtc1.forward_steps(noise1, inner, out=self._g_out[0])
                    st2 = self.gtc2.forward_steps(noise2, beats, out=self._g_out[1])
                    y1 = y2 = None
                    i1 = i2 = -1
                    while y1 is None or y2 is None:
                        if y1 is None:
                            try:
                                i1 = next(st1)
                            except StopIteration as fin:
                                y1, i1 = fin.value, None
                        if y2 is None:
                            try:
                                i2 = next(st2)
                            except StopIteration as fin:
                                y2, i2 = fin.value, None
                        if i1 is not None and i2 is not None:
                            assert i1 == i2, "the generators' SyncBN points went out of step"
                            torch.distributed.all_reduce(self._sync_pairs[i1], group=self.pg)
                        elif i1 is not None:
                            torch.distributed.all_reduce(self.gtc1.sum_views[i1], group=self.pg)
                        elif i2 is not None:
                            torch.distributed.all_reduce(self.gtc2.sum_views[i2], group=self.pg)
                    self.g2_out = y2
                    self.g1_out = y1.view(B, -1, a[0], a[1])
                else:
                    # the two generators are independent: the beat generator runs on a side stream (fork / join, also under graph capture)
                    cur = torch.cuda.current_stream()
                    self._side.wait_stream(cur)
                    with torch.cuda.stream(self._side):
                        self.g2_out = self.gtc2.forward(noise2, beats, out=self._g_out[1])
                    self.g1_out = self.gtc1.forward(noise1, inner, out=self._g_out[0]).view(B, -1, a[0], a[1])
                    cur.wait_stream(self._side)
            else:
                self.g1_out = m.generator1(noise1, inner)
                self.g2_out = m.generator2(noise2, beats)

    def _d_pass(self, x, target, loss, accumulate, index=None):
        """forward + BCE + backward of the discriminator on one batch (rows ``x[index]`` when ``index`` is given); grads accumulate into flat_grad"""
        B = x.shape[0] if index is None else index.numel()
        if self.tc is not None and self.one_kernel_pass:
            # forward -> BCE -> backward per sample inside one persistent kernel (csrc/disc_tc_pass.cu): the activations never reach HBM
            if not accumulate:
                N.call("mmg_zero", N.ptr(loss), 4, N.stream())
            return self.tc.pass_fused(x, float(target), loss, index)
        if self.tc is not None:
            logits = self.tc.forward(x, index)
            dl = self.dlogit[:B]
            N.call("mmg_bce_logits_f32", N.ptr(logits), None, float(target), B, N.ptr(loss), int(accumulate), N.ptr(dl), 1.0 / B, None, N.stream())
            self.tc.backward(dl)
            return logits
        if index is not None:
            x = x.index_select(0, index)
        xf = x if x.dtype == torch.float32 else x.float()
        logits = self.m.discriminator(xf).squeeze(-1)
        lv = Fn.bce_with_logits(logits, float(target))
        lv.backward()
        if accumulate:
            loss += lv.detach()
        else:
            loss.copy_(lv.detach().reshape(1))
        return logits.detach()

    # ------------------------------------------------------------------ one iteration
    def _seg_d_gen(self, noise1, noise2, beats, real, fake_d, inner_d, real_index=None):
        """The D step's generator forwards (:294 -> :177-178)"""
        self._generators(noise1, noise2, beats, inner_d)

    def _seg_d_disc(self, noise1, noise2, beats, real, fake_d, inner_d, real_index=None):
        """D step up to the gradients (:293, :304-307)"""
        self._zero_d_grads()
        if self.tc is not None:
            self.tc.pack()               # one tiny launch: weights edited since the last optimiser step (load_state_dict, manual surgery) are picked up
        self.logit_fake_d = self._d_pass(fake_d, 0.0, self.loss_d, False)
        if self.tc is not None:
            self.logit_fake_d = self.logit_fake_d.clone()
        self.logit_real = self._d_pass(real, 1.0, self.loss_d, True, real_index)
        if self.tc is not None:
            self.logit_real = self.logit_real.clone()

    def _seg_d(self, *a):
        self._seg_d_gen(*a)
        self._seg_d_disc(*a)

    def _seg_g_gen(self, noise1, noise2, beats, fake_g, inner_g):
        """The G step's generator forwards (:312 -> :177-178).  They depend on nothing the D step produces (generator weights never change, SURVEY
        3.1), so when sharded they run while the D-gradient all-reduce is in flight."""
        self._generators(noise1, noise2, beats, inner_g)

    def _seg_d_opt(self):
        """Adam on D (:308) and the bf16 operand copies of the new weights"""
        self._disc_adam()
        if self.tc is not None:
            self.tc.pack()

    def _seg_g_disc(self, fake_g):
        """the rest of the G step (:311-315): gen_opt.zero_grad() leaves the D grads in place, gen_loss.backward() adds to them"""
        self.logit_fake_g = self._d_pass(fake_g, 1.0, self.loss_g, False)

    def _seg_g(self, noise1, noise2, beats, fake_g, inner_g):
        self._seg_d_opt()
        self._seg_g_disc(fake_g)

    def _capture(self, fn, *args):
        g = torch.cuda.CUDAGraph()
        l0 = N.lib().mmg_launch_count()
        # thread-local capture mode: other threads (NCCL's watchdog polling events, the bench's clock sampler) may call CUDA meanwhile
        with torch.cuda.graph(g, capture_error_mode="thread_local" if self.world > 1 else "global"):
            fn(*args)
        self.graph_launches += N.lib().mmg_launch_count() - l0
        return g

    # ------------------------------------------------------------------ public segments
    def _segment(self, name, fn, tensors, extra=()):
        """Run ``fn()`` eagerly the first time it is seen on these buffers, capture it in a CUDA graph the second time, replay afterwards."""
        if not (self.use_graph and self.on_d_grads is None):
            return fn()
        key = (name,) + tuple(t.data_ptr() if t is not None else 0 for t in tensors) + tuple(extra)
        g = self._seg_graphs.get(key)
        if g is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] == 1:
                return fn()
            torch.cuda.synchronize()
            l0 = N.lib().mmg_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local" if self.world > 1 else "global"):
                fn()
            self._seg_graphs[key] = (g, N.lib().mmg_launch_count() - l0)
            g = self._seg_graphs[key]
        self.replayed_launches += g[1]
        g[0].replay()

    def generators(self, noise1, noise2, beats, inner=None, out=None):
        """One pair of train-mode generator forwards (network_tests.py:294 / :312 -> :177-178; BatchNorm running statistics advance once).
        Returns ``(g1_out (B,1,S,S), g2_out (B,output_dim))`` fp32 -- what ``matrix_to_midi`` takes (matrix_sim_process.py:15,28-29).
        ``out=(buf1 (B,S*S), buf2 (B,output_dim))``: write into these buffers (the bf16 path otherwise reuses one internal pair per call);
        ``inner``: the Generator's in-forward randn (network_tests.py:83-84), drawn here when None."""
        B = noise1.shape[0]
        if inner is None:
            inner = self._draw_inner(B, "d")
        sync = self.sync_bn and not self.capture_collectives      # host-side collectives between the layer kernels: not capturable
        if sync or self.tc is None:
            self._generators(noise1, noise2, beats, inner, out)
        else:
            self._segment("gen", lambda: self._generators(noise1, noise2, beats, inner, out), (noise1, noise2, beats, inner) + tuple(out or ()), (B,))
            if out is not None:
                a = self.m.generator1.adj_size
                self.g1_out, self.g2_out = out[0].view(B, -1, a[0], a[1]), out[1]
        return self.g1_out, self.g2_out

    def d_step(self, real, fake_d, real_index=None):
        """The D step after the bridge has returned ``fake_d`` (network_tests.py:293, :304-308): zero grads, D on the fake rolls vs 0 and on the real
        rolls vs 1, backward, gradient all-reduce when sharded, Adam.  Returns the loss (device scalar)."""
        B = real.shape[0] if real_index is None else real_index.numel()
        a_d = (None, None, None, real, fake_d, None, real_index)
        self._sync_hyper()
        self._segment("d_bwd", lambda: self._seg_d_disc(*a_d), (real, fake_d, real_index), (B, real.dtype, fake_d.dtype))
        work = self._allreduce_d_grads(async_op=True)
        if work is not None:
            work.wait()
        if self.on_d_grads is not None:
            self.on_d_grads(self)
        self._segment("d_opt", self._seg_d_opt, ())
        return self.loss_d[0]

    def g_step(self, fake_g):
        """The G step after the bridge has returned ``fake_g`` (network_tests.py:311-315): D on the fake rolls vs 1, backward onto the D-step
        gradients (gen_opt.zero_grad() leaves them), generator Adam = no-op (no generator parameter has a grad).  Returns the loss."""
        self.gen_opt.zero_grad(set_to_none=True)
        self._segment("g", lambda: self._seg_g_disc(fake_g), (fake_g,), (fake_g.shape[0], fake_g.dtype))
        self.gen_opt.step()
        return self.loss_g[0]

    def step(self, noise1, noise2, beats, real, fake_d, fake_g, inner_d=None, inner_g=None, real_index=None):
        """``real_index`` (B,) int64 CUDA tensor: the real batch is ``real[real_index]`` -- ``real`` is then the whole HBM-resident training set and
        the discriminator kernel gathers the rows itself (no gathered copy of the batch)."""
        B = real.shape[0] if real_index is None else real_index.numel()
        if inner_d is None:
            inner_d = self._draw_inner(B, "d")
        if inner_g is None:
            inner_g = self._draw_inner(B, "g")
        self.gen_opt.zero_grad(set_to_none=True)
        a_d, a_g = (noise1, noise2, beats, real, fake_d, inner_d, real_index), (noise1, noise2, beats, fake_g, inner_g)
        graphs = None
        if self.use_graph and self.on_d_grads is None:
            key = tuple(t.data_ptr() for t in (*a_d, fake_g, inner_g) if t is not None) + (B, real.dtype, fake_d.dtype, fake_g.dtype, real_index is None)
            graphs = self._graphs.get(key)
            if graphs is None:
                self._seen[key] = self._seen.get(key, 0) + 1
                if self._seen[key] > 1:              # one eager iteration on these buffers first (allocator / lazy-init warm-up)
                    torch.cuda.synchronize()
                    self.graph_launches = 0
                    if self.world == 1:
                        graphs = (self._capture(lambda: (self._seg_d(*a_d), self._seg_g_gen(*a_g), self._seg_g(*a_g))),)
                    else:       # sharded: graphs around the asynchronous D-gradient all-reduce; SyncBN generators are captured with their NCCL all-reduces (eager under gloo)
                        eager_gen = self.sync_bn and not self.capture_collectives
                        graphs = (None if eager_gen else self._capture(self._seg_d_gen, *a_d), self._capture(self._seg_d_disc, *a_d),
                                  None if eager_gen else self._capture(self._seg_g_gen, *a_g), self._capture(self._seg_g, *a_g))
                    self._graphs[key] = graphs       # capture does not execute: fall through to the replay below
        self._sync_hyper()
        if graphs is None:
            self._seg_d(*a_d)
            work = self._allreduce_d_grads(async_op=True)       # 84 KB over NVLink, hidden behind the G step's generator forwards
            self._seg_g_gen(*a_g)
            if work is not None:
                work.wait()
            if self.on_d_grads is not None:
                self.on_d_grads(self)              # observer hook: flat_grad holds the (summed) D-step gradients Adam is about to consume
            self._seg_g(*a_g)
        else:
            self.replayed_launches += self.graph_launches
            if len(graphs) == 1:
                graphs[0].replay()
            else:
                gd_gen, gd, gg_gen, gg = graphs
                gd_gen.replay() if gd_gen is not None else self._seg_d_gen(*a_d)
                gd.replay()
                work = self._allreduce_d_grads(async_op=True)
                gg_gen.replay() if gg_gen is not None else self._seg_g_gen(*a_g)
                work.wait()
                gg.replay()
        self.gen_opt.step()          # no-op: generator grads are None
        return self.loss_d[0], self.loss_g[0]

class HostBatchPipeline:
    """Feeds HOST batches (pinned tensors, e.g. what the DataLoader and the host DES bridge hand over) to
    ``MMGANTrainer.step`` with the H2D copies of batch i+1 running on a copy stream underneath the compute of
    batch i (double-buffered device staging; SURVEY 8f-1: network_tests.py:192-193 does B small blocking copies).
    Every batch is copied exactly once; the two losses are read back (blocking, like the ``.item()`` calls at
    network_tests.py:320-321) after every step.

    ``dataset=(rolls, beats)``: the training set resident in HBM -- ``rolls`` (N,2,128,W) uint8/float32 and ``beats`` (N,50)
    CUDA tensors (the reference's ``MaestroDatasetPickle(..., device=device)`` also keeps its items on the device,
    datasets.py:73-87; 5.4 k MAESTRO slices are 69 MB).  A batch then carries ``real_idx`` (host int64 indices, what a
    sampler yields) instead of ``real`` / ``beats``; on the tensor-core path the discriminator kernel gathers the real rolls by index itself
    (``mmg_disc_fwd_fused_gather``), otherwise the gather runs on the copy stream.  The fake rolls always come from the host:

    * as rolls: ``fake_d`` / ``fake_g`` (B,2,128,W) uint8 / float32 (what ``matrix_to_midi`` returns, matrix_sim_process.py:191-195), or
    * as note events: ``fake_d_events`` / ``fake_g_events`` = ``(dt float64 (E,), meta int32 (E,), offsets int64 (B+1,))`` pinned tensors, the
      post-mido message streams of the B simulated songs (what ``process_adjsim_log`` hands to ``generate_piano_roll``,
      sim_log_to_midi.py:277).  They are copied H2D (12 bytes per message instead of 12.8 KB per roll) and rasterised on the device by
      ``mmg_raster_piano_roll`` straight into the uint8 roll buffers the discriminator kernels read -- bit-exact with the host rasterisation
      (SURVEY 8f-3).  ``raster=(sequence_length, start, end)`` are ``generate_piano_roll``'s arguments; ``max_events`` bounds E per pass."""

    KEYS = ("beats", "real", "fake_d", "fake_g")

    def __init__(self, trainer, example, dataset=None, max_events=None, raster=(100, 0, 50), g_out_host=False):
        self.t = trainer
        dev = trainer.flat_grad.device
        self.g_out_host = bool(g_out_host)
        self.dataset = dataset
        self.events = "fake_d_events" in example
        self.gather_in_kernel = dataset is not None and trainer.tc is not None and getattr(trainer.tc, "fused_forward", False)
        self.raster = tuple(int(v) for v in raster)
        if self.events:
            from .MMGAN_MIDI_DES import datasets as ds
            self._ds = ds
            B = example["fake_d_events"][2].numel() - 1
            W = ds.out_width(self.raster[1], self.raster[2])
            cap = int(max_events if max_events is not None else max(example[k][0].numel() for k in ("fake_d_events", "fake_g_events")))
            self.max_events = cap
            roll = lambda: torch.empty(B, 2, 128, W, dtype=torch.uint8, device=dev)
            ev = lambda: (torch.empty(cap, dtype=torch.float64, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
                          torch.empty(B + 1, dtype=torch.int64, device=dev))
            self.ws = [torch.empty(ds.raster_workspace_bytes(B, cap), dtype=torch.uint8, device=dev) for _ in range(2)]
        else:
            B = example["fake_d"].shape[0]
            roll = None
        fake = (lambda k: roll()) if self.events else (lambda k: torch.empty_like(example[k], device=dev))
        if dataset is not None:
            rolls, beats = dataset
            N.require_cuda(rolls, beats)
            self.copy_keys = () if self.events else ("fake_d", "fake_g")
            self.stage = [{"fake_d": fake("fake_d"), "fake_g": fake("fake_g"),
                           "real": None if self.gather_in_kernel else torch.empty((B,) + tuple(rolls.shape[1:]), dtype=rolls.dtype, device=dev),
                           "beats": torch.empty((B,) + tuple(beats.shape[1:]), dtype=beats.dtype, device=dev),
                           "idx": torch.empty(B, dtype=torch.int64, device=dev)} for _ in range(2)]
            self.h2d_bytes = sum(example[k].numel() * example[k].element_size() for k in self.copy_keys) + B * 8
        else:
            self.copy_keys = ("beats", "real") if self.events else self.KEYS
            self.stage = [{k: (fake(k) if k.startswith("fake") else torch.empty_like(example[k], device=dev)) for k in self.KEYS} for _ in range(2)]
            self.h2d_bytes = sum(example[k].numel() * example[k].element_size() for k in self.copy_keys)
        if self.events:
            for st in self.stage:
                st["ev_d"], st["ev_g"] = ev(), ev()
            self.h2d_bytes += sum(t.numel() * t.element_size() for k in ("fake_d_events", "fake_g_events") for t in example[k])
        self.noise = [[torch.empty(B, trainer.m.z_dim, device=dev) for _ in range(2)] for _ in range(2)]     # static: graph replay
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.losses_host = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.read = [torch.cuda.Event() for _ in range(2)]
        self.d2h_bytes = 8
        self.g_out = None
        if self.g_out_host:
            # the host DES consumes the generator outputs of BOTH forwards of an iteration (matrix_sim_process.py:28-29 does .cpu().numpy() on
            # (B,1,S,S) and (B,output_dim) fp32): device staging + pinned host buffers per pipeline slot, copies on their own stream
            g1, g2 = trainer.m.generator1, trainer.m.generator2
            n1, n2 = g1.gen[-1][0].out_features, g2.gen[-1][0].out_features
            self.g_dev = [[(torch.empty(B, n1, device=dev), torch.empty(B, n2, device=dev)) for _ in range(2)] for _ in range(2)]
            self.g_host = [[(torch.empty(B, n1).pin_memory(), torch.empty(B, n2).pin_memory()) for _ in range(2)] for _ in range(2)]
            self.d2h_stream = torch.cuda.Stream(device=dev)
            self.g_made = [[torch.cuda.Event() for _ in range(2)] for _ in range(2)]
            self.g_done = [torch.cuda.Event() for _ in range(2)]
            self.d2h_bytes += 2 * B * (n1 + n2) * 4

    def _raster(self, slot, dev_ev, host_ev, out):
        """H2D of one pass's message streams + device rasterisation into the pass's uint8 roll buffer (on the copy stream)."""
        dt, meta, off = host_ev
        E = dt.numel()
        if E > self.max_events:
            raise ValueError(f"{E} messages exceed the pipeline's max_events = {self.max_events}")
        if off.numel() != dev_ev[2].numel():
            raise ValueError("offsets must have batch + 1 entries")
        d_dt, d_meta = dev_ev[0][:E], dev_ev[1][:E]
        d_dt.copy_(dt, non_blocking=True)
        d_meta.copy_(meta, non_blocking=True)
        dev_ev[2].copy_(off, non_blocking=True)
        S, a, b = self.raster
        self._ds.rasterize_events(d_dt, d_meta, dev_ev[2], S, a, b, out=out, workspace=self.ws[slot])

    def _issue(self, slot, batch, first_use):
        with torch.cuda.stream(self.copy_stream):
            if not first_use:
                self.copy_stream.wait_event(self.free[slot])
            st = self.stage[slot]
            for k in self.copy_keys:
                st[k].copy_(batch[k], non_blocking=True)
            if self.events:
                self._raster(slot, st["ev_d"], batch["fake_d_events"], st["fake_d"])
                self._raster(slot, st["ev_g"], batch["fake_g_events"], st["fake_g"])
            if self.dataset is not None:
                st["idx"].copy_(batch["real_idx"], non_blocking=True)
                if not self.gather_in_kernel:
                    torch.index_select(self.dataset[0], 0, st["idx"], out=st["real"])
                torch.index_select(self.dataset[1], 0, st["idx"], out=st["beats"])
            self.ready[slot].record(self.copy_stream)

    def _iteration(self, slot, i, n1, n2, st):
        """One iteration on staged inputs.  With ``g_out_host`` the iteration runs as the public segments and the outputs of both generator
        forwards go to pinned host memory on the D2H stream, underneath the discriminator passes."""
        t = self.t
        gather = self.dataset is not None and self.gather_in_kernel
        real, idx = (self.dataset[0], st["idx"]) if gather else (st["real"], None)
        if not self.g_out_host:
            return t.step(n1, n2, st["beats"], real, st["fake_d"], st["fake_g"], real_index=idx)
        main = torch.cuda.current_stream()
        if i >= 2:
            main.wait_event(self.g_done[slot])                         # the copies of two iterations ago have left the device staging buffers
        for which in range(2):                                         # D-step forward (:294), G-step forward (:312): neither depends on the D step
            t.generators(n1, n2, st["beats"], out=self.g_dev[slot][which])
            self.g_made[slot][which].record(main)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(self.g_made[slot][which])
                for h, d in zip(self.g_host[slot][which], self.g_dev[slot][which]):
                    h.copy_(d, non_blocking=True)
        self.g_done[slot].record(self.d2h_stream)
        dl = t.d_step(real, st["fake_d"], real_index=idx)
        gl = t.g_step(st["fake_g"])
        return dl, gl

    def run(self, batches):
        """Generator: yields a (2,) CPU tensor [disc_loss, gen_loss] for every batch, in order (a copy: it stays valid).  The read-back of
        iteration i is enqueued right behind it (stream order) and waited for AFTER iteration i + 1 has been enqueued, so the host-side work of
        the next iteration (copies, rasteriser launches, graph launch) does not leave the GPU idle behind a blocking ``.item()``
        (network_tests.py:320-321 blocks; the values are the same, they arrive one iteration later; the last one is flushed at the end).
        With ``g_out_host``, ``self.g_out`` = ((g1, g2) of the D-step forward, (g1, g2) of the G-step forward) of the batch being yielded: pinned
        host tensors that are overwritten two batches later.  Buffer lifetime of the caller's pinned input batches: batch i may be reused once
        batch i + 1 has been yielded (its H2D copies were enqueued before iteration i and have completed by then)."""
        main = torch.cuda.current_stream()
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        self._issue(0, cur, True)
        i = 0
        pending = None
        while cur is not None:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                self._issue(1 - slot, nxt, i == 0)
            main.wait_event(self.ready[slot])
            st = self.stage[slot]
            n1, n2 = self.noise[slot]
            n1.normal_()                                                          # network_tests.py:284-285
            n2.normal_()
            dl, gl = self._iteration(slot, i, n1, n2, st)        # (resident set: the discriminator kernel reads the real rolls by index itself)
            self.free[slot].record(main)
            host = self.losses_host[slot]
            host[0:1].copy_(dl.reshape(1), non_blocking=True)
            host[1:2].copy_(gl.reshape(1), non_blocking=True)
            self.read[slot].record(main)
            if pending is not None:
                yield self._deliver(pending)
            pending = slot
            cur, i = nxt, i + 1
        yield self._deliver(pending)

    def _deliver(self, slot):
        self.read[slot].synchronize()
        if self.g_out_host:
            self.g_done[slot].synchronize()
            self.g_out = tuple(self.g_host[slot])
        return self.losses_host[slot].clone()

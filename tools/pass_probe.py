"""GPU probe of the one-kernel discriminator pass (csrc/disc_tc_pass.cu): parity against the two-kernel path (forward kernel + BCE kernel +
backward kernel) on the same inputs, with a watchdog that prints the kernel's progress words if it hangs, then timing at the bench batch.
usage: python tools/pass_probe.py [parity|time|all]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch

import mmgan_oracle as mo
from gan_des_midi_music_gen_b200 import _native as N
from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
from gan_des_midi_music_gen_b200.disc_tc import DiscTC

DEV = "cuda"
dbg = torch.zeros(2048, dtype=torch.int32).pin_memory()
state = {"what": "start"}


def watchdog(limit):
    t0 = time.time()
    while time.time() - t0 < limit:
        time.sleep(0.5)
        if state["what"] == "done":
            return
    d = dbg.numpy()[:592].reshape(148, 4)
    print("WATCHDOG: hung in", state["what"], flush=True)
    print("progress words [cta: producer, mma, worker0] (it*16 + stage):", flush=True)
    for c in range(0, 148, 37):
        print(c, d[c].tolist(), flush=True)
    print("distinct:", sorted({tuple(r) for r in d.tolist()})[:12], flush=True)
    os._exit(3)


def make_disc(seed=31):
    sd = mo.synth_state(mo.mmgan_shapes(), seed=seed, d_scale=0.25)
    D = nt.DiscriminatorCNN(roll_size=(2, 128, 50)).to(DEV)
    D.load_state_dict({k[len("discriminator."):]: v for k, v in sd.items() if k.startswith("discriminator.")})
    return D


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def parity():
    D = make_disc()
    names = [n for n, _ in D.named_parameters()]
    for B, dtype, target in ((5, torch.uint8, 0.0), (3, torch.float32, 1.0), (37, torch.uint8, 1.0), (300, torch.uint8, 0.0), (1000, torch.uint8, 1.0)):
        tc = DiscTC(D, max_batch=B + 2)
        x8 = torch.from_numpy(mo.synth_rolls(B, 50, seed=32 + B, p=0.05)).to(DEV)
        x = x8 if dtype == torch.uint8 else x8.float()
        # two-kernel path
        for p in D.parameters():
            p.grad = None
        logits = tc.forward(x).clone()
        dl = torch.empty(B, device=DEV)
        loss = torch.zeros(1, device=DEV)
        N.call("mmg_bce_logits_f32", N.ptr(logits), None, float(target), B, N.ptr(loss), 0, N.ptr(dl), 1.0 / B, None, N.stream())
        tc.backward(dl)
        torch.cuda.synchronize()
        want = {n: p.grad.clone() for n, p in D.named_parameters()}
        # one-kernel pass
        for p in D.parameters():
            p.grad = None
        loss2 = torch.zeros(1, device=DEV)
        state["what"] = f"pass_fused B={B}"
        dbg.zero_()
        l2 = tc.pass_fused(x, target, loss2, dbg=dbg.data_ptr())
        torch.cuda.synchronize()
        state["what"] = "between"
        got = {n: p.grad.clone() for n, p in D.named_parameters()}
        dlog = (l2 - logits).abs().max().item()
        print(f"B={B} {dtype} y={target}: |dlogit|max {dlog:.3e} (scale {logits.abs().max().item():.3e}) loss {loss.item():.6f} vs {loss2.item():.6f}", flush=True)
        errs = {n: rel(got[n], want[n]) for n in names}
        print("   grad rel-L2 vs two-kernel path:", {k: f"{v:.2e}" for k, v in errs.items()}, flush=True)
        ok = dlog <= 2e-4 * logits.abs().max().item() + 1e-6 and abs(loss.item() - loss2.item()) <= 1e-5 * abs(loss.item()) + 1e-7 and all(v < 2e-3 for v in errs.values())
        print("   OK" if ok else "   MISMATCH", flush=True)
        # second call accumulates
        tc.pass_fused(x, target, loss2)
        torch.cuda.synchronize()
        acc = {n: rel(p.grad, 2 * want[n]) for n, p in D.named_parameters()}
        print("   accumulate:", {k: f"{v:.1e}" for k, v in acc.items()}, "loss2", loss2.item(), flush=True)


def timeline():
    """SM-clock timeline of CTA 0 for samples 6 and 7 of a B = 16384 pass (debug entry point)."""
    D = make_disc()
    B = 16384
    tc = DiscTC(D, max_batch=B)
    g = torch.Generator(device=DEV).manual_seed(5)
    x = ((torch.rand(B, 2, 128, 50, device=DEV, generator=g) < 0.02) * torch.randint(1, 128, (B, 2, 128, 50), device=DEV, generator=g)).to(torch.uint8)
    loss = torch.zeros(1, device=DEV)
    for _ in range(2):
        tc.pass_fused(x, 1.0, loss)
    torch.cuda.synchronize()
    dbg.zero_()
    state["what"] = "timeline"
    tc.pass_fused(x, 1.0, loss, dbg=dbg.data_ptr())
    torch.cuda.synchronize()
    print("progress words of CTA 0 / 147:", dbg.numpy()[:4].tolist(), dbg.numpy()[588:592].tolist(), "raw stamps:", dbg.numpy()[1024:1032].tolist(), dbg.numpy()[1040:1048].tolist())
    t = dbg.numpy()[1024:1024 + 64].astype(np.int64) & 0xFFFFFFFF
    names_w = ["loop top", "wg1a_done(prev) + c1a_done seen", "S3 done", "XSb(next) done", "S5 done", "W1 done", "W3 done"]
    names_m = ["loop top", "C2 issued (all tiles)", "conv1(next) tiles 0..7 + conv2 wgrad + dgrad issued (all tiles)", "conv1 wgrad parts 0..2 issued", "conv1 wgrad part 3 issued", "conv1(next) tiles 8..13 issued"]
    t0 = int(t[16])
    for smp in range(2):
        print(f"-- sample {6 + smp} of CTA 0 (cycles relative to the MMA thread's loop top of sample 6)")
        for k, n in enumerate(names_w):
            print(f"   W {n:28s} {(int(t[smp * 32 + k]) - t0) & 0xFFFFFFFF:8d}")
        for k, n in enumerate(names_m):
            print(f"   M {n:28s} {(int(t[smp * 32 + 16 + k]) - t0) & 0xFFFFFFFF:8d}")
    state["what"] = "between"


def timing():
    D = make_disc()
    B = 16384
    tc = DiscTC(D, max_batch=B)
    g = torch.Generator(device=DEV).manual_seed(5)
    x = ((torch.rand(B, 2, 128, 50, device=DEV, generator=g) < 0.02) * torch.randint(1, 128, (B, 2, 128, 50), device=DEV, generator=g)).to(torch.uint8)
    loss = torch.zeros(1, device=DEV)
    dl = torch.empty(B, device=DEV)
    for p in D.parameters():
        p.grad = None

    def two():
        logits = tc.forward(x)
        N.call("mmg_bce_logits_f32", N.ptr(logits), None, 1.0, B, N.ptr(loss), 0, N.ptr(dl), 1.0 / B, None, N.stream())
        tc.backward(dl)

    def one():
        tc.pass_fused(x, 1.0, loss)

    for name, fn in (("two-kernel", two), ("one-kernel", one)):
        state["what"] = "timing " + name
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"{name}: {us:.1f} us per pass at B={B}  ({2 * 11112448 * B / us / 1e6:.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    threading.Thread(target=watchdog, args=(120,), daemon=True).start()
    for flags in (0, 1):
        N.lib().mmg_disc_pass_set_flags(flags)
        print(f"==== pass flags {flags} (bit 0: biases in the epilogues instead of the bias MMAs)", flush=True)
        if mode in ("parity", "all"):
            parity()
        if mode in ("time", "all"):
            timing()
        if mode in ("timeline", "all"):
            timeline()
    N.lib().mmg_disc_pass_set_flags(0)
    state["what"] = "done"

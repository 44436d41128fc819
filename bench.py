"""bench.py -- MM-GAN training throughput (piano-rolls/sec) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision bf16|fp32] [--impl ours|reference]

A "step" is one iteration of the reference loop body (MMGAN_MIDI_DES/network_tests.py:292-315) on a
synthetic batch of B piano rolls per GPU: D step (G1+G2 train-mode forward, D forward on the fake and
the real rolls, BCE, backward, Adam) then G step (G1+G2 forward, D forward on the fake rolls, BCE vs
ones, backward, no-op generator update).  The host DES is excluded on both arms: the
fake rolls it would return are synthetic inputs.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FLOP_PER_ROLL = 68.26e6          # 34 130 432 MAC per roll per iteration
METRIC = "mmgan_piano_rolls_per_sec_trained"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the loop body (oracle port of the
    unmodified PyTorch code -- the reference itself is Python and cannot travel to the GPU box), fp32,
    all host threads, on a bounded sample of the workload (same per-roll work, smaller batch)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import mmgan_oracle as mo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = args.ref_batch
    sd = mo.synth_state(mo.mmgan_shapes(), seed=0, d_scale=0.25)
    adam = {}
    inp = mo.synth_inputs(B, seed=1)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        mo.mmgan_iteration(sd, adam, inp, lr=0.01)
        times.append(time.perf_counter() - t0)
    dt = sum(times[args.warmup:]) / args.steps
    val = B / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "rolls/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MM-GAN G+D training iteration (network_tests.py:292-315), DES excluded", "batch": B,
                       "roll_size": [2, 128, 50], "adj_size": [64, 64], "z_dim": 50},
            "cpu_baseline": {"value": val, "unit": "rolls/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} iterations of batch {B} (bounded sample of the per-GPU batch)"},
            "e2e": {"value": val, "unit": "rolls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_eager(args):
    """Library-GPU bar: the reference's modules and loop body under PyTorch EAGER on the same B200 (cuBLAS / cuDNN / ATen),
    fp32 as the reference runs them and with bf16 autocast; same batch, same synthetic rolls (float32 device tensors, DES excluded), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    import eager_baseline as eb
    from gan_des_midi_music_gen_b200 import benchmark
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B = args.batch or 16384
    rolls = {k: benchmark._synth_rolls_u8(B, 50, i, "cpu").to(dev).float() for i, k in enumerate(("real", "fake_d", "fake_g"))}
    noise1, noise2, beats = torch.randn(B, 50, device=dev), torch.randn(B, 50, device=dev), 25 * torch.rand(B, 50, device=dev)
    out = {}
    for name, kw in (("fp32", {}), ("fp32_tf32", {"tf32": True}), ("bf16_autocast", {"autocast": True}), ("bf16_autocast_channels_last", {"autocast": True, "channels_last": True})):
        tf32 = kw.pop("tf32", False)
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        m = eb.EagerMMGAN(dev, **kw)
        x = {k: (v.contiguous(memory_format=torch.channels_last) if kw.get("channels_last") else v) for k, v in rolls.items()}
        fn = lambda: m.iteration(noise1, noise2, beats, x["real"], x["fake_d"], x["fake_g"])
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / args.steps
        out[name] = {"rolls_per_sec": B / sec, "ms_per_step": sec * 1e3}
        del m
        torch.cuda.empty_cache()
    best = max(out, key=lambda k: out[k]["rolls_per_sec"])
    print(json.dumps({"impl": "eager", "metric": METRIC, "value": out[best]["rolls_per_sec"], "unit": "rolls/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": out[best]["ms_per_step"], "higher_is_better": True, "dtype": best, "data": "synthetic",
                      "config": {"workload": "MM-GAN G+D training iteration (network_tests.py:292-315) under PyTorch eager CUDA (stock nn modules, torch.optim.Adam), DES excluded",
                                 "batch": B, "inputs": "float32 rolls resident on the device"}, "variants": out}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="piano rolls per GPU per step")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--ref-batch", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-raster", action="store_true")
    ap.add_argument("--sync-bn", action="store_true", help="N > 1: generator BatchNorm over the global batch (statistics all-reduced between the layer kernels)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "eager":
        return run_eager(args)
    from gan_des_midi_music_gen_b200 import benchmark
    benchmark.run(args)


if __name__ == "__main__":
    main()

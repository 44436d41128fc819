"""Measurement harness behind bench.py (our arm).  See bench.py for the contract."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from .MMGAN_MIDI_DES import datasets as ds
from .MMGAN_MIDI_DES import network_tests as nt
from .trainer import HostBatchPipeline, MMGANTrainer

FLOP_PER_ROLL = 68.26e6
METRIC = "mmgan_piano_rolls_per_sec_trained"


def _peaks():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(root, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def _synth_rolls_u8(B, W, seed, device):
    """(B,2,128,W) uint8 rolls: each cell non-zero with p=0.02, generated on host in
    chunks so that large batches do not need a float64 staging array."""
    rng = np.random.default_rng(seed)
    out = torch.empty(B, 2, 128, W, dtype=torch.uint8)
    for b0 in range(0, B, 4096):
        n = min(4096, B - b0)
        mask = rng.random((n, 2, 128, W), dtype=np.float32) < 0.02
        vel = rng.integers(1, 128, size=(n, 128, W), dtype=np.uint8)
        dur = rng.integers(1, W + 1, size=(n, 128, W), dtype=np.uint8)
        out[b0:b0 + n] = torch.from_numpy(np.stack([vel, dur], axis=1) * mask)
    return out


def _synth_fake_events(B, seed, E=320, T=60.0):
    """Post-mido message streams of B simulated songs, the form in which the host DES bridge hands its output to ``generate_piano_roll``
    (sim_log_to_midi.py:277): E messages per song, dt ~ Exp(mean T/E) float64 seconds, 40 % note_on / 40 % note_off / 20 % other, pitch
    21..108, velocity 0..127.  With the reference's call shape (sequence_length 100, window 50) about 2 % of the roll plane is non-zero,
    like the synthetic uint8 rolls.  Returns pinned (dt, meta as int32, offsets) host tensors."""
    rng = np.random.default_rng(seed)
    n = B * E
    dt = rng.exponential(T / E, size=n)
    u = rng.random(n, dtype=np.float32)
    kind = np.where(u < 0.4, 1, np.where(u < 0.8, 2, 0)).astype(np.uint32)
    meta = kind | (rng.integers(21, 109, size=n, dtype=np.uint32) << 8) | (rng.integers(0, 128, size=n, dtype=np.uint32) << 16)
    meta[kind == 0] = 0
    off = np.arange(B + 1, dtype=np.int64) * E
    return tuple(torch.from_numpy(a).pin_memory() for a in (dt, meta.view(np.int32), off))


def _timed(fn, steps, sync):
    """CUDA-event timing of `steps` calls of fn on the current stream; returns seconds."""
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    sync()
    return e0.elapsed_time(e1) * 1e-3


def _raster_leg(device, peaks, quick):
    """Secondary metric (BASELINE config 4): notes/s of the rasteriser on MAESTRO-scale synthetic streams."""
    import raster_oracle as ro        # bench.py's oracle use is limited to the cpu_baseline sample below
    S, E, T = (1276, 15000, 300.0) if not quick else (128, 15000, 300.0)
    dt, meta, off = ro.synth_songs(S, E, T, seed=0)
    kinds = meta & 0xFF
    d_dt, d_meta, d_off = (torch.from_numpy(a).to(device) for a in (dt, meta.view(np.int32), off))
    fn = lambda i: ds.rasterize_events(d_dt, d_meta, d_off, 300, 0, 300)
    out = fn(0)
    for _ in range(3):
        fn(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    times = []
    for _ in range(5):
        flush.zero_()                                   # L2 flush between timed iterations
        times.append(_timed(fn, 1, torch.cuda.synchronize))
    sec = min(times)
    # notes actually applied = note_on messages before each song's cut-off (datasets.py:33-35), counted by the C oracle: a timed single-thread
    # sample for the cpu_baseline, the remaining songs on all host threads for the count only
    sample = min(S, 64)
    t0 = time.perf_counter()
    _, notes_sample = ro.raster_batch_c(dt[:off[sample]], meta[:off[sample]], off[:sample + 1], 300, 0, 300)
    cpu_sec = time.perf_counter() - t0
    n_on = int((kinds == 1).sum())
    n_applied = int(notes_sample)
    if S > sample:
        rest = off[sample:] - off[sample]
        n_applied += int(ro.raster_batch_c(dt[off[sample]:], meta[off[sample]:], rest, 300, 0, 300, n_threads=os.cpu_count() or 1)[1])
    alg_bytes = 12.0 * S * E + 2 * 128 * 300 * 4.0 * S          # what the kernels read: f64 dt + u32 meta per message, + the f32 roll
    alg_bytes_8 = 8.0 * S * E + 2 * 128 * 300 * 4.0 * S         # SURVEY 8d's packed 8-byte record
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    if os.path.exists(tpath):                                # both kernels of the sort path, one call, from the committed ncu --set full capture
        t = json.load(open(tpath)).get("raster_sort_path")
        if t and t.get("songs") == S and t.get("messages_per_song") == E:
            traffic = t["dram_bytes"]
    return {"notes_per_sec": n_applied / sec, "note_on_messages_per_sec": n_on / sec, "messages_per_sec": S * E / sec, "ms": sec * 1e3, "songs": S,
            "messages_per_song": E, "window": 300, "notes_applied": n_applied, "note_on_messages": n_on,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": alg_bytes / sec / 1e9 / peaks["hbm_gbs"], "traffic": traffic, "peak_src": peaks["src"],
                         "algorithmic_bytes": alg_bytes, "algorithmic_bytes_8B_records": alg_bytes_8,
                         "frac_8B_records": alg_bytes_8 / sec / 1e9 / peaks["hbm_gbs"]},
            "cpu_baseline": {"value": notes_sample / cpu_sec, "unit": "notes/s", "cores": 1, "kind": "port",
                             "sample": f"first {sample} songs through oracle/raster_oracle.c"},
            "checksum": float(out.double().sum().item())}


def _gandes_leg(device):
    """BASELINE config 2: GAN-DES (GAN_DES/SIMNN.py:275-334) G+D training iteration on one synthetic song matrix (B = 30 spectrograms of
    128 x 216); headline = every contraction on the tcgen05 GEMM kernel (bf16 operands, fp32 accumulation), the fp32 SIMT kernels of this
    library beside it; fused BCE / multi-tensor Adam; DES + FluidSynth bridge excluded (the fake spectrograms are inputs)."""
    import mmgan_oracle as mo               # cpu_baseline sample only
    from .GAN_DES import SIMNN
    from . import optim as fo
    B = 30
    gshapes, dshapes = mo.gandes_shapes()
    gsd, dsd = mo.synth_state(gshapes, seed=11), mo.synth_state(dshapes, seed=12)
    g = torch.Generator().manual_seed(3)
    noise_h, real_h, fake_h = torch.randn(B, 100, 1, 1, generator=g), torch.randn(B, 128, 216, generator=g), torch.randn(B, 128, 216, generator=g)
    noise, real, fake = noise_h.to(device), real_h.to(device), fake_h.to(device)
    t09, t01, t1 = torch.full((B,), 0.9, device=device), torch.full((B,), 0.1, device=device), torch.ones(B, device=device)
    res = {}
    for name, tcores in (("fp32_simt", False), ("tensor_cores", True)):
        gen, disc = SIMNN.Generator().to(device).enable_tensor_cores(tcores), SIMNN.Discriminator().to(device).enable_tensor_cores(tcores)
        gen.load_state_dict(gsd); disc.load_state_dict(dsd)
        crit = fo.BCEWithLogitsLoss()
        gen_opt = fo.FusedAdam(gen.parameters(), lr=2e-5, betas=(0.5, 0.999))
        disc_opt = fo.FusedAdam(disc.parameters(), lr=2e-5, betas=(0.5, 0.999))

        def it(i):
            disc_opt.zero_grad()
            l_real = crit(disc(real).reshape(-1), t09)
            with torch.no_grad():
                gen(noise)
            l_fake = crit(disc(fake.detach()).reshape(-1), t01)
            (l_fake + l_real).backward()
            disc_opt.step()
            gen_opt.zero_grad()
            crit(disc(fake).squeeze(), t1).backward()
            gen_opt.step()

        for _ in range(3):
            it(0)
        l0 = N.lib().mmg_launch_count()
        sec = _timed(it, 10, torch.cuda.synchronize) / 10
        res[name] = (sec, (N.lib().mmg_launch_count() - l0) // 10)
        del gen, disc, gen_opt, disc_opt
    # the same iteration through GANDESTrainer: generate / d_step / g_step replayed from CUDA graphs (tensor-core path)
    from .gandes_trainer import GANDESTrainer
    gen, disc = SIMNN.Generator().to(device).enable_tensor_cores(), SIMNN.Discriminator().to(device).enable_tensor_cores()
    gen.load_state_dict(gsd); disc.load_state_dict(dsd)
    gtr = GANDESTrainer(gen, disc, lr=2e-5, betas=(0.5, 0.999))

    def it_graph(i):
        gtr.generate(noise)
        gtr.d_step(real, fake)
        gtr.g_step(fake)

    for _ in range(4):
        it_graph(0)
    r0 = gtr.replayed_launches
    sec_graph = _timed(it_graph, 10, torch.cuda.synchronize) / 10
    res["tensor_cores_graph"] = (sec_graph, (gtr.replayed_launches - r0) // 10)
    adam = {}
    mo.gandes_iteration(gsd, dsd, adam, noise_h, real_h, fake_h)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 5.0 and n < 20:
        mo.gandes_iteration(gsd, dsd, adam, noise_h, real_h, fake_h)
        n += 1
    cs = (time.perf_counter() - t0) / n
    sec, launches = res["tensor_cores_graph"]
    # algorithmic HBM traffic of one iteration, fp32 tensors (the reference's): 3 D passes read a (B,128,216) batch and touch the conv1 / conv2
    # activations (written, read by the pool, re-read by the backward) + fc1's 28.3 MB weight (3 forward + 3 dgrad reads, 3 wgrad writes) + Adam
    # on 7.1 M parameters (28 B each)
    act = B * 4 * (128 * 216 + 16 * 129 * 217 + 16 * 64 * 108 + 32 * 64 * 108 + 32 * 32 * 54)
    alg_bytes = 3 * (3 * act) + 9 * 128 * 55296 * 4 + 7.1e6 * 28
    peaks = _peaks()
    return {"spectrograms_per_sec": B / sec, "ms_per_step": sec * 1e3, "batch": B, "dtype": "bf16 operands, fp32 accumulation (tcgen05)",
            "launches_per_step": launches, "api": "gandes_trainer.GANDESTrainer: generate / d_step / g_step, each replayed from its CUDA graph; the D step runs real and fake as one batch of 2B (no BatchNorm in D: same logits and gradients)",
            "module_loop": {"ms_per_step": res["tensor_cores"][0] * 1e3, "spectrograms_per_sec": B / res["tensor_cores"][0], "launches_per_step": res["tensor_cores"][1],
                            "api": "the reference's loop written with the drop-in modules (enable_tensor_cores), FusedAdam and the fused BCE, eager launches"},
            "fp32_simt": {"ms_per_step": res["fp32_simt"][0] * 1e3, "spectrograms_per_sec": B / res["fp32_simt"][0], "launches_per_step": res["fp32_simt"][1]},
            "roofline": {"bound": "hbm", "achieved": alg_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / sec / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes": alg_bytes, "traffic": None, "peak_src": peaks["src"],
                         "note": "whole iteration (about 90 small dependent launches at B = 30: launch latency, not bandwidth, bounds it)"},
            "cpu_baseline": {"value": B / cs, "unit": "spectrograms/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": f"{n} iterations of batch {B} through oracle/mmgan_oracle.gandes_iteration"}}


def _mel_leg(device, peaks):
    """the GAN-DES mel front end (GAN_DES/util.py:37-61) for one batch of real windows: 30 five-second windows at 44.1 kHz
    (datasets.py:85-90, SIMNN.py:236) -> (30, 128, 216) dB spectrograms; waveforms resident on the device."""
    import mel_oracle as mel                # cpu_baseline sample only
    from .GAN_DES import util
    B, L = 30, 220500
    waves_h = np.stack([mel.synth_wave(L, 100 + b) for b in range(B)])
    waves = torch.from_numpy(waves_h).to(device)
    fn = lambda i: util.get_melspectrogram_db_tensor(waves)
    out = fn(0)
    for _ in range(3):
        fn(0)
    l0 = N.lib().mmg_launch_count()
    sec = _timed(fn, 20, torch.cuda.synchronize) / 20
    launches = (N.lib().mmg_launch_count() - l0) // 20
    t0 = time.perf_counter()
    want = mel.get_melspectrogram_db_tensor(waves_h[0])
    cs = time.perf_counter() - t0
    err = float(np.abs(out[0].cpu().numpy() - want).max())
    alg_bytes = B * L * 4.0 + B * 128 * 216 * 4.0
    return {"windows_per_sec": B / sec, "us": sec * 1e6, "batch": B, "samples_per_window": L, "launches": launches, "max_abs_db_error_vs_oracle": err,
            "dtype": "f32 FFT, tf32 mel projection",
            "roofline": {"bound": "hbm", "achieved": alg_bytes / sec / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg_bytes / sec / 1e9 / peaks["hbm_gbs"],
                         "algorithmic_bytes": alg_bytes, "traffic": None, "peak_src": peaks["src"]},
            "cpu_baseline": {"value": 1.0 / cs, "unit": "windows/s", "cores": 1, "kind": "port", "sample": "1 window through oracle/mel_oracle.py (numpy float64)"}}


def _inference_leg(mmgan, device):
    """BASELINE config 5: eval-mode G1 + G2 (running-stat BatchNorm folded into the tcgen05 prologue / epilogue) for B = 1 .. 16384, outputs
    (B,1,64,64) + (B,20) fp32 delivered to pinned host memory -- what matrix_to_midi consumes (matrix_sim_process.py:28-29)."""
    from .gen_tc import GenTC
    out = {}
    for B in (1, 16, 256, 4096, 16384):
        g1, g2 = GenTC(mmgan.generator1, B), GenTC(mmgan.generator2, B)
        n = [torch.randn(B, 50, device=device) for _ in range(4)]
        o1, o2 = torch.empty(B, 4096, device=device), torch.empty(B, 20, device=device)
        h1, h2 = torch.empty(B, 4096).pin_memory(), torch.empty(B, 20).pin_memory()

        def it(i):
            g1.forward(n[0], n[1], training=False, out=o1)
            g2.forward(n[2], n[3], training=False, out=o2)
            h1.copy_(o1, non_blocking=True)
            h2.copy_(o2, non_blocking=True)

        for _ in range(3):
            it(0)
        sec = _timed(it, 10, torch.cuda.synchronize) / 10
        out[str(B)] = {"samples_per_sec": B / sec, "us": sec * 1e6}
    return out


def _host_front_end_leg():
    """The host code on the input side of the rasteriser (no GPU involved): the native sim-log -> note-event conversion of a batch of simulated songs
    (csrc/simlog.cu, all host threads) beside the Python state machine it mirrors, and the native Standard MIDI File reader (csrc/smf.cu) beside its
    Python checker.  Synthetic inputs; a few seconds of CPU time."""
    import struct
    import smf_oracle as so                      # cpu baseline sample only
    from .MMGAN_MIDI_DES import datasets as ds, sim_log_to_midi as sl
    rng = np.random.default_rng(0)
    S, n_lines = 2048, 320
    one = []
    for _ in range(32):
        t = np.sort(rng.random(n_lines) * 60.0)
        one.append("".join(f"INFO:root:{t[i]:.4f} - {int(rng.integers(0, 40))} - {int(rng.integers(0, 16))} - {'arrival' if rng.random() < 0.55 else 'departure'}\n"
                           for i in range(n_lines)).encode())
    logs = [one[i % 32] for i in range(S)]
    ins, nl, g = rng.integers(0, 100, (S, 16)), rng.integers(0, 128, (S, 16)), rng.random((S, 10)).astype(np.float32)
    sl.sim_logs_to_event_batch(logs[:64], ins[:64], nl[:64], g[:64], True)
    t0 = time.perf_counter()
    dt, _, _ = sl.sim_logs_to_event_batch(logs, ins, nl, g, True)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    n_py = 16
    for i in range(n_py):
        sl.sim_log_to_event_stream(logs[i].decode().splitlines(True), ins[i], nl[i], g[i], True)
    t_py = (time.perf_counter() - t0) / n_py
    # one MAESTRO-scale file: 3 tracks x 12 000 note events with running status and tempo changes
    def vlq(n):
        out = [n & 0x7F]
        n >>= 7
        while n:
            out.append(0x80 | (n & 0x7F))
            n >>= 7
        return bytes(reversed(out))
    tracks = []
    for _ in range(3):
        body = bytearray()
        for k in range(12000):
            body += vlq(int(rng.integers(0, 240)))
            body += (b"\xff\x51\x03" + int(rng.integers(300000, 900000)).to_bytes(3, "big")) if k % 500 == 0 else bytes([0x90 if k % 2 == 0 else 0x80, int(rng.integers(21, 109)), int(rng.integers(1, 128))])
        body += b"\x00\xff\x2f\x00"
        tracks.append(b"MTrk" + struct.pack(">I", len(body)) + bytes(body))
    raw = b"MThd" + struct.pack(">IHHH", 6, 1, 3, 480) + b"".join(tracks)
    ev = ds.parse_smf_bytes(raw)
    t0 = time.perf_counter()
    for _ in range(5):
        ds.parse_smf_bytes(raw)
    t_smf = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    so.read_smf_bytes(raw)
    t_smf_py = time.perf_counter() - t0
    return {"simlog_songs_per_sec": S / t_native, "simlog_messages_per_sec": int(dt.numel()) / t_native, "songs": S, "log_lines_per_song": n_lines,
            "host_threads": os.cpu_count() or 1, "api": "sim_log_to_midi.sim_logs_to_event_batch -> mmg_simlog_batch_to_events (pinned dt / meta / offsets of one H2D copy)",
            "cpu_baseline": {"value": 1.0 / t_py, "unit": "songs/s", "cores": 1, "kind": "port",
                             "sample": f"{n_py} songs through the Python mirror of MidiGenerator (sim_log_to_midi.sim_log_to_event_stream)"},
            "smf_messages_per_sec": len(ev) / t_smf, "smf_messages": len(ev),
            "smf_cpu_baseline": {"value": len(ev) / t_smf_py, "unit": "messages/s", "cores": 1, "kind": "port", "sample": "the same file through oracle/smf_oracle.py"}}


def run(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    precision = args.precision or "bf16"
    B = args.batch or (4096 if precision == "fp32" else 16384)      # BASELINE config 3: global batch 16 384 at N = 1, weak scaling beyond
    W = 50
    peaks = _peaks()
    sync = torch.cuda.synchronize

    torch.manual_seed(1234 + rank)
    mmgan = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, W), input_dim=50, output_dim=20, instrument=0, start=100,
                             end=150, device=device)
    mmgan.train()
    trainer = MMGANTrainer(mmgan, lr=0.01, precision=precision, max_batch=B, inner_rng="device", sync_bn=getattr(args, "sync_bn", False))

    # ---- host (pinned) and device copies of one step's inputs
    h = {k: _synth_rolls_u8(B, W, 100 * rank + i, "cpu").pin_memory() for i, k in enumerate(("real", "fake_d", "fake_g"))}
    h["beats"] = (25.0 * torch.rand(B, 50)).pin_memory()
    d = {k: v.to(device) for k, v in h.items()}
    noise = [torch.randn(B, 50, device=device) for _ in range(4)]

    def step_resident(i):
        return trainer.step(noise[0], noise[1], d["beats"], d["real"], d["fake_d"], d["fake_g"], noise[2], noise[3])

    # end-to-end arm: the real rolls / beats are gathered from a training set resident in HBM (indices come from the host sampler);
    # the fake songs -- what the host DES bridge produces for the two generator forwards -- arrive from the HOST every step, as note-event
    # streams rasterised on the device (primary: generate_piano_roll is part of this path) or as ready-made uint8 rolls (secondary)
    ds_n = 2 * B
    dataset = (_synth_rolls_u8(ds_n, W, 7777 + rank, "cpu").to(device), (25.0 * torch.rand(ds_n, 50)).to(device))
    pipe_rolls = HostBatchPipeline(trainer, h, dataset=dataset)
    hb_rolls = [dict(fake_d=h["fake_d"], fake_g=h["fake_g"], real_idx=torch.randint(0, ds_n, (B,)).pin_memory()) for _ in range(4)]
    ev = [_synth_fake_events(B, 31 * rank + i) for i in range(4)]
    hb_ev = [dict(fake_d_events=ev[i], fake_g_events=ev[(i + 1) % 4], real_idx=hb_rolls[i]["real_idx"]) for i in range(4)]
    pipe_ev = HostBatchPipeline(trainer, hb_ev[0], dataset=dataset, raster=(100, 0, W))
    # the headline end-to-end arm also delivers the outputs of BOTH generator forwards of every iteration to pinned host memory, where the
    # unchanged host DES consumes them (matrix_sim_process.py:28-29)
    pipe_full = HostBatchPipeline(trainer, hb_ev[0], dataset=dataset, raster=(100, 0, W), g_out_host=True) if precision == "bf16" else None

    def barrier():
        if world > 1:
            dist.barrier()
        sync()

    def measure(fn):
        for i in range(args.warmup):
            fn(i)
        barrier()
        l0 = N.lib().mmg_launch_count() + trainer.replayed_launches
        sec = _timed(fn, args.steps, barrier)
        launches = N.lib().mmg_launch_count() + trainer.replayed_launches - l0
        if world > 1:
            t = torch.tensor([sec], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = t.item()
        return sec, launches

    import bench as _b
    clocks = _b.ClockSampler(local)
    clocks.__enter__()                                   # sampled over the resident and the end-to-end timed regions
    sec, launches = measure(step_resident)
    # ---- end to end: HOST (pinned) inputs every step through the public API, H2D inside the timed region
    def measure_e2e(pipe, hb):
        for _ in pipe.run([hb[i % 4] for i in range(max(args.warmup, 4))]):      # both staging slots reach graph replay before the timed region
            pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in pipe.run([hb[i % 4] for i in range(args.steps)]):
            pass
        e1.record()
        barrier()
        s_ = e0.elapsed_time(e1) * 1e-3
        if world > 1:
            t = torch.tensor([s_], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ = t.item()
        return s_

    sec_e2e_nog = measure_e2e(pipe_ev, hb_ev)
    sec_e2e = measure_e2e(pipe_full, hb_ev) if pipe_full is not None else sec_e2e_nog
    sec_e2e_rolls = measure_e2e(pipe_rolls, hb_rolls)
    clocks.__exit__(None, None, None)
    rolls = B * world * args.steps
    value, e2e = rolls / sec, rolls / sec_e2e
    h2d = pipe_ev.h2d_bytes
    d2h = pipe_full.d2h_bytes if pipe_full is not None else 8

    # ---- roofline of the dominant kernel, timed alone on this stream with CUDA events (after the step measurements: it touches D's grads)
    MAC_FWD, MAC_BWD = 3977216, 7135232                      # per roll and D pass
    if precision == "bf16":
        tcd = trainer.tc
        Bk = min(B, tcd.cap)
        xk, dlk = d["real"][:Bk], torch.full((Bk,), 1.0 / Bk, device=device)
        lossk = torch.zeros(1, device=device)
        cands = {"disc_pass_fused_kernel (one whole D pass: conv1, conv2, fc, BCE, fc', conv2 wgrad+dgrad, conv1 wgrad; one persistent tcgen05 kernel, bf16)":
                     (lambda i: tcd.pass_fused(xk, 1.0, lossk, want_logits=False), 2.0 * (MAC_FWD + MAC_BWD) * Bk, Bk * 12800.0),
                 "disc_bwd_fused_kernel (D backward of one pass: fc', conv2 wgrad+dgrad, conv1 wgrad; tcgen05, bf16)": (lambda i: tcd.backward(dlk), 2.0 * MAC_BWD * Bk,
                                                                                                                   Bk * (1690 * 16 + 429 * 128 + 429 * 64.0)),
                 "disc_fwd_fused_kernel (D forward of one pass: conv1, conv2, fc; tcgen05, bf16)": (lambda i: tcd.forward(xk), 2.0 * MAC_FWD * Bk,
                                                                                                     Bk * (12800 + 1690 * 16 + 429 * 128 + 429 * 64.0))}
    else:
        Bk = min(B, 4096)
        x1 = torch.randn(Bk, 16, 64, 25, device=device)
        w2, b2 = mmgan.discriminator.conv2.weight.detach(), mmgan.discriminator.conv2.bias.detach()
        y2 = torch.empty(Bk, 32, 32, 12, device=device)
        cands = {"conv2d_fwd_kernel (D conv2, fp32 SIMT)": (lambda i: N.call("mmg_conv2d_fwd_f32", N.ptr(x1), N.ptr(w2), N.ptr(b2), N.ptr(y2), Bk, 16, 64, 25, 32, 4, 4, 2, 1, 1,
                                                                                 N.stream()), 2.0 * 3145728 * Bk, None)}
    timed = {}
    for kname, (fn, kflops, kbytes) in cands.items():
        for _ in range(3):
            fn(0)
        timed[kname] = (_timed(fn, 10, sync) / 10, kflops, kbytes)
    if precision == "bf16" and trainer.one_kernel_pass:
        tcd.forward(xk)                                       # (the two-kernel path is timed for comparison only; leave consistent activations behind)
        kname = next(k for k in timed if k.startswith("disc_pass_fused_kernel"))      # the kernel the step launches 3 times per iteration
    else:
        kname = max(timed, key=lambda k: timed[k][0])        # the longest one is the dominant kernel of the step (3 launches each per iteration)
    ksec, kflops, kbytes = timed[kname]
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    if os.path.exists(tpath):                                # dram__bytes_read+write of one launch from the committed ncu --set full capture
        tj = json.load(open(tpath))
        t = tj.get(kname.split(" ")[0])
        if t and t.get("batch") == Bk:
            traffic = t["dram_bytes"]
        elif str(Bk) in tj.get("by_batch", {}) and kname.split(" ")[0] in tj["by_batch"][str(Bk)]:
            traffic = tj["by_batch"][str(Bk)][kname.split(" ")[0]]["dram_bytes"]
    roofline = {"bound": "tensor", "kernel": kname, "achieved": kflops / ksec / 1e12, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": kflops / ksec / 1e12 / peaks["bf16_tflops_sustained"], "traffic": traffic, "peak_src": peaks["src"],
                "kernel_us": ksec * 1e6, "kernel_algorithmic_flops": kflops, "kernel_algorithmic_bytes": kbytes,
                "kernel_hbm_gbs": kbytes / ksec / 1e9 if kbytes else None,
                "other_kernels": {k.split(" ")[0]: {"us": v[0] * 1e6, "tflops": v[1] / v[0] / 1e12} for k, v in timed.items() if k != kname},
                "step_achieved_tflops_per_gpu": FLOP_PER_ROLL * B * args.steps / sec / 1e12,
                "step_frac_of_bf16_peak": FLOP_PER_ROLL * B * args.steps / sec / 1e12 / peaks["bf16_tflops_sustained"]}
    if kname.startswith("disc_pass_fused_kernel"):
        # what actually bounds this kernel: the tensor cores read their operands from shared memory at 128 B per clock and SM
        # (ncu l1tex__data_pipe_tc_wavefronts_mem_shared); the MMA program of one sample reads 42 x 4608 + 68 x 5120 + 89 x 6144 bytes
        op_bytes = (42 * 4608 + 68 * 5120 + 89 * 6144) * float(Bk)
        sm_clock = 1.965e9 if not clocks.summary().get("sm_mhz") else clocks.summary()["sm_mhz"] * 1e6
        feed_peak = 148 * 128 * sm_clock
        roofline["operand_feed"] = {"bound": "shared memory -> tensor core operand reads", "achieved_tbs": op_bytes / ksec / 1e12, "peak_tbs": feed_peak / 1e12,
                                    "frac": op_bytes / ksec / feed_peak, "operand_bytes_per_roll": op_bytes / Bk,
                                    "note": "N = 16 / 32 / 64 MMAs: the A tile (4 KB per M128 x K16 MMA) is re-read once per tap; ncu: 61-66 % of the pipe's peak"}

    line = {"metric": METRIC, "value": value, "unit": "rolls/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": "MM-GAN G+D training iteration (network_tests.py:292-315), DES excluded", "batch_per_gpu": B, "global_batch": B * world,
                       "roll_size": [2, 128, W], "adj_size": [64, 64], "z_dim": 50, "parallelism": f"dp{world}", "precision": precision, "cuda_graph": bool(trainer.use_graph), "generator_bn": "sync (global batch)" if trainer.sync_bn else "per-replica",
                       "l2": "inputs larger than L2 (per-step working set > 126 MB)" if B * 3 * 12800 > 126e6 else "working set may fit L2"},
            "e2e": {"value": e2e, "unit": "rolls/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": sec_e2e / args.steps * 1e3,
                    "api": "trainer.HostBatchPipeline.run(g_out_host=True): pinned host batches = the note-event streams of the simulated songs (what the DES bridge hands to "
                           "generate_piano_roll: 320 messages per song, 12 B per message) + sampler indices; events copied H2D and rasterised on the device "
                           "(mmg_raster_piano_roll, bit-exact) into the uint8 fake rolls, real rolls / beats gathered from the HBM-resident training set, "
                           "real rolls read by index inside the discriminator kernel, copies + rasterisation of batch i+1 overlapped with the iteration of batch i; "
                           "the iteration runs as the public segments generators / generators / d_step / g_step and the outputs of BOTH generator forwards "
                           "((B,1,64,64) + (B,20) fp32 each: what matrix_to_midi reads, matrix_sim_process.py:28-29) go to pinned host memory on a D2H stream underneath the "
                           "discriminator passes; the two losses of every step are read back too (waited for one iteration later).  The DES itself is excluded: the fake songs "
                           "are synthetic and do not depend on these generator outputs",
                    "no_g_d2h": {"value": rolls / sec_e2e_nog, "unit": "rolls/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": sec_e2e_nog / args.steps * 1e3,
                                 "api": "the same pipeline without the generator-output read-back (round 1's end-to-end figure)"},
                    "rolls_u8_variant": {"value": rolls / sec_e2e_rolls, "unit": "rolls/s", "h2d_bytes_per_step": pipe_rolls.h2d_bytes,
                                         "ms_per_step": sec_e2e_rolls / args.steps * 1e3,
                                         "api": "same pipeline fed with ready-made uint8 fake rolls (12.8 KB per roll over PCIe)"}},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline}

    if rank == 0 and world == 1:
        if not args.no_cpu_baseline:
            import mmgan_oracle as mo
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            Bc = 256
            sd = mo.synth_state(mo.mmgan_shapes(), seed=0, d_scale=0.25)
            inp = mo.synth_inputs(Bc, seed=1)
            adam = {}
            mo.mmgan_iteration(sd, adam, inp)
            t0 = time.perf_counter()
            n_it = 0
            while time.perf_counter() - t0 < 10.0 and n_it < 50:
                mo.mmgan_iteration(sd, adam, inp)
                n_it += 1
            cs = (time.perf_counter() - t0) / n_it
            line["cpu_baseline"] = {"value": Bc / cs, "unit": "rolls/s", "cores": cores, "kind": "port",
                                    "sample": f"{n_it} iterations of batch {Bc} through oracle/mmgan_oracle.py (fp32, DES excluded)"}
        if not args.no_raster:
            line["raster"] = _raster_leg(device, peaks, quick=False)
            mmgan.eval()
            line["inference_sweep"] = _inference_leg(mmgan, device)
            line["gandes"] = _gandes_leg(device)
            line["gandes_mel"] = _mel_leg(device, peaks)
            try:
                line["host_front_end"] = _host_front_end_leg()
            except Exception as e:                  # a secondary, CPU-only leg must never cost the headline line
                line["host_front_end"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # teardown: CUDA graphs that hold NCCL kernels (captured SyncBN all-reduces) go before the communicator; the JSON line is out, so a
        # stalled communicator teardown must not keep the job alive
        import gc
        import threading
        trainer._graphs.clear()
        trainer._seg_graphs.clear()
        gc.collect()
        torch.cuda.synchronize()
        killer = threading.Timer(20.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()

"""Golden-vector generator (test infrastructure; runs ONLY in the build container).

Executes the UNMODIFIED reference code from /root/reference (through the import
harness in oracle/_refimport.py) on seeded synthetic inputs and freezes what it
returns under tests/golden/.  The GPU box has no /root/reference: tests there read
only the committed .npz files.

    python oracle/make_golden.py raster     -> tests/golden/raster_cases.npz
    python oracle/make_golden.py mmgan      -> tests/golden/mmgan_b16.npz, mmgan_b4_small.npz
    python oracle/make_golden.py gandes     -> tests/golden/gandes_b3.npz   (separate process: module-name clash)
    python oracle/make_golden.py simlog     -> tests/golden/simlog_cases.npz

Reference entry points exercised:
  MMGAN_MIDI_DES/datasets.py:13-70      generate_piano_roll (via the mido / pretty_midi shim)
  MMGAN_MIDI_DES/network_tests.py:58-206 Generator, BeatGenerator, DiscriminatorCNN, MultiModalGAN
  MMGAN_MIDI_DES/network_tests.py:248-315 criterion, the two Adam optimisers and the loop body
  GAN_DES/SIMNN.py:62-142,256-331        Generator, Discriminator, loop body
  MMGAN_MIDI_DES/sim_log_to_midi.py:13-277 MidiGenerator, LogLineProcessor, process_adjsim_log (mido build-side shim)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

import _refimport as R          # noqa: E402
import raster_oracle as ro      # noqa: E402
import mmgan_oracle as mo       # noqa: E402

KINDS = {"cc": 0, "on": 1, "off": 2}
NAMES = {0: "control_change", 1: "note_on", 2: "note_off"}


def _shim_file(dt, kind, pitch, vel, beats=()):
    ev = [R.ShimMessage(NAMES[int(k)], float(t), int(p), int(v)) for t, k, p, v in zip(dt, kind, pitch, vel)]
    return R.ShimMidiFile(ev, beats=beats)


def make_raster():
    _, ds = R.import_mmgan()
    import io, contextlib
    cases = {}

    def add(name, msgs=None, arrays=None, **kw):
        if arrays is None:
            dt = np.array([m[1] for m in msgs], dtype=np.float64)
            kind = np.array([KINDS[m[0]] for m in msgs], dtype=np.uint8)
            pitch = np.array([m[2] if len(m) > 2 else 0 for m in msgs], dtype=np.uint8)
            vel = np.array([m[3] if len(m) > 3 else 0 for m in msgs], dtype=np.uint8)
        else:
            dt, kind, pitch, vel = arrays
        args = dict(sequence_length=100, beats_length=50, start=0, end=50)
        args.update(kw)
        with contextlib.redirect_stdout(io.StringIO()):
            roll, dur, beats = ds.generate_piano_roll(_shim_file(dt, kind, pitch, vel, beats=[0.5, 1.0]), **args)
        sl = -1 if args["sequence_length"] is None else args["sequence_length"]
        cases[name + ".dt"] = dt
        cases[name + ".meta"] = ro.pack_meta(kind, pitch, vel)
        cases[name + ".args"] = np.array([sl, args["start"], args["end"]], dtype=np.int64)
        cases[name + ".roll"] = roll.astype(np.float32)
        cases[name + ".dur"] = dur.astype(np.float32)
        assert roll.dtype == np.float64 and np.array_equal(roll, roll.astype(np.float32))
        assert np.array_equal(beats[:2], [0.5, 1.0]) and len(beats) == 50 and not beats[2:].any()

    # SURVEY.md Appendix A, K1-K8
    K1 = [("on", 0.4, 60, 80), ("on", 0.1, 64, 70), ("off", 2.0, 60), ("on", 0.0, 60, 0), ("off", 1.0, 64),
          ("on", 45.9, 72, 100), ("on", 0.7, 73, 101), ("on", 10, 74, 102)]
    add("K1", K1)
    add("K2", [("on", 49.4, 60, 80), ("on", 1.0, 61, 81), ("on", 1.0, 62, 82)])
    add("K3", K1, start=100, end=150)
    add("K4", K1, start=2, end=52)
    add("K5", [("on", 3, 60, 90), ("on", 5, 61, 91), ("off", 4, 60), ("off", 8, 61), ("off", 5, 62), ("cc", 5),
               ("on", 1, 63, 99), ("off", 9, 61)], sequence_length=100, start=0, end=10)
    add("K6", [("on", 3, 60, 90), ("cc", 4), ("on", 0, 61, 91), ("off", 1, 60)], sequence_length=7, start=0, end=10)
    add("K7", [("on", 3, 60, 90), ("off", 26, 60), ("off", 1, 60)], sequence_length=None, start=0, end=10)
    add("K8", [("on", 2, 60, 50), ("on", 0.2, 60, 70), ("on", 2, 60, 0), ("off", 3, 60)], start=0, end=10)
    add("empty", [])
    add("half_even", [("on", 0.5, 10, 1), ("on", 1.0, 11, 2), ("on", 1.0, 12, 3), ("on", 1.0, 13, 4), ("off", 1.0, 10),
                      ("off", 0.0, 11), ("off", 0.0, 13)], start=0, end=10)   # t=.5,1.5,2.5,3.5,4.5 -> 0,2,2,4,4
    add("pitch_edges", [("on", 1, 0, 127), ("on", 1, 127, 1), ("off", 3, 0), ("off", 0, 127)], start=0, end=8)
    add("end_ge_128", K1, sequence_length=200, start=0, end=130)
    add("start_gt0_end_ge128", K1, sequence_length=200, start=30, end=160)
    # random streams: the reference call shape (S=100, W=50), the notebook shape (S=W=300) and odd windows
    rng = np.random.default_rng(7)
    for i, (E, T, kw) in enumerate([(400, 60.0, dict()), (400, 120.0, dict()), (3000, 300.0, dict(sequence_length=300, start=0, end=300)),
                                    (800, 80.0, dict(sequence_length=64, start=0, end=70)), (800, 40.0, dict(sequence_length=90, start=5, end=45)),
                                    (500, 30.0, dict(sequence_length=None, start=0, end=17)), (2000, 200.0, dict(sequence_length=150, start=0, end=140)),
                                    (50, 10.0, dict(sequence_length=100, start=0, end=0))]):
        dt = rng.exponential(T / E, size=E)
        if i == 1:
            dt = np.round(dt * 2) / 2          # many exact .5 boundaries -> half-even rounding matters
        u = rng.random(E)
        kind = np.where(u < 0.3, 1, np.where(u < 0.6, 2, 0)).astype(np.uint8)
        pitch = rng.integers(0, 128, size=E).astype(np.uint8) if i % 2 else rng.integers(21, 109, size=E).astype(np.uint8)
        vel = rng.integers(0, 128, size=E).astype(np.uint8)
        add(f"rand{i}", arrays=(dt, kind, pitch, vel), **kw)
    names = sorted({k.split(".")[0] for k in cases})
    cases["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "raster_cases.npz"), **cases)
    # K9: the reference raises ValueError for a non-path, non-MidiFile input (datasets.py:24)
    try:
        ds.generate_piano_roll(123)
        raise SystemExit("K9 failed")
    except ValueError as e:
        assert "midi_input must be a file path" in str(e)
    print("raster goldens:", names)


def _run_mmgan_case(nt, B, adj, out_dim, seed, iters, full_g1):
    shapes = mo.mmgan_shapes(adj_size=(adj, adj), output_dim=out_dim)
    sd0 = mo.synth_state(shapes, seed=seed, d_scale=0.25)
    mmgan = nt.MultiModalGAN(z_dim=50, adj_size=(adj, adj), roll_size=(2, 128, 50), input_dim=50, output_dim=out_dim,
                             instrument=0, start=100, end=150, device="cpu")
    assert list(mmgan.state_dict().keys()) == list(shapes.keys())
    mmgan.load_state_dict(sd0)
    criterion = torch.nn.BCEWithLogitsLoss()                                                   # network_tests.py:248
    gen_opt = torch.optim.Adam(list(mmgan.generator1.parameters()) + list(mmgan.generator2.parameters()), lr=0.01)  # :253
    disc_opt = torch.optim.Adam(mmgan.discriminator.parameters(), lr=0.01)                     # :254
    mmgan.train()
    g_out = {}
    mmgan.generator1.register_forward_hook(lambda m, i, o: g_out.__setitem__("g1", o.detach().clone()))
    mmgan.generator2.register_forward_hook(lambda m, i, o: g_out.__setitem__("g2", o.detach().clone()))
    gold = {"meta": np.array([B, adj, out_dim, seed, iters], dtype=np.int64)}
    for it in range(iters):
        inp = mo.synth_inputs(B, seed=seed * 1000 + it)
        pre = f"it{it}."
        # the fake rolls the host DES would have returned: list of B float64 (2,128,W) arrays
        fake = {"d": [a.double().numpy() for a in inp["fake_d"]], "g": [a.double().numpy() for a in inp["fake_g"]]}
        which = {"k": "d"}
        nt.matrix_to_midi = lambda *a, **k: (fake[which["k"]], 0)
        piano_roll, durations, beats = inp["real"][:, 0], inp["real"][:, 1], inp["beats"]
        noise1, noise2 = inp["noise1"], inp["noise2"]
        real = torch.ones(B)
        fake_label = torch.zeros(B)
        real_data = torch.stack([piano_roll, durations]).permute(1, 0, 2, 3)                   # :290
        # ---- :293-308
        disc_opt.zero_grad()
        torch.manual_seed(4242 + it)            # pins the randn drawn INSIDE Generator.forward (:83-84)
        inner = torch.randn(B, 50)
        torch.manual_seed(4242 + it)
        fake_output, failed = mmgan(noise1, noise2, beats, it + 1, False)
        gold[pre + "inner_d"] = inner.numpy()
        gold[pre + "g1_d"], gold[pre + "g2_d"] = g_out["g1"].numpy(), g_out["g2"].numpy()
        disc_fake_loss = criterion(fake_output.squeeze(), fake_label)
        logit_real = mmgan.discriminator(real_data)
        disc_real_loss = criterion(logit_real.squeeze(), real)
        disc_loss = disc_fake_loss + disc_real_loss
        disc_loss.backward()
        for k, p in mmgan.discriminator.named_parameters():
            gold[pre + "grad_d.discriminator." + k] = p.grad.detach().clone().numpy()
        disc_opt.step()
        for k, p in mmgan.discriminator.named_parameters():
            gold[pre + "param_d.discriminator." + k] = p.detach().clone().numpy()
        gold[pre + "logit_fake_d"] = fake_output.detach().numpy()
        gold[pre + "logit_real"] = logit_real.detach().numpy()
        gold[pre + "disc_loss"] = disc_loss.detach().numpy()
        # ---- :311-315
        gen_opt.zero_grad()
        which["k"] = "g"
        torch.manual_seed(9000 + it)
        inner_g = torch.randn(B, 50)
        torch.manual_seed(9000 + it)
        fake_output, failed = mmgan(noise1, noise2, beats, it + 1)          # make_dot_png defaults True (:176,:312)
        gold[pre + "inner_g"] = inner_g.numpy()
        gold[pre + "g1_g"], gold[pre + "g2_g"] = g_out["g1"].numpy(), g_out["g2"].numpy()
        gen_loss = criterion(fake_output.squeeze(), real)
        gen_loss.backward()
        gen_opt.step()
        assert all(p.grad is None for p in mmgan.generator1.parameters())
        assert all(p.grad is None for p in mmgan.generator2.parameters())
        assert len(gen_opt.state) == 0
        for k, p in mmgan.discriminator.named_parameters():
            gold[pre + "grad_g.discriminator." + k] = p.grad.detach().clone().numpy()
        gold[pre + "logit_fake_g"] = fake_output.detach().numpy()
        gold[pre + "gen_loss"] = gen_loss.detach().numpy()
        if not full_g1:
            for nm in ("g1_d", "g1_g"):
                a = gold.pop(pre + nm)
                gold[pre + nm + ".sub"] = a[:, :, ::4, ::4].copy()
                gold[pre + nm + ".sum"] = np.array([a.astype(np.float64).sum(), (a.astype(np.float64) ** 2).sum()])
    for k, v in mmgan.state_dict().items():
        if "running" in k or "num_batches" in k:
            gold["final." + k] = v.detach().clone().numpy()
    # eval-mode generator outputs (generate_midi path, :198-203) after the training iterations
    mmgan.generator1.eval(); mmgan.generator2.eval()
    inp = mo.synth_inputs(B, seed=seed * 1000 + 77)
    torch.manual_seed(555)
    gold["eval.inner"] = torch.randn(B, 50).numpy()
    torch.manual_seed(555)
    with torch.no_grad():
        g1 = mmgan.generator1(inp["noise1"])
        g2 = mmgan.generator2(inp["noise2"], inp["beats"])
    gold["eval.g2"] = g2.numpy()
    gold["eval.g1.sub"] = g1.numpy()[:, :, ::4, ::4].copy()
    gold["eval.g1.sum"] = np.array([g1.double().sum().item(), (g1.double() ** 2).sum().item()])
    return gold


def make_mmgan():
    nt, _ = R.import_mmgan()
    torch.set_num_threads(1)
    g = _run_mmgan_case(nt, B=16, adj=64, out_dim=20, seed=3, iters=2, full_g1=False)
    np.savez_compressed(os.path.join(GOLD, "mmgan_b16.npz"), **g)
    g = _run_mmgan_case(nt, B=4, adj=16, out_dim=16, seed=5, iters=2, full_g1=True)
    np.savez_compressed(os.path.join(GOLD, "mmgan_b4_small.npz"), **g)
    print("mmgan goldens written")


def make_gandes():
    sim = R.import_gandes()
    torch.set_num_threads(1)
    B = 3
    gshapes, dshapes = mo.gandes_shapes()
    gsd = mo.synth_state(gshapes, seed=11)
    dsd = mo.synth_state(dshapes, seed=12)
    gen, disc = sim.Generator(), sim.Discriminator()
    assert list(gen.state_dict().keys()) == list(gshapes.keys()), (list(gen.state_dict().keys()), list(gshapes.keys()))
    assert list(disc.state_dict().keys()) == list(dshapes.keys())
    gen.load_state_dict(gsd); disc.load_state_dict(dsd)
    criterion = torch.nn.BCEWithLogitsLoss()                                           # SIMNN.py:257
    gen_opt = torch.optim.Adam(gen.parameters(), lr=2e-5, betas=(0.5, 0.999))          # :258
    disc_opt = torch.optim.Adam(disc.parameters(), lr=2e-5, betas=(0.5, 0.999))        # :259
    rng = np.random.default_rng(13)
    gold = {"meta": np.array([B], dtype=np.int64)}
    real = torch.from_numpy(rng.standard_normal((B, 128, 216)).astype(np.float32))
    fake = torch.from_numpy(rng.standard_normal((B, 128, 216)).astype(np.float32))
    noise = torch.from_numpy(rng.standard_normal((B, 100, 1, 1)).astype(np.float32))
    # ---- :283-316
    disc_opt.zero_grad()
    p_real = disc(real).reshape(-1)
    l_real = criterion(p_real, torch.ones(B) * 0.9)
    gen_out = gen(noise)
    p_fake = disc(fake.detach()).reshape(-1)
    l_fake = criterion(p_fake, torch.ones(B) * 0.1)
    d_loss = l_fake + l_real
    d_loss.backward()
    sl = {"fc1.weight": (slice(0, 128, 16), slice(0, None, 97))}
    for k, p in disc.named_parameters():
        g = p.grad.detach().clone().numpy()
        gold["grad_d." + k + ".sum"] = np.array([g.astype(np.float64).sum(), (g.astype(np.float64) ** 2).sum()])
        gold["grad_d." + k] = g[sl[k]].copy() if k in sl else g
    disc_opt.step()
    for k, p in disc.named_parameters():
        a = p.detach().clone().numpy()
        gold["param_d." + k] = a[sl[k]].copy() if k in sl else a
    gold["p_real"], gold["p_fake"], gold["disc_loss"] = p_real.detach().numpy(), p_fake.detach().numpy(), d_loss.detach().numpy()
    gold["gen_out"] = gen_out.detach().numpy()
    # ---- :322-331
    gen_opt.zero_grad()
    p_g = disc(fake).squeeze()
    g_loss = criterion(p_g, torch.ones(B))
    g_loss.backward()
    gen_opt.step()
    assert all(p.grad is None for p in gen.parameters())
    gold["p_fake_g"], gold["gen_loss"] = p_g.detach().numpy(), g_loss.detach().numpy()
    for k, v in gen.state_dict().items():
        if "running" in k or "num_batches" in k:
            gold["final.gen." + k] = v.detach().clone().numpy()
    gen.eval()
    with torch.no_grad():
        gold["eval.gen_out"] = gen(noise).numpy()
    gold["real"], gold["fake"], gold["noise"] = real.numpy(), fake.numpy(), noise.numpy()
    np.savez_compressed(os.path.join(GOLD, "gandes_b3.npz"), **gold)
    print("gandes golden written")


def synth_sim_log(rng, n_lines, t_max, n_servers=16, n_customers=40, junk=0.1):
    """Log lines in the simulator's 'Music' logging format (simulation_v3.py:341,546,604,617): time - customer - server - event."""
    t = np.sort(rng.random(n_lines) * t_max)
    lines = []
    for i in range(n_lines):
        if rng.random() < junk:
            lines.append("INFO:root:queue length 3\n")                      # lines the regex rejects
            continue
        ev = "arrival" if rng.random() < 0.55 else "departure"
        tt = f"{t[i]:.4f}" if rng.random() < 0.8 else f"{int(t[i])}"
        lines.append(f"INFO:root:{tt} - {int(rng.integers(0, n_customers))} - {int(rng.integers(0, n_servers))} - {ev}\n")
    return lines


def make_simlog():
    import contextlib, io, tempfile
    m = R.import_simlog()
    rng = np.random.default_rng(2024)
    out, names = {}, []
    # (name, lines, t_max, gen2[0:6], generate, start, end)
    specs = [("hundred", 200, 60.0, (0.2, 0.3, 0.5, 0.7, 0.5, 0.4), False, 0, 50),
             ("not_multiple_of_100", 150, 60.0, (0.2, 0.3, 0.5, 0.7, 0.5, 0.4), False, 0, 50),
             ("generate", 150, 60.0, (0.2, 0.3, 0.5, 0.7, 0.5, 0.4), True, 0, 50),
             ("slow_tempo", 300, 40.0, (0.21, 0.35, 0.5, 0.9, 0.95, 0.9), False, 0, 50),
             ("very_slow_tempo", 100, 20.0, (0.2, 0.3, 0.7, 0.3, 9.0, 0.05), False, 0, 50),
             ("track_cap", 1500, 150.0, (0.2, 0.3, 0.5, 0.6, 0.3, 0.5), False, 0, 50),
             ("late_times", 400, 320.0, (0.2, 0.2, 0.2, 0.6, 0.1, 0.5), False, 0, 50),
             ("window", 200, 60.0, (0.2, 0.3, 0.5, 0.7, 0.9, 0.4), True, 100, 150),
             ("sparse_skips", 300, 60.0, (0.7, 0.9, 1.1, 0.7, 0.6, 0.0), False, 0, 50),
             ("zero_tempo", 100, 30.0, (0.2, 0.3, 0.5, 0.7, 0.0, 0.4), False, 0, 50),
             ("too_long", 5100, 190.0, (0.2, 0.3, 0.5, 0.7, 0.5, 0.4), False, 0, 50)]
    for name, n_lines, t_max, g6, generate, start, end in specs:
        lines = synth_sim_log(rng, n_lines, t_max)
        gen2 = np.concatenate([np.array(g6, dtype=np.float32), rng.random(4).astype(np.float32)])
        instruments = rng.integers(0, 100, 16)
        note_levels = rng.integers(30, 100, 16)
        seen = {}
        real_gpr = m.generate_piano_roll

        def spy(midi, **kw):
            seen["msgs"] = [(x.type, float(x.time), int(getattr(x, "note", 0)), int(getattr(x, "velocity", 0))) for x in midi]
            return real_gpr(midi, **kw)

        m.generate_piano_roll = spy
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as td:
            os.makedirs(os.path.join(td, "logs"))
            os.makedirs(os.path.join(td, "adj_sim_outputs", "midi"))
            open(os.path.join(td, "logs", "simulation.log"), "w").writelines(lines)
            os.chdir(td)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    roll, dur, _ = m.process_adjsim_log(instruments=instruments, note_levels=note_levels, gen2_output=gen2, count=0, start=start, end=end,
                                                        generate=generate)
            finally:
                os.chdir(cwd)
                m.generate_piano_roll = real_gpr
        kinds = {"note_on": 1, "note_off": 2}
        msgs = seen["msgs"]
        names.append(name)
        out[name + ".lines"] = np.array(lines)
        out[name + ".gen2"] = gen2
        out[name + ".instruments"] = instruments
        out[name + ".note_levels"] = note_levels
        out[name + ".args"] = np.array([int(generate), start, end])
        out[name + ".dt"] = np.array([x[1] for x in msgs], dtype=np.float64)
        out[name + ".meta"] = np.array([(kinds[x[0]] | (x[2] << 8) | (x[3] << 16)) if x[0] in kinds else 0 for x in msgs], dtype=np.uint32)
        out[name + ".roll"] = np.asarray(roll)
        out[name + ".dur"] = np.asarray(dur)
        print(name, "messages", len(msgs), "notes", int((out[name + ".meta"] != 0).sum()), "roll nnz", int((roll != 0).sum()), "dur nnz", int((dur != 0).sum()))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "simlog_cases.npz"), **out)



def make_artifacts():
    """Shipped artefacts of the reference as drop-in fixtures: the MM-GAN checkpoint
    MMGAN_MIDI_DES/models/mmgan_64_64_epoch_1.pth run through the UNMODIFIED reference classes
    (network_tests.py:58-206: eval-mode generators as in generate_midi :198-203, discriminator forward and one
    BCE backward per target), and the 30 shipped .mid files as parsed message streams.
      tests/golden/ckpt_epoch1.npz  the state dict (so the GPU box can rebuild the model) + reference outputs
      tests/golden/disc_epoch1.npz  SURVEY 8d's tolerance probe: discriminator of that checkpoint, B = 256 synthetic rolls,
                                    fp32 reference logits / loss / gradients for target 0 and target 1
      tests/golden/ckpt_keys.npz    key / shape / checksum tables of all shipped MM-GAN checkpoints (load contract)
    GAN-DES: python oracle/make_golden.py artifacts_gandes -> tests/golden/ckpt_gandes_gen.npz (separate process)."""
    nt, _ = R.import_mmgan()
    ref_root = os.path.join(R.REF_ROOT, "MMGAN_MIDI_DES")
    ck_path = os.path.join(ref_root, "models", "mmgan_64_64_epoch_1.pth")
    sd = torch.load(ck_path, map_location="cpu")
    mmgan = nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device="cpu")
    mmgan.load_state_dict(sd)                                                      # network_tests.py:240-245
    gold = {"sd." + k: v.numpy() for k, v in sd.items()}
    gold["keys"] = np.array(list(sd.keys()))
    B = 8
    inp = mo.synth_inputs(B, seed=2024)
    mmgan.generator1.eval(); mmgan.generator2.eval()                                # generate_midi, :199-200
    torch.manual_seed(777)
    gold["eval.inner"] = torch.randn(B, 50).numpy()
    torch.manual_seed(777)
    with torch.no_grad():
        g1 = mmgan.generator1(inp["noise1"])                                        # draws the inner randn
        g2 = mmgan.generator2(inp["noise2"], inp["beats"])
        logit = mmgan.discriminator(inp["real"])
    gold["eval.g1"], gold["eval.g2"], gold["disc.logit_real"] = g1.numpy()[:, :, ::2, ::2].copy(), g2.numpy(), logit.numpy()
    gold["eval.g1.sum"] = np.array([g1.double().sum().item(), (g1.double() ** 2).sum().item()])
    gold["seed"] = np.array([2024, B], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, "ckpt_epoch1.npz"), **gold)
    # ---- SURVEY 8d probe conditions: this discriminator, B = 256
    Bp = 256
    D = mmgan.discriminator
    crit = torch.nn.BCEWithLogitsLoss()
    probe = {"meta": np.array([Bp, 50, 8], dtype=np.int64)}                        # batch, width, roll seed (mo.synth_rolls(B, 50, seed=8))
    for k, p in D.named_parameters():
        probe["w." + k] = p.detach().numpy().copy()
    x = torch.from_numpy(mo.synth_rolls(Bp, 50, seed=8)).float()
    for y in (0.0, 1.0):
        D.zero_grad()
        lg = D(x)
        loss = crit(lg.squeeze(), torch.full((Bp,), y))
        loss.backward()
        probe[f"y{int(y)}.logits"], probe[f"y{int(y)}.loss"] = lg.detach().numpy().reshape(-1), loss.detach().numpy()
        for k, p in D.named_parameters():
            probe[f"y{int(y)}.grad." + k] = p.grad.detach().numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "disc_epoch1.npz"), **probe)
    # ---- load contract of every shipped MM-GAN checkpoint
    tab = {}
    for rel in ("mmgan_64_64_epoch_1.pth", "MAE_loss/mmgan_64_64_epoch_35.pth", "V1_bad/mmgan_64_64_epoch_50.pth"):
        s2 = torch.load(os.path.join(ref_root, "models", rel), map_location="cpu")
        tab[rel + ".keys"] = np.array(list(s2.keys()))
        tab[rel + ".shapes"] = np.array([",".join(map(str, v.shape)) for v in s2.values()])
        tab[rel + ".sums"] = np.array([v.double().sum().item() for v in s2.values()])
    np.savez_compressed(os.path.join(GOLD, "ckpt_keys.npz"), **tab)
    print("artifact goldens: ckpt_epoch1 (%d tensors), disc_epoch1 (B=%d), ckpt_keys" % (len(sd), Bp))


def make_artifacts_gandes():
    """GAN_DES/models/gen_100_*.pt through the UNMODIFIED GAN_DES/SIMNN.py Generator (eval mode, demo.ipynb cells 25-27)."""
    sim = R.import_gandes()
    import glob
    path = sorted(glob.glob(os.path.join(R.REF_ROOT, "GAN_DES", "models", "gen_100_*.pt")))[0]
    sd = torch.load(path, map_location="cpu")
    gen = sim.Generator()
    gen.load_state_dict(sd)
    gen.eval()
    rng = np.random.default_rng(99)
    noise = torch.from_numpy(rng.standard_normal((4, 100, 1, 1)).astype(np.float32))
    with torch.no_grad():
        out = gen(noise)
    gold = {"sd." + k: v.numpy() for k, v in sd.items()}
    gold["keys"] = np.array(list(sd.keys()))
    gold["noise"], gold["out"] = noise.numpy(), out.numpy()
    np.savez_compressed(os.path.join(GOLD, "ckpt_gandes_gen.npz"), **gold)
    print("artifact golden: ckpt_gandes_gen", os.path.basename(path), tuple(out.shape))



def make_midi():
    """The 30 .mid files the reference ships (SURVEY 4 / 8d config 4 "plus the 30 shipped .mid") as post-mido message streams.
    mido is absent here, so the streams come from this repo's own SMF reader (datasets.read_smf, which restates mido's merge / tempo
    rules): parity of the PARSER is unpinned; what the fixture pins is the rasteriser on real songs -- the GPU test
    rasterises these streams on the device and compares with the C oracle -- plus the facts the survey probed with its own reader
    (simulation.mid: 202 events, 106 note_on / 91 note_off, 480 ticks per beat, tempo 423 130)."""
    import glob
    root = os.path.dirname(HERE)
    sys.path.insert(0, root)
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    files = sorted(glob.glob(os.path.join(R.REF_ROOT, "**", "*.mid"), recursive=True))
    gold, names = {}, []
    for i, f in enumerate(files):
        ev = ds.read_smf(f)
        rel = os.path.relpath(f, R.REF_ROOT)
        names.append(rel)
        gold[f"s{i}.dt"], gold[f"s{i}.meta"], gold[f"s{i}.beats"] = ev.dt, ev.meta, ev.beats
    gold["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "midi_streams.npz"), **gold)
    print("midi goldens:", len(files), "files,", sum(len(gold[f"s{i}.dt"]) for i in range(len(files))), "messages")


def make_mel():
    """GAN_DES/util.py:37-87 UNMODIFIED (torchaudio of this image) on deterministic test signals (mel_oracle.synth_wave: regenerated from the
    seed by the tests, so only the outputs are stored): the three call shapes of the reference -- datasets.py:51 (110 250 samples at 22 050 Hz),
    datasets.py:88 (a 5 s split at 44 100 Hz, default sr) and a short clip -- dB spectrograms, plus the power spectrogram `_maestro` returns."""
    import mel_oracle as mel
    u = R.import_gandes_util()
    out = {"meta": np.array([[110250, 22050, 11], [220500, 44100, 12], [30000, 44100, 13], [2049 * 215 + 7, 44100, 14]], dtype=np.int64)}
    for i, (L, sr, seed) in enumerate(out["meta"]):
        w = torch.from_numpy(mel.synth_wave(int(L), int(seed)))
        db = u.get_melspectrogram_db_tensor(w, int(sr))
        out[f"db{i}"] = db.numpy().astype(np.float32)
        if i == 1:
            out["power1"] = u.get_melspectrogram_db_tensor_maestro(w, int(sr)).numpy().astype(np.float32)
        print("mel case", i, tuple(db.shape), float(db.min()), float(db.max()))
    np.savez_compressed(os.path.join(GOLD, "mel_cases.npz"), **out)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("raster", "all"):
        make_raster()
    if what in ("mmgan", "all"):
        make_mmgan()
    if what == "gandes":
        make_gandes()
    if what in ("simlog", "all"):
        make_simlog()
    if what in ("artifacts", "all"):
        make_artifacts()
    if what in ("midi", "all"):
        make_midi()
    if what == "artifacts_gandes":
        make_artifacts_gandes()
    if what == "mel":
        make_mel()
    if what == "all":
        import subprocess
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "artifacts_gandes"])
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "gandes"])
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "mel"])

"""The reference's shipped artefacts as drop-in fixtures (network_tests.py:240-245, demo.ipynb cells 25-27):
  * the three MM-GAN checkpoints and the GAN-DES generator checkpoint load into the mirror classes with strict keys -- from the
    original files where /root/reference exists (build container), from the frozen key / shape tables everywhere;
  * tests/golden/ckpt_epoch1.npz (state dict of models/mmgan_64_64_epoch_1.pth + outputs of the UNMODIFIED reference classes, frozen by
    oracle/make_golden.py artifacts): eval-mode generators and the discriminator forward on the GPU against the reference;
  * tests/golden/disc_epoch1.npz = SURVEY 8d's tolerance probe (that checkpoint's discriminator, B = 256): the bf16 one-kernel pass meets
    logits 0.5 % of scale, loss rel 1e-3 and gradients rel-L2 1e-2 on ALL six tensors;
  * tests/golden/ckpt_gandes_gen.npz: GAN-DES generator checkpoint, eval forward;
  * tests/golden/midi_streams.npz: the 30 shipped .mid files as message streams, rasterised on the device vs the C oracle."""
import glob
import os

import numpy as np
import pytest
import torch

import mmgan_oracle as mo
import raster_oracle as ro

REF = os.environ.get("MMG_REFERENCE_ROOT", "/root/reference")
HAVE_REF = os.path.isdir(os.path.join(REF, "MMGAN_MIDI_DES"))
DEV = "cuda"


def _mmgan(device):
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    return nt.MultiModalGAN(z_dim=50, adj_size=(64, 64), roll_size=(2, 128, 50), input_dim=50, output_dim=20, instrument=0, start=100, end=150, device=device)


def _sd(npz, prefix="sd."):
    return {k[len(prefix):]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------ CPU
def test_checkpoint_fixture_loads_strict(golden_dir):
    g = np.load(os.path.join(golden_dir, "ckpt_epoch1.npz"))
    m = _mmgan("cpu")
    assert list(m.state_dict().keys()) == list(g["keys"])
    m.load_state_dict(_sd(g), strict=True)
    assert int(m.state_dict()["generator1.gen.0.1.num_batches_tracked"]) > 0
    assert sum(v.numel() for v in m.state_dict().values()) == 442653


def test_key_tables_of_all_shipped_checkpoints(golden_dir):
    t = np.load(os.path.join(golden_dir, "ckpt_keys.npz"))
    m = _mmgan("cpu")
    mine = m.state_dict()
    for rel in ("mmgan_64_64_epoch_1.pth", "MAE_loss/mmgan_64_64_epoch_35.pth", "V1_bad/mmgan_64_64_epoch_50.pth"):
        keys, shapes = list(t[rel + ".keys"]), list(t[rel + ".shapes"])
        if rel.startswith("V1_bad"):
            continue                                                            # older architecture (different widths): not loadable by the reference class either
        assert keys == list(mine.keys()), rel
        assert shapes == [",".join(map(str, v.shape)) for v in mine.values()], rel


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_shipped_checkpoint_files_load_into_mirror():
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    loaded = 0
    for rel in ("mmgan_64_64_epoch_1.pth", "MAE_loss/mmgan_64_64_epoch_35.pth"):
        sd = torch.load(os.path.join(REF, "MMGAN_MIDI_DES", "models", rel), map_location="cpu")
        m = _mmgan("cpu")
        m.load_state_dict(sd, strict=True)
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k
        loaded += 1
    for f in glob.glob(os.path.join(REF, "GAN_DES", "models", "gen_100_*.pt")):
        sd = torch.load(f, map_location="cpu")
        gen = SIMNN.Generator()
        gen.load_state_dict(sd, strict=True)
        loaded += 1
    assert loaded == 3


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_midi_fixture_matches_shipped_files(golden_dir):
    """the committed streams are what read_smf returns for the reference's own .mid files today; survey facts for simulation.mid"""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    g = np.load(os.path.join(golden_dir, "midi_streams.npz"))
    names = list(g["names"])
    assert len(names) == 30
    for i, rel in enumerate(names):
        ev = ds.read_smf(os.path.join(REF, rel))
        assert np.array_equal(ev.dt, g[f"s{i}.dt"]) and np.array_equal(ev.meta, g[f"s{i}.meta"]), rel
    i = names.index("MMGAN_MIDI_DES/adj_sim_outputs/midi/simulation.mid")
    kind = g[f"s{i}.meta"] & 0xFF
    assert len(kind) == 202 and int((kind == 1).sum()) == 106 and int((kind == 2).sum()) == 91


def test_midi_streams_oracle_numpy_vs_c(golden_dir):
    g = np.load(os.path.join(golden_dir, "midi_streams.npz"))
    for i in range(len(g["names"])):
        dt, meta = g[f"s{i}.dt"], g[f"s{i}.meta"]
        kind, pitch, vel = meta & 0xFF, (meta >> 8) & 0xFF, (meta >> 16) & 0xFF
        a, b = ro.raster_events(dt, kind, pitch, vel, 100, 0, 50)
        c, _ = ro.raster_batch_c(dt, meta, np.array([0, len(dt)]), 100, 0, 50)
        assert np.array_equal(c[0, 0], a) and np.array_equal(c[0, 1], b), i


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_checkpoint_eval_forward_vs_reference(golden_dir):
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    from gan_des_midi_music_gen_b200.gen_tc import GenTC
    g = np.load(os.path.join(golden_dir, "ckpt_epoch1.npz"))
    m = _mmgan(DEV)
    m.load_state_dict(_sd(g))
    seed, B = (int(v) for v in g["seed"])
    inp = mo.synth_inputs(B, seed=seed)
    m.generator1.eval(); m.generator2.eval()
    with torch.no_grad():
        g1 = m.generator1(inp["noise1"].to(DEV), torch.from_numpy(g["eval.inner"]).to(DEV))
        g2 = m.generator2(inp["noise2"].to(DEV), inp["beats"].to(DEV))
        logit = m.discriminator(inp["real"].to(DEV))
    assert np.abs(g1.cpu().numpy()[:, :, ::2, ::2] - g["eval.g1"]).max() <= 1e-5
    s = g["eval.g1.sum"]
    assert abs(g1.double().sum().item() - s[0]) <= 1e-5 * abs(s[0]) and abs((g1.double() ** 2).sum().item() - s[1]) <= 1e-5 * abs(s[1])
    assert np.abs(g2.cpu().numpy() - g["eval.g2"]).max() <= 1e-5
    want = g["disc.logit_real"].reshape(-1)
    assert np.abs(logit.cpu().numpy().reshape(-1) - want).max() <= 2e-5 * np.abs(want).max()
    # the bf16 tensor-core paths on the same checkpoint
    got = DiscTC(m.discriminator, max_batch=B).forward(inp["real"].to(DEV).to(torch.uint8))
    assert np.abs(got.cpu().numpy() - want).max() <= 5e-3 * np.abs(want).max()
    o1 = GenTC(m.generator1, B).forward(inp["noise1"].to(DEV), torch.from_numpy(g["eval.inner"]).to(DEV), training=False)
    o2 = GenTC(m.generator2, B).forward(inp["noise2"].to(DEV), inp["beats"].to(DEV), training=False)
    assert np.abs(o1.view(B, 1, 64, 64).cpu().numpy()[:, :, ::2, ::2] - g["eval.g1"]).max() <= 2e-2
    assert np.abs(o2.cpu().numpy() - g["eval.g2"]).max() <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("y", [0, 1])
def test_bf16_pass_under_survey_probe_conditions(golden_dir, y):
    """SURVEY 8d derived its bf16 bars from this discriminator at B = 256: the one-kernel pass meets every one of them."""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
    from gan_des_midi_music_gen_b200.disc_tc import DiscTC
    g = np.load(os.path.join(golden_dir, "disc_epoch1.npz"))
    B, W, seed = (int(v) for v in g["meta"])
    D = nt.DiscriminatorCNN(roll_size=(2, 128, W)).to(DEV)
    D.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")})
    x = torch.from_numpy(mo.synth_rolls(B, W, seed=seed)).to(DEV)
    tc = DiscTC(D, max_batch=B)
    loss = torch.zeros(1, device=DEV)
    logits = tc.pass_fused(x, float(y), loss)
    torch.cuda.synchronize()
    want = g[f"y{y}.logits"]
    assert np.abs(logits.cpu().numpy() - want).max() <= 5e-3 * np.abs(want).max()
    assert abs(loss.item() - g[f"y{y}.loss"].item()) <= 1e-3 * abs(g[f"y{y}.loss"].item())
    errs = {}
    for k, p in D.named_parameters():
        w = g[f"y{y}.grad." + k].astype(np.float64).ravel()
        errs[k] = np.linalg.norm(p.grad.cpu().double().numpy().ravel() - w) / np.linalg.norm(w)
    assert all(v <= 1e-2 for v in errs.values()), errs


@pytest.mark.gpu
def test_gandes_generator_checkpoint_eval_forward(golden_dir):
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    g = np.load(os.path.join(golden_dir, "ckpt_gandes_gen.npz"))
    gen = SIMNN.Generator().to(DEV)
    assert list(gen.state_dict().keys()) == list(g["keys"])
    gen.load_state_dict(_sd(g))
    gen.eval()
    with torch.no_grad():
        out = gen(torch.from_numpy(g["noise"]).to(DEV))
    assert out.shape == (4, 1, 20, 20)
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-5


@pytest.mark.gpu
def test_shipped_midi_streams_rasterise_bit_exact(golden_dir):
    """config 4 "plus the 30 shipped .mid": one batched device rasterisation of all 30 songs (both kernel paths, three call shapes)"""
    from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import datasets as ds
    g = np.load(os.path.join(golden_dir, "midi_streams.npz"))
    n = len(g["names"])
    dts, metas = [g[f"s{i}.dt"] for i in range(n)], [g[f"s{i}.meta"] for i in range(n)]
    off = np.concatenate([[0], np.cumsum([len(d) for d in dts])]).astype(np.int64)
    dt, meta = np.concatenate(dts), np.concatenate(metas)
    ev = (torch.from_numpy(dt).to(DEV), torch.from_numpy(meta.view(np.int32)).to(DEV), torch.from_numpy(off).to(DEV))
    for S, a, b in ((100, 0, 50), (300, 0, 300), (None, 3, 33)):
        want, _ = ro.raster_batch_c(dt, meta, off, S, a, b)
        for path in ("stream", "sort"):
            got = ds.rasterize_events(*ev, S, a, b, path=path)
            assert np.array_equal(got.cpu().numpy(), want), (S, a, b, path)
    assert want.any()

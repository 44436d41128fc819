"""GAN-DES leg of the bench alone (BASELINE config 2): tensor-core path and fp32 SIMT path, one JSON object.  `--once tc|simt` runs three
iterations of one path only (what the ncu launch list is taken from)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import __graft_entry__  # noqa: E402,F401  (registers the package alias)
from gan_des_midi_music_gen_b200 import benchmark as bm  # noqa: E402

if len(sys.argv) > 2 and sys.argv[1] == "--once" and sys.argv[2] == "trainer":
    # four iterations through GANDESTrainer (eager, capture, two replays): the last iteration's launches in an ncu list = the replayed graphs' kernel nodes
    import mmgan_oracle as mo
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    from gan_des_midi_music_gen_b200.gandes_trainer import GANDESTrainer
    B, dev = 30, "cuda"
    gshapes, dshapes = mo.gandes_shapes()
    gen, disc = SIMNN.Generator().to(dev).enable_tensor_cores(), SIMNN.Discriminator().to(dev).enable_tensor_cores()
    gen.load_state_dict(mo.synth_state(gshapes, seed=11)); disc.load_state_dict(mo.synth_state(dshapes, seed=12))
    g = torch.Generator().manual_seed(3)
    noise, real, fake = torch.randn(B, 100, 1, 1, generator=g).to(dev), torch.randn(B, 128, 216, generator=g).to(dev), torch.randn(B, 128, 216, generator=g).to(dev)
    tr = GANDESTrainer(gen, disc, lr=2e-5, betas=(0.5, 0.999))
    for _ in range(4):
        tr.generate(noise); tr.d_step(real, fake); tr.g_step(fake)
    torch.cuda.synchronize()
    print("ok")
elif len(sys.argv) > 2 and sys.argv[1] == "--once":
    import mmgan_oracle as mo
    from gan_des_midi_music_gen_b200.GAN_DES import SIMNN
    from gan_des_midi_music_gen_b200 import optim as fo
    tcores = sys.argv[2] == "tc"
    B, dev = 30, "cuda"
    gshapes, dshapes = mo.gandes_shapes()
    gen, disc = SIMNN.Generator().to(dev).enable_tensor_cores(tcores), SIMNN.Discriminator().to(dev).enable_tensor_cores(tcores)
    gen.load_state_dict(mo.synth_state(gshapes, seed=11)); disc.load_state_dict(mo.synth_state(dshapes, seed=12))
    crit = fo.BCEWithLogitsLoss()
    gen_opt, disc_opt = fo.FusedAdam(gen.parameters(), lr=2e-5, betas=(0.5, 0.999)), fo.FusedAdam(disc.parameters(), lr=2e-5, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(3)
    noise, real, fake = torch.randn(B, 100, 1, 1, generator=g).to(dev), torch.randn(B, 128, 216, generator=g).to(dev), torch.randn(B, 128, 216, generator=g).to(dev)
    t09, t01, t1 = torch.full((B,), 0.9, device=dev), torch.full((B,), 0.1, device=dev), torch.ones(B, device=dev)
    for _ in range(3):
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
    torch.cuda.synchronize()
    print("ok")
else:
    print(json.dumps(bm._gandes_leg("cuda")))

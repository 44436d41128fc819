"""World-size-2 checks of the data-parallel host logic on CPU (gloo): batch sharding and the single flat
D-gradient all-reduce of MMGANTrainer (the kernels themselves need a GPU and are covered by the -m gpu tests)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gan_des_midi_music_gen_b200.MMGAN_MIDI_DES import network_tests as nt
        from gan_des_midi_music_gen_b200.trainer import MMGANTrainer, shard_batch
        torch.manual_seed(0)
        m = nt.MultiModalGAN(z_dim=50, adj_size=(8, 8), roll_size=(2, 128, 50), input_dim=50, output_dim=16, device="cpu")
        tr = MMGANTrainer(m, precision="fp32")
        assert tr.world == world
        n = tr.flat_grad.numel()
        assert n == 21041 and all(p.grad.data_ptr() >= tr.flat_grad.data_ptr() for p in tr.d_params)
        # every p.grad is a view of the flat buffer: writing through .grad shows up in flat_grad
        tr.d_params[0].grad.fill_(float(rank + 1))
        tr.flat_grad[512:] = float(10 * (rank + 1))
        tr._allreduce_d_grads()
        ok = bool((tr.flat_grad[:512] == 3.0).all() and (tr.flat_grad[512:] == 30.0).all())
        # sharding: contiguous, equal, covers the global batch exactly once
        x = torch.arange(16 * 3).view(16, 3)
        mine = shard_batch(x, rank, world)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        ok = ok and torch.equal(torch.cat(gathered), x) and mine.shape[0] == 8
        # global-batch generator BatchNorm is a tensor-core-path feature: asking for it on the fp32 path must fail loudly, not silently fall back
        try:
            MMGANTrainer(m, precision="fp32", sync_bn=True)
            ok = False
        except ValueError as e:
            ok = ok and "sync_bn" in str(e)
        q.put((rank, ok, None))
    except Exception as e:       # pragma: no cover
        q.put((rank, False, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_grad_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def test_shard_batch_rejects_uneven():
    from gan_des_midi_music_gen_b200.trainer import shard_batch
    with pytest.raises(ValueError):
        shard_batch(torch.zeros(10, 2), 0, 4)

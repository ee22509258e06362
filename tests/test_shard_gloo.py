"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the shard partition, the final gather and the
max-over-ranks timing reduction; and the oracle confirms the batch semantics the sharding relies on (a batch of B is B
independent B=1 chains, so splitting a batch across ranks cannot change any image)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from osmosis_diffusion_code_b200.sharding import shard_indices, gather_images, max_over_ranks


def test_shard_indices_partition_exactly_and_match_the_loader():
    from osmosis_diffusion_code_b200.osmosis_utils.data import ShardedImageLoader

    class DS:
        def __init__(self, n): self.n = n
        def __len__(self): return self.n
    for n in (1, 2, 7, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
            for r in range(world):   # the same partition as the input loader (image i -> rank i mod world)
                assert ShardedImageLoader(DS(n), 4, rank=r, world=world).indices == parts[r]


def _worker(rank, world, port, n_images, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard_indices(n_images, rank, world)
    local = torch.tensor(idx, dtype=torch.float32)[:, None, None].expand(len(idx), 2, 3).contiguous() * 10.0
    full = gather_images(local, n_images)
    t = max_over_ranks(1.0 + rank)
    q.put((rank, full[:, 0, 0].tolist(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_max_over_two_gloo_ranks():
    world, n_images = 2, 5   # ragged: 3 + 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_images, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs: p.join(timeout=60)
    for rank, vals, t in res:
        assert vals == [0.0, 10.0, 20.0, 30.0, 40.0]
        assert t == 2.0


def test_oracle_batch_is_independent_images():
    """Per-image semantics: guidance losses / gradients of a batch equal those of its images taken one at a time."""
    from oracle import osmosis_oracle as orc
    from tests.helpers import load_yaml_cfg
    cfg = load_yaml_cfg("osmosis_sample_config.yaml", 6)
    tab, op, gs, phis, names = orc.specs_from_config(cfg, 2)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 4, 16, 16, generator=g).requires_grad_(True)
    y = torch.randn(2, 3, 16, 16, generator=g)
    total, norm, _ = orc.guidance_losses(op, x0, y, phis, gs.loss_weight, gs.weight_fn, gs.aux)
    (gfull,) = torch.autograd.grad(total.sum(), x0)
    for b in range(2):
        xb = x0[b:b + 1].detach().requires_grad_(True)
        tb, nb, _ = orc.guidance_losses(op, xb, y[b:b + 1], [p[b:b + 1] for p in phis], gs.loss_weight, gs.weight_fn, gs.aux)
        (gb,) = torch.autograd.grad(tb.sum(), xb)
        assert torch.equal(nb[0], norm[b]) and torch.allclose(gb[0], gfull[b], rtol=0, atol=0)

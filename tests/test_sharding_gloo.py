"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the single weight
broadcast and the batch sharding used by bench.py / the multi-GPU driver."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sudo_rm_rf_b200 as P
        from sudo_rm_rf_b200 import sharding
        from oracle import sudormrf_oracle as O
        kw = dict(out_channels=16, in_channels=32, num_blocks=2, upsampling_depth=3,
                  enc_kernel_size=21, enc_num_basis=24, num_sources=2)
        torch.manual_seed(100 + rank)                 # ranks start from DIFFERENT weights
        model = P.SuDORMRF(**kw)
        cfg = O.Config(variant="improved", **kw)
        ref_sd = O.make_state_dict(cfg, seed=3)
        if rank == 0:
            model.load_state_dict(ref_sd)
        nbytes = sharding.broadcast_parameters(model, src=0)
        same = all(torch.equal(v, ref_sd[k]) for k, v in model.state_dict().items())
        # shard a global batch of 7 mixtures, run the CPU oracle on the local shard, gather
        total = 7
        lo, hi = sharding.shard_bounds(total, world, rank)
        x = torch.randn(total, 1, 300, generator=torch.Generator().manual_seed(0))
        local = O.forward(cfg, model.state_dict(), x[lo:hi])
        full = sharding.gather_estimates(local, total)
        want = O.forward(cfg, ref_sd, x)
        q.put((rank, nbytes, same, (lo, hi), float((full - want).abs().max())))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_params = None
    for rank, nbytes, same, bounds, err in res:
        assert same, f"rank {rank}: parameters differ from rank 0 after the broadcast"
        assert err < 1e-5, err
        n_params = nbytes if n_params is None else n_params
        assert nbytes == n_params
    assert [r[3] for r in res] == [(0, 4), (4, 7)]


def test_shard_bounds_cover_batch():
    from sudo_rm_rf_b200.sharding import shard_bounds
    for total in (0, 1, 7, 32, 256, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)

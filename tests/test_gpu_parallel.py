"""GPU: the channel-block rotation and the sharded OT step (NCCL) against the single-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import optimaltextures_b200 as ob

    return ob


@pytest.mark.parametrize("n,c,c0,nc", [(4096, 512, 128, 64), (1000, 181, 60, 61), (16384, 256, 0, 128), (256, 64, 32, 32)])
def test_forward_block_equals_rows_of_the_full_rotation(ob, n, c, c0, nc):
    from optimaltextures_b200 import parallel

    g = torch.Generator().manual_seed(c0 + n)
    x = torch.relu(torch.randn(n, c, generator=g)).cuda()
    r = ob.random_rotation(c, "cuda", seed=2, counter=1)
    full = ob.rotate_forward(x, r)
    blk = parallel.cuda_ops().rotate_forward_block(x, r, c0, nc)
    assert torch.equal(blk, full[c0:c0 + nc])          # same K order per output element: bit-identical


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import optimaltextures_b200 as ob
    from optimaltextures_b200 import parallel

    g = torch.Generator().manual_seed(0)
    p = torch.relu(torch.randn(1, 64, 64, 256, generator=g)).cuda()
    s = torch.relu(1.3 * torch.randn(1, 48, 80, 256, generator=g) + 0.2).cuda()
    r = ob.random_rotation(256, "cuda", seed=5, counter=0)        # same (seed, counter) => same R on every rank
    out = parallel.optimal_transport_sharded(p, s, mode, r)
    single = ob.optimal_transport(p, s, mode, rotation=r)
    ok = bool(torch.equal(out, single))
    gathered = [None] * world
    dist.all_gather_object(gathered, ok)
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["cdf", "sort"])
def test_sharded_step_equals_single_gpu(mode):
    world = min(torch.cuda.device_count(), 2)
    import torch.multiprocessing as mp

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=300)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(got), f"sharded != single-GPU on ranks {got} (world={world})"

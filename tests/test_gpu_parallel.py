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


def _pixel_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import optimaltextures_b200 as ob
    from optimaltextures_b200 import parallel

    comm = parallel.Communicator()
    results = {}
    for (n_p, n_s, c) in [(4096, 6144, 256), (16384, 16384, 512), (65536, 32768, 64), (3000, 2000, 23)]:
        g = torch.Generator().manual_seed(n_p + c)
        p = torch.relu(torch.randn(n_p, c, generator=g)).cuda()
        s = torch.relu(1.3 * torch.randn(n_s, c, generator=g) + 0.2).cuda()
        r = ob.random_rotation(c, "cuda", seed=5, counter=0)        # same (seed, counter) => same R on every rank
        (p0, pk), (s0, sk) = parallel.row_slices(n_p, world)[rank], parallel.row_slices(n_s, world)[rank]
        for mode in ("cdf", "pca", "chol", "sym"):
            single = ob.optimal_transport(p.view(1, n_p, 1, c), s.view(1, n_s, 1, c), mode, rotation=r).view(n_p, c)
            mine = parallel.optimal_transport_pixel_sharded(p[p0:p0 + pk], s[s0:s0 + sk], mode, r, comm, n_p, n_s)
            want = single[p0:p0 + pk]
            if mode == "cdf":
                ok = bool(torch.equal(mine, want))                  # integer counts, exact extrema: bit-identical
            else:
                ok = float((mine - want).abs().max()) <= 5e-4 * max(1.0, float(want.abs().max()))
            results[f"{mode}@{n_p}x{c}"] = ok
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        q.put(gathered)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def _spawn(worker, world, *args):
    import torch.multiprocessing as mp

    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, *args, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    return got


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_pixel_sharded_step_equals_single_gpu():
    """optex_ot_step_sharded over NCCL: cdf bit-identical to one GPU, covariance modes within the fp32 tolerance."""
    world = min(torch.cuda.device_count(), 4)
    got = _spawn(_pixel_worker, world)
    bad = [(r, k) for r, res in enumerate(got) for k, ok in res.items() if not ok]
    assert not bad, f"pixel-sharded != single GPU: {bad} (world={world})"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
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

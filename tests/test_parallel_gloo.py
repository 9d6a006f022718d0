"""N > 1 path on CPU: world_size-2 gloo run of the channel-sharded OT step.  The device kernels are replaced by
the oracle's matchers through `parallel.Ops`, so what is tested is the sharding / all-gather logic itself:
the 2-rank result must EQUAL the single-process result bit for bit (no reduction order changes)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from optimaltextures_b200 import parallel


def test_channel_blocks_cover_and_align():
    for c, world in [(512, 8), (512, 3), (64, 4), (23, 2), (320, 8), (96, 8), (3, 4)]:
        blocks = parallel.channel_blocks(c, world)
        assert len(blocks) == world
        assert sum(k for _, k in blocks) == c
        pos = 0
        for s, k in blocks:
            assert s == pos and (s % 32 == 0 or k == 0)
            pos += k


def _oracle_ops():
    from oracle import ot_oracle, sort_oracle

    def fwd(x, r, c0, nc):
        return (x @ r.to(x))[:, c0:c0 + nc].T.contiguous()

    def match(t, s, mode):
        return sort_oracle.sort_match_channels(t, s)[0] if mode == "sort" else ot_oracle.cdf_match_channels(t, s)

    def inv(mt, r, content, w):
        out = mt.T @ r.to(mt).T
        if content is not None:
            out += w * (content - out)
        return out

    return parallel.Ops(fwd, match, inv)


def _worker(rank, world, port, c, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(0)
    p = torch.relu(torch.randn(1, 12, 10, c, generator=g))
    s = torch.relu(1.3 * torch.randn(1, 9, 14, c, generator=g) + 0.2)
    content = torch.randn(1, 12, 10, c, generator=g)
    from oracle import rotation

    r = torch.from_numpy(rotation.haar_rotation_qr(c, 7)).float()
    out = parallel.optimal_transport_sharded(p, s, mode, r, content=content, content_strength=0.1, ops=_oracle_ops())
    if rank == 0:
        q.put(out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("c,mode", [(64, "cdf"), (96, "sort"), (40, "cdf")])
def test_two_rank_sharded_step_equals_single_process(c, mode):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, c, mode, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process reference with the same ops
    ops = _oracle_ops()
    g = torch.Generator().manual_seed(0)
    p = torch.relu(torch.randn(1, 12, 10, c, generator=g))
    s = torch.relu(1.3 * torch.randn(1, 9, 14, c, generator=g) + 0.2)
    content = torch.randn(1, 12, 10, c, generator=g)
    from oracle import rotation

    r = torch.from_numpy(rotation.haar_rotation_qr(c, 7)).float()
    n = 120
    full = ops.match(ops.rotate_forward_block(p.reshape(n, c), r, 0, c), ops.rotate_forward_block(s.reshape(-1, c), r, 0, c), mode)
    ref = ops.rotate_inverse(full, r, content.reshape(n, c), 0.1).reshape(p.shape)
    np.testing.assert_array_equal(got, ref.numpy())


# ---------------------------------------------------------------- pixel sharding (what optex_ot_step_sharded does)
def test_row_slices_cover_and_align():
    for n, world in [(16384, 8), (1000, 3), (31, 2), (4096, 4), (65, 8)]:
        sl = parallel.row_slices(n, world)
        assert len(sl) == world and sum(k for _, k in sl) == n
        pos = 0
        for s, k in sl:
            assert s == pos and (s % 32 == 0 or k == 0)
            pos += k


def _pixel_worker(rank, world, port, c, q):
    """The pixel-sharded `cdf` step restated with torch ops + gloo collectives: local rows, all-reduce MIN/MAX of the
    range, all-reduce SUM of the histograms, identical tables on every rank (the C-ABI's algorithm, sharded.cu)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle import ot_oracle, rotation

    g = torch.Generator().manual_seed(1)
    n_p, n_s = 200, 170
    p = torch.relu(torch.randn(n_p, c, generator=g))
    s = torch.relu(1.3 * torch.randn(n_s, c, generator=g) + 0.2)
    r = torch.from_numpy(rotation.haar_rotation_qr(c, 3)).float()
    (p0, pk), (s0, sk) = parallel.row_slices(n_p, world)[rank], parallel.row_slices(n_s, world)[rank]
    rp = (p[p0:p0 + pk] @ r).T.contiguous()          # [c, n_local]: rotations are row-local
    rs = (s[s0:s0 + sk] @ r).T.contiguous()
    out = torch.empty_like(rp)
    for ch in range(c):
        lo = torch.min(rp[ch].min(), rs[ch].min()).reshape(1)
        hi = torch.max(rp[ch].max(), rs[ch].max()).reshape(1)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        th, sh = torch.histc(rp[ch], 256, lo[0], hi[0]), torch.histc(rs[ch], 256, lo[0], hi[0])
        dist.all_reduce(th)
        dist.all_reduce(sh)                              # integer counts: exact in fp32 below 2^24
        edges = torch.linspace(lo[0], hi[0], 257)[1:]
        tc, sc = th.cumsum(0), sh.cumsum(0)
        remap = ot_oracle.interp_backward(tc / tc[-1], sc / sc[-1], edges)
        out[ch] = ot_oracle.interp_backward(rp[ch], edges, remap)
    mine = out.T @ r.T
    parts = [None] * world
    dist.all_gather_object(parts, mine.numpy())
    if rank == 0:
        q.put(np.concatenate(parts, 0))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("c", [16, 40])
def test_two_rank_pixel_sharded_cdf_equals_single_process(c):
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pixel_worker, args=(r, 2, port, c, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    from oracle import ot_oracle, rotation

    g = torch.Generator().manual_seed(1)
    p = torch.relu(torch.randn(200, c, generator=g))
    s = torch.relu(1.3 * torch.randn(170, c, generator=g) + 0.2)
    r = torch.from_numpy(rotation.haar_rotation_qr(c, 3)).float()
    # row-local rotation (a row's product does not depend on the other rows), then the reference's matcher
    rp = torch.cat([(p[a:a + k] @ r) for a, k in parallel.row_slices(200, 2)]).T.contiguous()
    rs = torch.cat([(s[a:a + k] @ r) for a, k in parallel.row_slices(170, 2)]).T.contiguous()
    m = ot_oracle.cdf_match_channels(rp, rs)
    ref = torch.cat([(m.T[a:a + k] @ r.T) for a, k in parallel.row_slices(200, 2)])
    np.testing.assert_array_equal(got, ref.numpy())

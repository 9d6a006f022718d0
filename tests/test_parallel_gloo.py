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

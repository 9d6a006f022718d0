"""Multi-GPU form of the OT step: rotated-channel sharding + one all-gather.

The reference has no distributed code at all.  After the rotation the C channels are independent 1-D
problems (histmatch.py:51 loops over them), so for the per-channel modes (`cdf`, `sort`):

    rank g                                                   (one process per GPU, torch.distributed / NCCL)
      Xt_g  = (P @ R[:, blk_g])^T      [C/G, N_p]   forward rotation of its channel block only (1/G of the FLOPs)
      St_g  = (S @ R[:, blk_g])^T      [C/G, N_s]
      Mt_g  = match(Xt_g, St_g)        [C/G, N_p]   cdf_match / sort_match on its channels
      Mt    = all_gather(Mt_g)         [C,   N_p]   channel-major => each rank's block is one contiguous slab
      out   = Mt^T @ R^T               [N_p, C]     full inverse rotation, redundantly on every rank

P, S and R are replicated (R: same seed/counter on every rank), and `out` is bit-identical on every rank and
identical to the 1-GPU result: no reduction is involved, so no summation order changes.  It pays only where one
GPU saturates (conv1_1 / conv2_1 at >= 1024^2, everything at 2048^2): the all-gather moves 4*N_p*C*(G-1)/G bytes.

The covariance modes (chol / pca / sym) run replicated - their per-step work after the algebraic folding
(cov_match.cu) is two N x C x C GEMMs, which a pixel-sharded Gram + all-reduce would split (next round).

PIXEL sharding (the `cdf` and covariance modes; `Communicator`, `optimal_transport_pixel_sharded`, C-ABI
`optex_ot_step_sharded`): every rank keeps a slice of the ROWS of P and S.  Rotations are row-local; `cdf` needs only
the per-channel range and histograms of the whole block (two tiny all-reduces: 2C words MIN, 2*256*C counts SUM) and is
then bit-identical to one GPU; the covariance modes all-reduce column sums and the centred Gram.  No feature data
crosses NVLink, and in a loop the block simply stays sharded between iterations.

`ops` makes the device kernels pluggable so the sharding logic itself is tested on CPU with gloo (tests/
test_parallel_gloo.py drives it with the oracle's matchers); the default ops are the CUDA kernels.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def channel_blocks(c: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous (start, count) channel blocks, one per rank.  Starts are multiples of `align` (the tensor-core
    path wants 32-wide blocks); trailing ranks may own nothing when c is small."""
    units = (c + align - 1) // align
    per, extra = divmod(units, world)
    blocks, start = [], 0
    for r in range(world):
        n_units = per + (1 if r < extra else 0)
        count = max(0, min(c, start + n_units * align) - start)
        blocks.append((start, count))
        start += count
    assert start == c
    return blocks


@dataclass
class Ops:
    """The three device operations of the sharded step."""
    rotate_forward_block: Callable[[Tensor, Tensor, int, int], Tensor]   # (x[n,c], R, c0, nc) -> [nc, n]
    match: Callable[[Tensor, Tensor, str], Tensor]                        # (t[nc,n], s[nc,m], mode) -> [nc, n]
    rotate_inverse: Callable[[Tensor, Tensor, Optional[Tensor], float], Tensor]  # (mt[c,n], R, content, w) -> [n,c]
    prepare: Optional[Callable[[Tensor], object]] = None   # context manager factory: R shared by the three GEMMs


def cuda_ops() -> Ops:
    from . import _lib, histmatch
    from ._runtime import call, f32c, ptr, stream_ptr
    from .optex import prepared_rotation, rotate_inverse

    def fwd(x: Tensor, r: Tensor, c0: int, nc: int) -> Tensor:
        xc, rc = f32c(x), f32c(r)
        c = xc.shape[-1]
        n = xc.numel() // c
        out = torch.empty(nc, n, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("optex_rotate_forward_block", ptr(xc), ptr(rc), ptr(out), n, c, c0, nc, stream_ptr(x.device))
        return out

    def match(t: Tensor, s: Tensor, mode: str) -> Tensor:
        return histmatch.sort_match(t, s) if mode == "sort" else histmatch.cdf_match(t, s)

    return Ops(fwd, match, rotate_inverse, prepared_rotation)


def optimal_transport_sharded(pastiche_feature: Tensor, style_feature: Tensor, hist_mode: str, rotation: Tensor,
                              content: Optional[Tensor] = None, content_strength: float = 0.0,
                              group=None, ops: Optional[Ops] = None) -> Tensor:
    """optex.py:167-177 over the ranks of `group` (see module docstring).  Every rank passes the same replicated
    tensors and receives the same full result."""
    if hist_mode not in ("cdf", "sort"):
        raise ValueError("channel sharding applies to the per-channel modes cdf / sort; run chol / pca / sym "
                         "through optimal_transport (replicated)")
    ops = ops or cuda_ops()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    c = pastiche_feature.shape[-1]
    n = pastiche_feature.numel() // c
    blocks = channel_blocks(c, world)
    c0, nc = blocks[rank]
    with (ops.prepare(rotation) if ops.prepare is not None else contextlib.nullcontext(rotation)) as rotation:
        mt = torch.empty(c, n, dtype=torch.float32, device=pastiche_feature.device)
        if nc > 0:
            xt = ops.rotate_forward_block(pastiche_feature.reshape(n, c), rotation, c0, nc)
            st = ops.rotate_forward_block(style_feature.reshape(-1, c), rotation, c0, nc)
            mine = ops.match(xt, st, hist_mode)
        else:
            mine = mt[:0]
        if len({b[1] for b in blocks}) == 1:
            dist.all_gather_into_tensor(mt, mine.contiguous(), group=group)           # equal slabs: one collective
        else:
            slabs = [mt[s:s + k] for s, k in blocks]
            pad = max(k for _, k in blocks)
            if all(k == pad for _, k in blocks):
                dist.all_gather(slabs, mine.contiguous(), group=group)
            else:                                                                     # ragged tail: pad to equal slabs
                buf = torch.zeros(world, pad, n, dtype=torch.float32, device=mt.device)
                send = torch.zeros(pad, n, dtype=torch.float32, device=mt.device)
                send[:nc] = mine
                dist.all_gather_into_tensor(buf.view(world * pad, n), send, group=group)
                for r, (s, k) in enumerate(blocks):
                    mt[s:s + k] = buf[r, :k]
        out = ops.rotate_inverse(mt, rotation, content.reshape(n, c) if content is not None else None, content_strength)
    return out.reshape(pastiche_feature.shape)


# ----------------------------------------------------------------------------------------- pixel sharding (C-ABI)
class Communicator:
    """NCCL communicator owned by liboptex_b200 (`optex_comm_init`).  The unique id travels through the already
    initialised torch.distributed group (any backend); one instance per process / GPU."""

    def __init__(self, group=None):
        import ctypes as C

        from . import _lib

        lib = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = int(lib.optex_comm_unique_id_bytes())
        buf = (C.c_ubyte * nbytes)()
        if self.rank == 0:
            _lib.check(lib.optex_comm_unique_id(buf, nbytes))
        box = [bytes(buf) if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        raw = (C.c_ubyte * nbytes).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        _lib.check(lib.optex_comm_init(raw, self.rank, self.world, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        from . import _lib

        if self._h:
            _lib.lib().optex_comm_destroy(self._h)
            self._h = None

    def all_gather_rows(self, local: Tensor) -> Tensor:
        """[n_local, c] of every rank (equal n_local) -> [world * n_local, c], rank order."""
        from ._runtime import call, ptr, stream_ptr

        out = torch.empty(self.world * local.shape[0], *local.shape[1:], dtype=torch.float32, device=local.device)
        call("optex_comm_allgather_f32", self._h, ptr(local), ptr(out), local.numel(), stream_ptr(local.device))
        return out


def row_slices(n: int, world: int, align: int = 32) -> List[Tuple[int, int]]:
    """Contiguous (start, count) row slices, one per rank, starts multiples of `align` (the tensor-core inverse
    rotation wants 32-row blocks); the last rank takes the remainder."""
    units = (n + align - 1) // align
    per, extra = divmod(units, world)
    out, start = [], 0
    for r in range(world):
        k = max(0, min(n, start + (per + (1 if r < extra else 0)) * align) - start)
        out.append((start, k))
        start += k
    assert start == n
    return out


def optimal_transport_pixel_sharded(p_local: Tensor, s_local: Tensor, hist_mode: str, rotation: Optional[Tensor],
                                    comm: Communicator, n_p_total: int, n_s_total: int,
                                    content_local: Optional[Tensor] = None, content_strength: float = 0.0,
                                    eps: float = 1.0, out: Optional[Tensor] = None) -> Tensor:
    """optex.py:167-177 on this rank's rows [n_local, c] of the pastiche / style blocks (`optex_ot_step_sharded`)."""
    from . import _lib
    from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace

    dev = require_cuda(p_local, s_local, rotation, content_local)
    p, s = f32c(p_local), f32c(s_local)
    c = p.shape[-1]
    n_p, n_s = p.numel() // c, s.numel() // c
    m = _lib.mode_id(hist_mode)
    if out is None:
        out = torch.empty_like(p)
    wsb = workspace(dev, _lib.lib().optex_ot_workspace_bytes(n_p, n_s, c, m))
    r = f32c(rotation) if rotation is not None else None
    ct = f32c(content_local) if content_local is not None else None
    with torch.cuda.device(dev):
        call("optex_ot_step_sharded", comm.handle, ptr(p), ptr(s), ptr(r), ptr(out), n_p, n_s, int(n_p_total),
             int(n_s_total), c, m, float(eps), ptr(ct), float(content_strength), ptr(wsb), wsb.numel(), stream_ptr(dev))
    return out

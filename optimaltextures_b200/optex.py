"""Drop-in for the hot-path functions of the reference's ``optex`` module (/root/reference/optex.py).

    optimal_transport(pastiche_feature, style_feature, hist_mode)   optex.py:167-177
    random_rotation(N, device, impl)                                optex.py:142-164
    ot_loop(...)                                                    optex.py:112-117 (the inner loop)
    fit_pca(tensor)                                                 optex.py:180-190
    pca_project(x, eigvecs, transpose)                              optex.py:110, :120 (the `@ style_eigvs[l]` pair)

Additive keyword arguments (the reference draws its rotation internally from numpy's global RNG,
which makes value parity impossible): ``rotation=`` injects the matrix, ``content=`` /
``content_strength=`` fuse the in-loop content blend (optex.py:115-117) into the inverse rotation.
``hist_mode="sort"`` selects the exact per-channel 1-D OT (not in the reference).

`install(module)` rebinds a loaded reference ``optex`` module to these functions.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from . import histmatch as _histmatch
from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace

_seed = 0
_counter = itertools.count()


def manual_seed(seed: int) -> None:
    """Seed the on-device rotation stream (Philox key).  The reference's --seed does NOT control its
    rotations (numpy's global RNG, optex.py:149 vs :253); here it does."""
    global _seed, _counter
    _seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    _counter = itertools.count()
    _rotation_pool.clear()


# optimal_transport() draws its own rotation like the reference does (optex.py:168).  One draw is latency-bound
# (117 us at N = 512), a batch costs 29 us per matrix, so the implicit draws come out of a small per-(device, N)
# pool filled 16 at a time; the stream stays a pure function of manual_seed() and the order of the calls.
_POOL = 16
_rotation_pool = {}


def _pooled_rotation(N: int, dev) -> Tensor:
    stream = torch.cuda.current_stream(dev).cuda_stream
    key = (dev.index, N)
    ent = _rotation_pool.get(key)
    if ent is None or ent[1] >= _POOL or ent[2] != stream or ent[3] != _seed:
        first = next(_counter)
        for _ in range(_POOL - 1):
            next(_counter)
        ent = [random_rotations(N, _POOL, dev, seed=_seed, first_counter=first), 0, stream, _seed]
        _rotation_pool[key] = ent
    r = ent[0][ent[1]]
    ent[1] += 1
    return r


def set_gemm_mode(mode: str) -> None:
    """Rotation-GEMM arithmetic: "auto" | "fp32" (SIMT FFMA) | "tf32x3" (tcgen05, fp32-grade) | "tf32"."""
    call("optex_set_gemm_mode", _lib.GEMM_MODES[mode])


def set_rotation_precision(precision: str) -> str:
    """Arithmetic of the on-device Householder construction: "fp64" (default, like the reference's scipy branch
    before `.to(pastiche_feature)`) or "fp32" (like its impl="torch" branch, optex.py:150-164).  Returns the
    previous setting."""
    if precision not in ("fp64", "fp32"):
        raise ValueError(f"rotation precision must be 'fp64' or 'fp32', got {precision!r}")
    prev = _lib.lib().optex_set_rotation_precision(1 if precision == "fp64" else 0)
    return "fp64" if prev else "fp32"


def random_rotation(N: int, device="cuda", impl: str = "device", seed: Optional[int] = None,
                    counter: Optional[int] = None, gauss: Optional[Tensor] = None) -> Tensor:
    """reference: optex.py:142-164.  Haar SO(N) matrix, fp32 [N, N] on `device`.

    impl="device" (default): Philox normals + Householder construction in fp64 on the GPU (the
    reference's impl="torch" algorithm).  impl="scipy": the reference's live branch, on the host,
    for callers that want scipy's stream (slow: 35 ms at N=512)."""
    dev = torch.device(device)
    if impl == "scipy":
        from scipy.stats import special_ortho_group

        return torch.tensor(special_ortho_group.rvs(N), device=dev).float()
    if dev.type != "cuda":
        raise RuntimeError("random_rotation(impl='device') generates on a B200; pass a cuda device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(N, N, dtype=torch.float32, device=dev)
    nbytes = _lib.lib().optex_rotation_workspace_bytes(N)
    wsb = workspace(dev, nbytes)
    g = None
    if gauss is not None:
        g = gauss.to(device=dev, dtype=torch.float64).contiguous()
        if tuple(g.shape) != (max(N - 1, 0), N):
            raise ValueError(f"gauss must be [N-1, N], got {tuple(g.shape)}")
    with torch.cuda.device(dev):
        call("optex_random_rotation", ptr(out), N, _seed if seed is None else int(seed),
             next(_counter) if counter is None else int(counter), ptr(g), ptr(wsb), wsb.numel(), stream_ptr(dev))
    return out


class prepared_rotation:
    """Context manager: split `rotation` into its tf32 hi / lo halves once (`optex_rotation_prepare`) for the
    `rotate_forward` / `rotate_inverse` calls inside the block that use the same tensor - what `optimal_transport`
    does internally for the three `@` of optex.py:170,171,175.  No-op for shapes the tensor-core path does not take."""

    def __init__(self, rotation: Tensor):
        self.r = f32c(rotation)
        self.ws = None

    def __enter__(self):
        r = self.r
        c = r.shape[-1]
        if r.is_cuda and r.dim() == 2 and c % 4 == 0:
            lib = _lib.lib()
            self.ws = torch.empty(lib.optex_rotation_prepare_workspace_bytes(c), dtype=torch.uint8, device=r.device)
            with torch.cuda.device(r.device):
                call("optex_rotation_prepare", ptr(r), c, ptr(self.ws), self.ws.numel(), stream_ptr(r.device))
        return r

    def __exit__(self, *exc):
        if self.ws is not None:
            call("optex_rotation_prepare", None, 0, None, 0, None)
            self.ws = None
        return False


def _nhwc_dims(x: Tensor):
    if x.dim() != 4:
        raise ValueError(f"expected an NHWC feature tensor [b,h,w,c], got shape {tuple(x.shape)}")
    b, h, w, c = x.shape
    return b, h * w, c


def optimal_transport(pastiche_feature: Tensor, style_feature: Tensor, hist_mode: str, *,
                      rotation: Optional[Tensor] = None, content: Optional[Tensor] = None,
                      content_strength: float = 0.0, eps: float = 1.0) -> Tensor:
    """reference: optex.py:167-177.  One sliced-OT step; returns a NEW contiguous [b,h,w,c] tensor."""
    dev = require_cuda(pastiche_feature, style_feature, rotation, content)
    b, hw, c = _nhwc_dims(pastiche_feature)
    bs, hws, cs = _nhwc_dims(style_feature)
    if cs != c:
        raise ValueError(f"channel mismatch: pastiche {c} vs style {cs}")
    p, s = f32c(pastiche_feature), f32c(style_feature)
    if rotation is None:
        rotation = _pooled_rotation(c, dev)
    r = f32c(rotation)  # `.to(pastiche_feature)` in the reference: float64 scipy matrix -> fp32
    if tuple(r.shape) != (c, c):
        raise ValueError(f"rotation must be [{c}, {c}], got {tuple(r.shape)}")
    ct = None
    if content is not None:
        if content.shape != pastiche_feature.shape:
            raise ValueError("content must have the pastiche's shape")
        ct = f32c(content)
    out = torch.empty_like(p)
    m = _lib.mode_id(hist_mode)
    nbytes = _lib.lib().optex_ot_workspace_bytes(b * hw, bs * hws, c, m)
    wsb = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        call("optex_ot_step", ptr(p), ptr(s), ptr(r), ptr(out), b, hw, bs, hws, c, m, float(eps), ptr(ct),
             float(content_strength), ptr(wsb), wsb.numel(), stream_ptr(dev))
    return out.to(pastiche_feature.dtype)


def ot_loop(pastiche_feature: Tensor, style_feature: Tensor, hist_mode: str, iters: int, *,
            rotations: Optional[Tensor] = None, content: Optional[Tensor] = None, content_strength: float = 0.0,
            eps: float = 1.0, seed: Optional[int] = None, first_counter: Optional[int] = None) -> Tensor:
    """reference: optex.py:112-117 - `iters` OT steps (+ content blend) as one native call.

    rotations: optional [iters, c, c]; otherwise generated on the device, `iters` at a time."""
    dev = require_cuda(pastiche_feature, style_feature, rotations, content)
    b, hw, c = _nhwc_dims(pastiche_feature)
    bs, hws, _ = _nhwc_dims(style_feature)
    feat = f32c(pastiche_feature).clone()
    s = f32c(style_feature)
    r = None
    if rotations is not None:
        r = f32c(rotations)
        if tuple(r.shape) != (iters, c, c):
            raise ValueError(f"rotations must be [{iters}, {c}, {c}]")
    ct = f32c(content) if content is not None else None
    m = _lib.mode_id(hist_mode)
    nbytes = _lib.lib().optex_ot_loop_workspace_bytes(b * hw, bs * hws, c, m)
    wsb = workspace(dev, nbytes)
    if first_counter is None:
        first_counter = next(_counter)
        for _ in range(max(iters - 1, 0)):
            next(_counter)
    with torch.cuda.device(dev):
        call("optex_ot_loop", ptr(feat), ptr(s), ptr(r), int(iters), _seed if seed is None else int(seed),
             int(first_counter), b, hw, bs, hws, c, m, float(eps), ptr(ct), float(content_strength), ptr(wsb),
             wsb.numel(), stream_ptr(dev))
    return feat


def random_rotations(N: int, count: int, device="cuda", seed: Optional[int] = None,
                     first_counter: int = 0) -> Tensor:
    """`count` Haar SO(N) matrices [count, N, N] in one batched launch pair (see `random_rotation`)."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("random_rotations generates on a B200; pass a cuda device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.empty(count, N, N, dtype=torch.float32, device=dev)
    wsb = workspace(dev, _lib.lib().optex_rotations_workspace_bytes(N, count))
    with torch.cuda.device(dev):
        call("optex_random_rotations", ptr(out), N, count, _seed if seed is None else int(seed), int(first_counter),
             None, ptr(wsb), wsb.numel(), stream_ptr(dev))
    return out


def set_host_style(style, stream=None) -> None:
    """Upload the (constant) style block once and keep it resident for `optimal_transport_host(style=None)`:
    optex.py:112-113 hands the same style_features[l] to every iteration of a layer.  style=None releases it."""
    st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
    if style is None:
        call("optex_ot_host_set_style", None, 0, 0, 0, st)
        return
    if style.is_cuda or style.dtype != torch.float32 or not style.is_contiguous():
        raise ValueError("set_host_style takes a contiguous fp32 CPU tensor")
    bs, hws, c = _nhwc_dims(style)
    call("optex_ot_host_set_style", ptr(style), bs, hws, c, st)


def optimal_transport_host(pastiche, style, rotation, hist_mode: str, content=None, content_strength: float = 0.0,
                           eps: float = 1.0, out=None, seed: Optional[int] = None, counter: int = 0,
                           slot: Optional[int] = None, stream=None, style_shape=None):
    """The same step through the HOST-buffer entry point (`optex_ot_step_host`): CPU tensors in,
    CPU tensor out, H2D/D2H inside the call.  This is what a non-torch caller of the C-ABI gets.
    rotation=None draws the rotation on the device from (seed, counter), like the reference draws its own.
    style=None reuses the block made resident by `set_host_style` (pass its `style_shape`).
    slot=0/1/2 + stream=: the asynchronous, multi-buffered form (`optex_ot_step_host_async`) for pipelining
    independent steps - step i+1 uploads while step i computes and step i-1 downloads."""
    for t in (pastiche, style, rotation, content):
        if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise ValueError("optimal_transport_host takes contiguous fp32 CPU tensors")
    b, hw, c = _nhwc_dims(pastiche)
    if style is not None:
        bs, hws, _ = _nhwc_dims(style)
    else:
        if style_shape is None or len(style_shape) != 4:
            raise ValueError("style=None (resident style) needs style_shape=[b,h,w,c]")
        bs, hws = int(style_shape[0]), int(style_shape[1]) * int(style_shape[2])
    if out is None:
        out = torch.empty_like(pastiche)
    st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
    if slot is None:
        call("optex_ot_step_host", ptr(pastiche), ptr(style), ptr(rotation), ptr(out), b, hw, bs, hws, c,
             _lib.mode_id(hist_mode), float(eps), ptr(content), float(content_strength),
             _seed if seed is None else int(seed), int(counter), st)
    else:  # pipelined: no synchronisation - the caller syncs `stream` before reading `out` (pinned buffers!)
        call("optex_ot_step_host_async", ptr(pastiche), ptr(style), ptr(rotation), ptr(out), b, hw, bs, hws, c,
             _lib.mode_id(hist_mode), float(eps), ptr(content), float(content_strength),
             _seed if seed is None else int(seed), int(counter), int(slot), st)
    return out


def rotate_forward(x: Tensor, rotation: Tensor) -> Tensor:
    """(x @ R)^T as channel-major [c, n] (optex.py:170-171 + the permute of histmatch.py:6-8)."""
    dev = require_cuda(x, rotation)
    xc, r = f32c(x), f32c(rotation)
    c = xc.shape[-1]
    n = xc.numel() // c
    out = torch.empty(c, n, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        call("optex_rotate_forward", ptr(xc), ptr(r), ptr(out), n, c, stream_ptr(dev))
    return out


def rotate_inverse(mt: Tensor, rotation: Tensor, content: Optional[Tensor] = None,
                   content_strength: float = 0.0) -> Tensor:
    """mt [c, n] channel-major -> (mt^T @ R^T) [n, c] (optex.py:175), optional blend (optex.py:117)."""
    dev = require_cuda(mt, rotation, content)
    m, r = f32c(mt), f32c(rotation)
    c, n = m.shape
    out = torch.empty(n, c, dtype=torch.float32, device=dev)
    ct = f32c(content) if content is not None else None
    with torch.cuda.device(dev):
        call("optex_rotate_inverse", ptr(m), ptr(r), ptr(out), n, c, ptr(ct), float(content_strength),
             stream_ptr(dev))
    return out


def pca_project(x: Tensor, eigvecs: Tensor, transpose: bool = False) -> Tensor:
    """`x @ eigvecs` (optex.py:110, :188) or, with transpose=True, `x @ eigvecs.T` (optex.py:120) on the B200 GEMMs.
    x [..., c] (or [..., k] for the transpose), eigvecs [c, k]."""
    dev = require_cuda(x, eigvecs)
    v = f32c(eigvecs)
    if v.dim() != 2:
        raise ValueError(f"eigvecs must be [c, k], got {tuple(v.shape)}")
    c, k = v.shape
    xin = f32c(x)
    cin, cout = (k, c) if transpose else (c, k)
    if xin.shape[-1] != cin:
        raise ValueError(f"last dimension {xin.shape[-1]} does not match eigvecs {tuple(v.shape)}"
                         f"{' (transposed)' if transpose else ''}")
    n = xin.numel() // max(cin, 1)
    out = torch.empty(*xin.shape[:-1], cout, dtype=torch.float32, device=dev)
    if n == 0 or k == 0 or c == 0:
        return out.zero_().to(x.dtype)      # an empty basis projects to nothing / back to zeros, like torch's `@`
    with torch.cuda.device(dev):
        call("optex_pca_project", ptr(xin), ptr(v), ptr(out), n, c, k, 1 if transpose else 0, stream_ptr(dev))
    return out.to(x.dtype)


def fit_pca(tensor: Tensor, *, round_k_to: int = 1, return_sigma: bool = False, basis: Optional[Tensor] = None,
            warm: bool = False):
    """reference: optex.py:180-190.  Returns (features, eigvecs) = (tensor @ V[:, :k], V[:, :k]).

    The reference's `torch.svd(tensor - tensor.mean())` is computed on the device as the FP64 eigendecomposition of
    the c x c Gram matrix (csrc/pca.cu).  k follows the reference's 90 % rule (optex.py:184); reading it is the one
    host synchronisation of the call (the reference has the same one: it slices by a tensor index).
    An SVD leaves the sign of every singular vector open: here the largest-magnitude component of each column of
    `eigvecs` is positive.

    round_k_to (additive): round k UP to a multiple (e.g. 32, which keeps the PCA'd channel count on the
    tensor-core path of the OT step); the extra components only add explained variance.

    basis / warm (additive): `basis` is a [c, c] float64 CUDA tensor that receives the eigensolver's orthogonal
    matrix; with warm=True the solve STARTS from it (optex_fit_pca_warm) - same result, fewer Jacobi sweeps when
    the block resembles the one the basis came from (the same style image at the previous pass's size)."""
    dev = require_cuda(tensor)
    c = tensor.shape[-1]
    x = f32c(tensor).reshape(-1, c)
    n = x.shape[0]
    lib = _lib.lib()
    vecs = torch.empty(c, c, dtype=torch.float32, device=dev)
    sigma = torch.empty(c, dtype=torch.float32, device=dev)
    k_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = workspace(dev, lib.optex_fit_pca_workspace_bytes(n, c))
    _check_basis(basis, c, dev, warm)
    with torch.cuda.device(dev):
        call("optex_fit_pca_warm", ptr(x), n, c, ptr(vecs), ptr(sigma), ptr(k_dev),
             None if basis is None else ptr(basis), 1 if warm else 0, None, ptr(wsb), wsb.numel(), stream_ptr(dev))
    k = int(k_dev.item())
    if round_k_to > 1:
        k = min(c, -(-k // round_k_to) * round_k_to)
    eigvecs = vecs[:, :k].contiguous()
    features = pca_project(tensor, eigvecs)
    if return_sigma:
        return features, eigvecs, sigma
    return features, eigvecs


def _check_basis(basis, c, dev, warm):
    if basis is None:
        if warm:
            raise ValueError("warm=True needs basis= (the float64 [c, c] tensor an earlier fit_pca call filled)")
        return
    if basis.dtype != torch.float64 or tuple(basis.shape) != (c, c) or not basis.is_contiguous() or basis.device != dev:
        raise ValueError(f"basis must be a contiguous float64 [{c}, {c}] tensor on {dev}, got {basis.dtype} "
                         f"{tuple(basis.shape)} on {basis.device}")


_pca_streams = {}


def fit_pca_many(tensors, *, round_k_to: int = 1, bases=None, warm=None, sweeps_out: Optional[list] = None):
    """`fit_pca` (optex.py:180-190) of several independent feature blocks - the five VGG layers of one pass
    (optex.py:62-67) - with the eigensolvers running CONCURRENTLY on side streams: one solve is bound by the latency
    of its grid-wide barriers, not by throughput, and the five cooperative grids (<= 32 CTAs each for the cold solve)
    fit on the device together.  Returns [(features, eigvecs), ...] like five `fit_pca` calls; one host
    synchronisation for all the k's.  bases / warm: per-tensor `basis=` / `warm=` of `fit_pca`; sweeps_out: a list
    that receives the Jacobi sweep count of every solve."""
    return fit_pca_many_finish(fit_pca_many_launch(tensors, bases=bases, warm=warm), round_k_to=round_k_to,
                               sweeps_out=sweeps_out)


def fit_pca_many_launch(tensors, *, bases=None, warm=None):
    """First half of `fit_pca_many`: everything up to (not including) the host read of the k's - the solves are
    enqueued (on side streams, joined back into the current stream) and the call returns without synchronising, so a
    caller can put other work between the launch and `fit_pca_many_finish` (OptimalTexture: the whole previous pass)."""
    if not tensors:
        return None
    dev = require_cuda(*tensors)
    bases = list(bases) if bases is not None else [None] * len(tensors)
    warm = list(warm) if warm is not None else [False] * len(tensors)
    if len(bases) != len(tensors) or len(warm) != len(tensors):
        raise ValueError("bases= and warm= need one entry per tensor")
    lib = _lib.lib()
    cur = torch.cuda.current_stream(dev)
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), cur.cuda_stream)
    streams = _pca_streams.setdefault(key, [])
    while len(streams) < len(tensors):
        streams.append(torch.cuda.Stream(device=dev))
    jobs = []
    # every buffer is allocated on the caller's stream; the side streams only borrow them between the two waits
    for t in tensors:
        c = t.shape[-1]
        x = f32c(t).reshape(-1, c)
        jobs.append((x, c, torch.empty(c, c, dtype=torch.float32, device=dev),
                     torch.empty(c, dtype=torch.float32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev),
                     torch.empty(max(int(lib.optex_fit_pca_workspace_bytes(x.shape[0], c)), 256), dtype=torch.uint8,
                                 device=dev)))
    for (x, c, *_), b, w in zip(jobs, bases, warm):
        _check_basis(b, c, dev, w)
    with torch.cuda.device(dev):
        # the producers of `tensors` were launched with programmatic serialisation: an ordinary fence kernel in front
        # of the fork (and of every join) keeps the cross-stream events behind their complete drain
        call("optex_fence", C.c_void_p(cur.cuda_stream))
        for st, (x, c, vecs, sigma, k_dev, wsb), b, w in zip(streams, jobs, bases, warm):
            st.wait_stream(cur)
            call("optex_fit_pca_warm", ptr(x), x.shape[0], c, ptr(vecs), ptr(sigma), ptr(k_dev),
                 None if b is None else ptr(b), 1 if w else 0, C.c_void_p(k_dev.data_ptr() + 4), ptr(wsb), wsb.numel(),
                 C.c_void_p(st.cuda_stream))
        for st in streams[:len(jobs)]:
            call("optex_fence", C.c_void_p(st.cuda_stream))
            cur.wait_stream(st)
        ks = torch.stack([j[4] for j in jobs])
    return tensors, jobs, ks


def fit_pca_many_finish(state, *, round_k_to: int = 1, sweeps_out: Optional[list] = None):
    """Second half of `fit_pca_many`: the one synchronisation ([k, sweeps] per solve), the slices and the projections
    of optex.py:188 (on the current stream, which must be ordered after the launch's)."""
    if state is None:
        return []
    tensors, jobs, ks = state
    ks_sw = ks.tolist()
    if sweeps_out is not None:
        sweeps_out.extend(v[1] for v in ks_sw)
    out = []
    for t, (x, c, vecs, sigma, k_dev, wsb), (k, _) in zip(tensors, jobs, ks_sw):
        if round_k_to > 1:
            k = min(c, -(-k // round_k_to) * round_k_to)
        eigvecs = vecs[:, :k].contiguous()
        out.append((pca_project(t, eigvecs), eigvecs))
    return out


def install(reference_optex_module) -> None:
    """Rebind a loaded reference ``optex`` module to the B200 path.  Both names must be patched:
    ``optex`` did `from histmatch import hist_match` (optex.py:9), so patching ``histmatch`` alone
    would not reach `mix_style_features` (optex.py:200-201)."""
    reference_optex_module.optimal_transport = optimal_transport
    reference_optex_module.hist_match = _histmatch.hist_match
    reference_optex_module.random_rotation = random_rotation
    reference_optex_module.fit_pca = fit_pca
    from . import texture as _texture, util as _util, vgg as _vgg          # late: texture imports this module

    reference_optex_module.mix_style_features = _texture.mix_style_features
    reference_optex_module.resize = _util.resize                          # optex.py:10 `from util import resize`
    reference_optex_module.rgb_to_hls = _texture.rgb_to_hls               # optex.py:5 (kornia)
    reference_optex_module.hls_to_rgb = _texture.hls_to_rgb

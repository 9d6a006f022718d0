"""The synthesis loop of the reference (`OptimalTexture`, /root/reference/optex.py:15-139) on the B200 kernels.

Same constructor arguments, same `forward(pastiche, styles, content=None, verbose=False)`, same pass / layer /
iteration schedule (including the `[l - 1]` column quirk of optex.py:112) - but every tensor stays on the GPU and
every stage is one of this library's kernels:

    resize (util.py:105)                      -> optex_resize_bicubic_aa
    Encoder / Decoder (vgg.py)                -> optex_conv3x3 stacks (vgg.py of this package)
    fit_pca, `@ eigvecs`, `@ eigvecs.T`       -> optex_fit_pca, optex_pca_project
    content re-centring (optex.py:76)         -> optex_recentre
    mix_style_features (optex.py:193-206)     -> optex_hist_match x2 + optex_mix_features per layer
    the inner loop (optex.py:112-117)         -> ONE optex_ot_loop call per layer (rotations drawn on the device,
                                                 content blend fused into the inverse rotation)
    colour transfer (optex.py:124-137)        -> optex_lightness_transfer (+ 3 cdf OT steps for "opt")

Additive over the reference:
  * `models_dir=` / `state_dicts=`: where the weights come from (the reference reads ./models/*.pth, vgg.py:144,162);
  * the five encoder files are prefix-identical: when the loaded weights confirm it, styles and content
    are encoded ONCE per pass by Encoder(5).forward_all instead of five times (optex.py:62-63,72);
  * `rotations=` (callable (c, index) -> [c, c] tensor), `mixing_noise=` (callable shape -> uniform noise): inject
    what the reference draws from global RNG state, for parity tests; `pca=` (callable features -> (projected,
    eigvecs)) replaces `fit_pca` (an SVD basis is defined only up to sign / rotations of near-degenerate subspaces, so
    element-wise comparisons of the whole loop need both sides in ONE basis); `pca_round_k=32` keeps a few more
    components than the reference's 90 % rule so that every layer's channel count is a multiple of 32;
  * `pca_warm_start=True` (default off): each layer's PCA eigensolver starts from the basis it found in the previous
    pass; passes of EQUAL size share one style-side result (encode + PCA) within a forward();
  * `overlap_style=True` (default): the style side of pass p + 1 is launched on a second stream before pass p's
    layers and read back after them; same results (the host-side order of every RNG draw and kernel argument is
    unchanged; `overlap_style=False` is the serial schedule);
  * `no_multires=True` works (the reference's own path raises AttributeError at util.py:86: a list has no .tolist()).
"""
from __future__ import annotations

import contextlib
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib
from . import histmatch as _histmatch
from . import optex as _optex
from . import util as _util
from . import vgg as _vgg
from ._runtime import call, f32c, ptr, require_cuda, stream_ptr, workspace


def recentre(content_feature: Tensor, style_feature: Tensor) -> Tensor:
    """reference: optex.py:76  `content_feature - content_feature.mean() + torch.mean(style_features[l])`."""
    dev = require_cuda(content_feature, style_feature)
    x, s = f32c(content_feature), f32c(style_feature)
    out = torch.empty_like(x)
    wsb = workspace(dev, _lib.lib().optex_recentre_workspace_bytes())
    with torch.cuda.device(dev):
        call("optex_recentre", ptr(x), x.numel(), ptr(s), s.numel(), ptr(out), ptr(wsb), wsb.numel(), stream_ptr(dev))
    return out


def lightness_transfer(content: Tensor, pastiche: Tensor) -> Tensor:
    """reference: optex.py:126-128 - hls_to_rgb(H(content), L(pastiche), S(content)); NCHW [b,3,h,w] in [0,1]."""
    dev = require_cuda(content, pastiche)
    if content.shape != pastiche.shape or content.dim() != 4 or content.shape[1] != 3:
        raise ValueError(f"colour transfer needs two [b,3,h,w] images of one shape, got {tuple(content.shape)} "
                         f"and {tuple(pastiche.shape)}")
    c, p = f32c(content), f32c(pastiche)
    out = torch.empty_like(c)
    b, _, h, w = c.shape
    with torch.cuda.device(dev):
        call("optex_lightness_transfer", ptr(c), ptr(p), ptr(out), b, h * w, stream_ptr(dev))
    return out


def _hls(x: Tensor, fn: str) -> Tensor:
    dev = require_cuda(x)
    if x.dim() != 4 or x.shape[1] != 3:
        raise ValueError(f"expected [b,3,h,w], got {tuple(x.shape)}")
    xin = f32c(x)
    out = torch.empty_like(xin)
    b, _, h, w = xin.shape
    with torch.cuda.device(dev):
        call(fn, ptr(xin), ptr(out), b, h * w, stream_ptr(dev))
    return out


def rgb_to_hls(image: Tensor) -> Tensor:
    """kornia.color.hls.rgb_to_hls as imported at optex.py:5 (H in radians)."""
    return _hls(image, "optex_rgb_to_hls")


def hls_to_rgb(image: Tensor) -> Tensor:
    """kornia.color.hls.hls_to_rgb as imported at optex.py:5."""
    return _hls(image, "optex_hls_to_rgb")


def mix_style_features(style_features: List[Tensor], mixing_mask: Tensor, mixing_alpha: float, hist_mode: str):
    """reference: optex.py:193-206.  style_features[l]: [2, h, w, c]; mixing_mask [1, 1, mh, mw] of 0 / 1."""
    mask = f32c(mixing_mask).reshape(mixing_mask.shape[-2], mixing_mask.shape[-1])
    mh, mw = mask.shape
    out = []
    for sf in style_features:
        dev = require_cuda(sf, mask)
        if sf.shape[0] != 2:
            raise ValueError(f"mixing needs exactly two style images, got a batch of {sf.shape[0]}")
        A, B = f32c(sf[0:1]), f32c(sf[1:2])
        AtoB = _histmatch.hist_match(A, B, mode=hist_mode)
        BtoA = _histmatch.hist_match(B, A, mode=hist_mode)
        target = torch.empty_like(A)
        _, h, w, c = A.shape
        with torch.cuda.device(dev):
            call("optex_mix_features", ptr(A), ptr(B), ptr(AtoB), ptr(BtoA), ptr(mask), ptr(target), h, w, c, mh, mw,
                 float(mixing_alpha), stream_ptr(dev))
        out.append(target)
    return out


def _prefix_identical(enc5: _vgg.Encoder, others: Dict[int, _vgg.Encoder]) -> bool:
    for d, e in others.items():
        for a, b in zip(e.layers, enc5.layers):
            if a.w.shape != b.w.shape or not (torch.equal(a.w, b.w) and torch.equal(a.b, b.b)):
                return False
    return True


class OptimalTexture:
    """reference: optex.py:15-139."""

    def __init__(self, size: int = 512, iters: int = 500, passes: int = 5, hist_mode: str = "chol",
                 color_transfer: Optional[str] = None, content_strength: float = 0.1, style_scale: float = 1,
                 mixing_alpha: float = 0.5, no_pca: bool = False, no_multires: bool = False, *,
                 models_dir: Optional[str] = None, state_dicts: Optional[Dict[Tuple[str, int], dict]] = None,
                 device="cuda", rotations: Optional[Callable[[int, int], Tensor]] = None,
                 mixing_noise: Optional[Callable[[Tuple[int, int]], Tensor]] = None,
                 pca: Optional[Callable[[Tensor], Tuple[Tensor, Tensor]]] = None, pca_round_k: int = 1,
                 pca_warm_start: bool = False, overlap_style: bool = True):
        self.hist_mode = hist_mode
        self.color_transfer = color_transfer
        self.content_strength = content_strength
        self.style_scale = style_scale
        self.mixing_alpha = mixing_alpha
        self.use_pca = not no_pca
        self.passes = passes
        self.iters_per_pass_and_layer, self.sizes = _util.get_iters_and_sizes(size, iters, passes, not no_multires)
        self.device = torch.device(device)
        sd = state_dicts or {}
        self.depths = list(range(5, 0, -1))                                                     # optex.py:42-43
        self.encoders = [_vgg.Encoder(d, state_dict=sd.get(("encoder", d)), models_dir=models_dir, device=device)
                         for d in self.depths]
        self.decoders = [_vgg.Decoder(d, state_dict=sd.get(("decoder", d)), models_dir=models_dir, device=device)
                         for d in self.depths]
        self.shared_encoder = _prefix_identical(self.encoders[0], {e.depth: e for e in self.encoders[1:]})
        self.rotations = rotations
        self.mixing_noise = mixing_noise
        # pca_round_k = 32 rounds every layer's component count UP to a multiple of 32 (fit_pca(round_k_to=)): the
        # C x C products of the covariance modes then run on the tensor cores instead of the fp32 SIMT tiles
        self.pca = pca                      # None: the device PCA, all layers of a pass solved concurrently
        self.pca_round_k = pca_round_k
        # the style's PCA is refitted at every pass's size (optex.py:62-67).  True: each layer's eigensolver starts from
        # the basis it found in the previous pass (round-robin solver carrying V; fewer sweeps, but every sweep costs
        # 4-5x the blocked-order cold solve's - off by default since that solver exists)
        self.pca_warm_start = pca_warm_start
        # pass p + 1's style side (resize, encoders, PCA solves: <= 92 CTAs) runs on a second stream beside pass p's
        # OT loops - both are latency-bound and leave most of the device idle
        self.overlap_style = overlap_style
        self._side: Optional[torch.cuda.Stream] = None
        self._pca_bases: Dict[int, Tensor] = {}
        self._prepared_cache: Dict[tuple, tuple] = {}       # style side of equal-size passes within one forward()
        self._pca_fitted: set = set()                        # layers whose basis belongs to the current forward()
        self.pca_sweeps: List[List[int]] = []                # per pass: sweeps of the five solves (conv5_1 .. conv1_1)
        # pca / sym: zero channels appended up to a multiple of 32 before a layer's OT loop.  Exact: the padded
        # covariances are blockdiag(Sigma + I, I), their square roots block-diagonal, the padded channels stay 0 -
        # and every C x C product of the chains runs on the tensor cores instead of the fp32 SIMT tiles.
        self.pad_channels = 32
        self.ot_calls = 0
        self.last_pca_k: List[int] = []
        self.profile: Optional[Dict[str, list]] = None      # set to {} to collect CUDA-event pairs per stage

    def to(self, *args, **kwargs):        # the reference is an nn.Module and is `.to(pastiche)`-ed (optex.py:278)
        return self

    # ------------------------------------------------------------------------------------------ stages
    @contextlib.contextmanager
    def _stage(self, name: str):
        if self.profile is None:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.profile.setdefault(name, []).append((a, b))

    def stage_ms(self) -> Dict[str, float]:
        """Milliseconds per stage from the collected events (synchronises)."""
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in (self.profile or {}).items()}

    def _encode_all(self, image: Tensor) -> List[Tensor]:
        """Features of `image` for the five encoders, deepest first (the order of self.encoders)."""
        if self.shared_encoder:
            return self.encoders[0].forward_all(image)[::-1]
        return [enc(image) for enc in self.encoders]

    def _ot_layer(self, feature: Tensor, style: Tensor, hist_mode: str, iters: int, content: Optional[Tensor],
                  strength: float) -> Tensor:
        """optex.py:112-117 for one layer."""
        if iters <= 0:
            return feature
        c = feature.shape[-1]
        rots = None
        if self.rotations is not None:
            rots = torch.stack([f32c(self.rotations(c, self.ot_calls + i).to(feature.device)) for i in range(iters)])
        self.ot_calls += iters
        # PCA'd channel counts (23, 85, 181, 310 ...) are not multiples of 32, the tensor-core GEMMs' granularity: the
        # block is zero-padded to the next multiple and rotated by diag(R_c, I).  Exact for every mode - the padded
        # channels never mix with the real ones (block-diagonal rotation / covariance) and are dropped afterwards - and
        # 1.4-2.9x faster than the SIMT fallback (scripts/pad_probe.py).
        pad = (-c) % self.pad_channels if (self.pad_channels > 1 and c >= 8) else 0
        if pad:
            def widen(t):
                w = torch.zeros(*t.shape[:-1], c + pad, dtype=torch.float32, device=t.device)
                w[..., :c] = t
                return w

            if rots is None and hist_mode not in ("pca", "sym"):      # pca / sym do not depend on the rotation
                first = next(_optex._counter)
                for _ in range(max(iters - 1, 0)):
                    next(_optex._counter)
                rots = _optex.random_rotations(c, iters, feature.device, first_counter=first)
            if rots is not None:
                eye = torch.eye(c + pad, dtype=torch.float32, device=feature.device).repeat(iters, 1, 1)
                eye[:, :c, :c] = rots
                rots = eye
            out = _optex.ot_loop(widen(feature), widen(style), hist_mode, iters, rotations=rots,
                                 content=None if content is None else widen(content), content_strength=strength)
            return out[..., :c].contiguous()
        return _optex.ot_loop(feature, style, hist_mode, iters, rotations=rots, content=content,
                              content_strength=strength)

    def prepare_pass(self, pastiche_shape, styles: List[Tensor], content: Optional[Tensor], size: int,
                     mix: bool = True):
        """Everything of a pass that does not depend on the pastiche: optex.py:45-79 without the pastiche's own resize
        (:55) and, with mix=True, the style mixing of optex.py:97-101.  Returns (cont_size, style_features,
        style_eigvs, content_features); cont_size is the size the pastiche has to be resized to, or None."""
        return self.prepare_finish(self.prepare_launch(pastiche_shape, styles, content, size), mix=mix)

    def prepare_launch(self, pastiche_shape, styles: List[Tensor], content: Optional[Tensor], size: int):
        """First half of `prepare_pass`: resizes, encoders and the PCA solves are ENQUEUED, nothing is read back - the
        overlapped schedule calls this a whole pass ahead of `prepare_finish`.  Passes of equal size (the reference
        refits the same style at the same size, e.g. every pass of `--no_multires` or `--size 256`) share one result."""
        resized = pastiche_shape[-2] != size and pastiche_shape[-1] != size
        key = (size, resized, tuple((s.data_ptr(), tuple(s.shape)) for s in styles),
               None if content is None else (content.data_ptr(), tuple(content.shape)))
        if key in self._prepared_cache:
            return {"cached": self._prepared_cache[key], "styles": styles}
        cont_size = None
        if resized:
            with self._stage("resize"):
                style_tens = [_util.resize(s, _util.get_size(size, self.style_scale, s.shape[2], s.shape[3]))
                              for s in styles]
                if content is not None:
                    cont_size = _util.get_size(size, 1.0, content.shape[2], content.shape[3], oversize=True)
                    cont_tens = _util.resize(content, cont_size)
                else:
                    cont_size = (size, size)
                    cont_tens = None
        else:
            style_tens, cont_tens = styles, content

        with self._stage("encode_inputs"):
            per_style = [self._encode_all(s) for s in style_tens]
            per_content = self._encode_all(cont_tens) if cont_tens is not None else None
        raw = []
        for l in range(len(self.encoders)):
            feats = [ps[l] for ps in per_style]
            raw.append(feats[0] if len(feats) == 1 else torch.cat(feats))
        pending = None
        if self.use_pca and self.pca is None:
            with self._stage("fit_pca"):
                bases, warm = [], []
                for l, t in enumerate(raw):
                    if not self.pca_warm_start:         # cold solves take the blocked-order solver (no basis kept)
                        bases.append(None)
                        warm.append(False)
                        continue
                    c = t.shape[-1]
                    b = self._pca_bases.get(l)
                    hit = l in self._pca_fitted and b is not None and tuple(b.shape) == (c, c) and b.device == t.device
                    if not hit:
                        b = self._pca_bases[l] = torch.empty(c, c, dtype=torch.float64, device=t.device)
                    bases.append(b)
                    warm.append(bool(hit))
                pending = _optex.fit_pca_many_launch(raw, bases=bases, warm=warm)
                self._pca_fitted.update(range(len(raw)))
        return {"key": key, "cont_size": cont_size, "raw": raw, "per_content": per_content, "pending": pending,
                "styles": styles}

    def prepare_finish(self, launched, mix: bool = True):
        """Second half of `prepare_pass`: reads the k's of the PCA solves (the pass's one host synchronisation), slices
        and projects (optex.py:188, :72-76), and with mix=True mixes the styles (optex.py:97-101)."""
        styles = launched["styles"]
        if "cached" in launched:
            cont_size, style_features, style_eigvs, content_features, ks = launched["cached"]
            style_features, style_eigvs, content_features = list(style_features), list(style_eigvs), list(content_features)
            self.last_pca_k = list(ks)
        else:
            raw, per_content, cont_size = launched["raw"], launched["per_content"], launched["cont_size"]
            fitted = None
            if launched["pending"] is not None:
                with self._stage("fit_pca"):
                    sweeps: List[int] = []
                    fitted = _optex.fit_pca_many_finish(launched["pending"], round_k_to=self.pca_round_k,
                                                        sweeps_out=sweeps)
                    self.pca_sweeps.append(sweeps)
            style_features, style_eigvs, content_features = [], [], []
            for l in range(len(self.encoders)):
                sf = raw[l]
                eigvecs = None
                if self.use_pca:
                    if fitted is not None:
                        sf, eigvecs = fitted[l]
                    else:
                        with self._stage("fit_pca"):
                            sf, eigvecs = self.pca(sf)
                    style_eigvs.append(eigvecs)
                style_features.append(sf)
                if per_content is not None:
                    cf = per_content[l]
                    if self.use_pca:
                        cf = _optex.pca_project(cf, eigvecs)
                    content_features.append(recentre(cf, sf))
            self.last_pca_k = [int(v.shape[1]) for v in style_eigvs]    # conv5_1 .. conv1_1 of the latest pass
            if self.sizes.count(launched["key"][0]) > 1:     # only a size that comes again is worth keeping alive
                self._prepared_cache[launched["key"]] = (cont_size, list(style_features), list(style_eigvs),
                                                         list(content_features), list(self.last_pca_k))
        if self.use_pca and min(self.last_pca_k) < 1:
            # optex.py:185-186: k = first index whose cumulative singular-value share exceeds 0.9; a layer whose FIRST
            # singular value already does gets k = 0 and the reference then fails inside its matmuls on [.., 0] tensors
            raise ValueError(f"fit_pca kept 0 components for a layer (k per layer, deepest first: {self.last_pca_k}): "
                             "the style's first singular value alone exceeds 90 % of the total (optex.py:185)")

        if mix and len(styles) > 1:                                                              # optex.py:97-101
            shape = tuple(style_features[1].shape[1:3])
            dev = styles[0].device
            noise = self.mixing_noise(shape) if self.mixing_noise is not None else torch.rand(shape, device=dev)
            mixing_mask = torch.ceil(noise.to(dev) - self.mixing_alpha)[None, None, ...]
            style_features = mix_style_features(style_features, mixing_mask, self.mixing_alpha, self.hist_mode)
        return cont_size, style_features, style_eigvs, content_features

    def encode_inputs(self, pastiche: Tensor, styles: List[Tensor], content: Optional[Tensor], size: int):
        """reference: optex.py:45-79 (same name, same four results)."""
        cont_size, style_features, style_eigvs, content_features = self.prepare_pass(
            pastiche.shape, styles, content, size, mix=False)
        if cont_size is not None:
            with self._stage("resize"):
                pastiche = _util.resize(pastiche, cont_size)
        return pastiche, style_features, style_eigvs, content_features

    def _streams_for(self, pastiche: Tensor):
        """(main stream, side stream, event marking the start of forward()) of the overlapped schedule."""
        if not pastiche.is_cuda:
            return None
        main = torch.cuda.current_stream(pastiche.device)
        if self._side is None or self._side.device != pastiche.device:
            self._side = torch.cuda.Stream(device=pastiche.device)
        return main, self._side, main.record_event()

    @contextlib.contextmanager
    def _on_side(self, streams):
        """Run the body on the side stream of the overlapped schedule (with the GEMMs' second scratch slot)."""
        if streams is None:
            yield
            return
        main, side, start = streams
        lib = _lib.lib()
        prev = lib.optex_set_scratch_slot(1)         # the GEMMs' operand-split scratch: one per concurrent stream
        try:
            with torch.cuda.stream(side):
                yield
        finally:
            lib.optex_set_scratch_slot(prev)

    def _launch(self, streams, pastiche_shape, styles, content, size):
        """prepare_launch, on the side stream when there is one."""
        with self._on_side(streams):
            if streams is not None:
                streams[1].wait_event(streams[2])    # the inputs exist since forward() began; NOT the main stream's
            launched = self.prepare_launch(pastiche_shape, styles, content, size)   # later work (that is the overlap)
        launched["args"] = (tuple(pastiche_shape), styles, content, size)
        return launched

    def _finish(self, streams, launched, pastiche_shape):
        """prepare_finish (same stream as the launch).  Returns its four results + the event the main stream waits for."""
        shape, styles, content, size = launched["args"]
        if (shape[-2] != size and shape[-1] != size) != (pastiche_shape[-2] != size and pastiche_shape[-1] != size):
            launched = self._launch(streams, pastiche_shape, styles, content, size)   # the pastiche changed shape
        with self._on_side(streams):
            out = self.prepare_finish(launched)
            ready = streams[1].record_event() if streams is not None else None
        if streams is not None:
            for group in out[1:]:
                for t in group:
                    if t is not None and t.is_cuda:
                        t.record_stream(streams[0])  # allocated on `side`, consumed (and released) on `main`
        return (*out, ready)

    def forward(self, pastiche: Tensor, styles: List[Tensor], content: Optional[Tensor] = None,
                verbose: bool = False) -> Tensor:
        """reference: optex.py:81-139."""
        require_cuda(pastiche, *styles, content)
        self._pca_fitted.clear()                             # pass 0 of every call starts cold
        self._prepared_cache = {}
        self.pca_sweeps = []
        streams = self._streams_for(pastiche) if self.overlap_style else None
        launched = self._launch(streams, pastiche.shape, styles, content, self.sizes[0])
        for p in range(self.passes):
            if verbose:
                print(f"Pass {p}, size {self.sizes[p]}")
            cont_size, style_features, style_eigvs, content_features, ready = self._finish(streams, launched,
                                                                                           pastiche.shape)
            if cont_size is not None:                                                            # optex.py:55
                with self._stage("resize"):
                    pastiche = _util.resize(pastiche, cont_size)
            if ready is not None:
                streams[0].wait_event(ready)
            if self.overlap_style and p + 1 < self.passes:
                # the next pass's style side (resize, encode, PCA solves) only needs the style / content images: it is
                # enqueued now on the side stream and runs beside this pass's OT loops; its results are read (the k's:
                # the pass's one host synchronisation) when this pass has been enqueued
                launched = self._launch(streams, pastiche.shape, styles, content, self.sizes[p + 1])

            for l, (encoder, decoder) in enumerate(zip(self.encoders, self.decoders)):
                if verbose:
                    print(f"Layer: relu{(4 - l) + 1}_1")
                with self._stage("encode"):
                    feature = encoder(pastiche)
                    if self.use_pca:
                        feature = _optex.pca_project(feature, style_eigvs[l])
                blend = len(content_features) > 0 and l <= 2
                with self._stage("ot_loop"):
                    feature = self._ot_layer(feature, style_features[l], self.hist_mode,
                                             self.iters_per_pass_and_layer[p][l - 1],    # [l - 1]: optex.py:112
                                             content_features[l] if blend else None,
                                             self.content_strength / 2 ** (4 - l) if blend else 0.0)
                with self._stage("decode"):
                    if self.use_pca:
                        feature = _optex.pca_project(feature, style_eigvs[l], transpose=True)
                    pastiche = decoder(feature)

            if not self.overlap_style and p + 1 < self.passes:
                launched = self._launch(None, pastiche.shape, styles, content, self.sizes[p + 1])

        if self.color_transfer is not None:
            assert content is not None, "Color transfer requires content image"
            target = lightness_transfer(content, pastiche)
            if self.color_transfer == "opt":
                feature = _util.to_nhwc(pastiche).contiguous()
                target = _util.to_nhwc(target).contiguous()
                feature = self._ot_layer(feature, target, "cdf", 3, None, 0.0)
                pastiche = _util.to_nchw(feature)
            elif self.color_transfer == "lum":
                pastiche = target
        return pastiche

    __call__ = forward

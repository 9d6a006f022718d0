"""BASELINE.json configs[2..4] (and [1]) on the reference's REAL inputs (baseline/_ref: models/*.pth, style/, content/)
through optimaltextures_b200.OptimalTexture on one B200: seconds and output pixels / second per config.

    python scripts/run_configs.py [cfg ...]        cfg in {cfg1, cfg2, cfg3, cfg4}   (default: all)

cfg1  texture synthesis 512^2, graffiti.jpg, hist pca                                   (BASELINE configs[1])
cfg2  style transfer lava-small -> rocket 1024^2, content_strength 0.2, hist pca       (configs[2])
cfg3  texture mixing zebra + pattern-small 1024^2, mixing_alpha 0.5, hist pca           (configs[3])
cfg4  colour transfer green-paint-large -> city 2048^2, hist cdf, color_transfer opt   (configs[4])
Loading follows optex.py:264-272 (the reference's own util.load_styles / maybe_load_content)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import optimaltextures_b200 as ob
from baseline import reference
from optimaltextures_b200 import texture

CFG = {
    "cfg1": dict(style=["graffiti.jpg"], content=None, size=512, kw=dict(hist_mode="pca")),
    "cfg2": dict(style=["lava-small.jpg"], content="rocket.jpg", size=1024, kw=dict(hist_mode="pca", content_strength=0.2)),
    "cfg3": dict(style=["zebra.jpg", "pattern-small.jpg"], content=None, size=1024, kw=dict(hist_mode="pca", mixing_alpha=0.5)),
    "cfg4": dict(style=["green-paint-large.jpg"], content="city.jpg", size=2048,
                 kw=dict(hist_mode="cdf", color_transfer="opt")),
}


def main():
    ref = reference.load()
    root = ref.path
    sd = {}
    for d in range(1, 6):
        sd[("encoder", d)] = torch.load(os.path.join(root, "models", f"vgg_normalised_conv{d}_1.pth"), map_location="cpu")
        sd[("decoder", d)] = torch.load(os.path.join(root, "models", f"feature_invertor_conv{d}_1.pth"), map_location="cpu")
    lib = ob._lib.lib()
    out = {}
    for name in (sys.argv[1:] or list(CFG)):
        cfg = CFG[name]
        styles = ref.util.load_styles([os.path.join(root, "style", s) for s in cfg["style"]], size=cfg["size"], scale=1.0)
        content = ref.util.maybe_load_content(os.path.join(root, "content", cfg["content"]) if cfg["content"] else None,
                                              size=cfg["size"])
        torch.manual_seed(0)
        pastiche = torch.rand(content.shape if content is not None else (1, 3, cfg["size"], cfg["size"]))
        model = texture.OptimalTexture(size=cfg["size"], iters=500, passes=5, state_dicts=sd, **cfg["kw"])
        dev_styles = [s.cuda() for s in styles]
        dev_content = content.cuda() if content is not None else None
        for rep in range(2):
            ob.manual_seed(0)
            model.profile = {} if rep == 1 else None
            l0 = lib.optex_launch_count()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = model.forward(pastiche.cuda(), dev_styles, dev_content)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[name] = {"seconds": dt, "px_per_s": res.shape[0] * res.shape[2] * res.shape[3] / dt,
                     "out_shape": list(res.shape), "style_shapes": [list(s.shape) for s in styles],
                     "gpu_launches": int(lib.optex_launch_count() - l0), "finite": bool(torch.isfinite(res).all()),
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                     "stage_ms": {k: round(v, 1) for k, v in model.stage_ms().items()}, "args": cfg["kw"]}
        print(name, json.dumps(out[name]), flush=True)
        del model, res
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

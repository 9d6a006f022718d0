"""Time Encoder(5) / Decoder(5) (random weights) at 512^2 and 1024^2, per GEMM mode (CUDA events after warm-up)."""
import sys

import torch

sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import vgg
from oracle import vgg_oracle

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
enc = vgg.Encoder(5, state_dict=vgg_oracle.random_state_dict("encoder", 5))
dec = vgg.Decoder(5, state_dict=vgg_oracle.random_state_dict("decoder", 5))
x = torch.rand(1, 3, size, size, device="cuda")
for mode in ("auto", "tf32"):
    ob.set_gemm_mode(mode)
    f = enc(x)
    img = dec(f)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(3):
        f = enc(x)
    e[1].record()
    for _ in range(3):
        img = dec(f)
    e[2].record()
    torch.cuda.synchronize()
    flop_e = sum(2.0 * 9 * ci * co * (size >> s) ** 2 for (ci, co, s) in
                 [(3, 64, 0), (64, 64, 0), (64, 128, 1), (128, 128, 1), (128, 256, 2), (256, 256, 2), (256, 256, 2),
                  (256, 256, 2), (256, 512, 3), (512, 512, 3), (512, 512, 3), (512, 512, 3), (512, 512, 4)])
    te, td = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
    print(f"{size}^2 gemm={mode}: Encoder(5) {te:.2f} ms ({flop_e / te / 1e9:.0f} TFLOP/s)  Decoder(5) {td:.2f} ms  "
          f"feat {tuple(f.shape)} img {tuple(img.shape)}", flush=True)

"""Is the VGG stack host-bound at synthesis sizes?  Encoder(d) / Decoder(d) at 256^2 .. 512^2: device time between events
(back-to-back calls) against the host time to enqueue one call and the library's launch count."""
import sys, time
import torch
sys.path.insert(0, ".")
import optimaltextures_b200 as ob
from optimaltextures_b200 import vgg
from oracle import vgg_oracle

lib = ob._lib.lib()
for size in (256, 384, 512):
    for d in (5, 3, 1):
        enc = vgg.Encoder(d, state_dict=vgg_oracle.random_state_dict("encoder", d))
        dec = vgg.Decoder(d, state_dict=vgg_oracle.random_state_dict("decoder", d))
        x = torch.rand(1, 3, size, size, device="cuda")
        f = enc(x); img = dec(f)
        torch.cuda.synchronize()
        reps = 10
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        l0 = lib.optex_launch_count()
        e[0].record()
        t0 = time.perf_counter()
        for _ in range(reps):
            f = enc(x)
        t1 = time.perf_counter()
        e[1].record()
        for _ in range(reps):
            img = dec(f)
        t2 = time.perf_counter()
        e[2].record()
        torch.cuda.synchronize()
        n = (lib.optex_launch_count() - l0) / reps
        print(f"{size}^2 depth {d}: Encoder device {e[0].elapsed_time(e[1]) / reps * 1e3:7.1f} us, host enqueue {(t1 - t0) / reps * 1e6:7.1f} us | "
              f"Decoder device {e[1].elapsed_time(e[2]) / reps * 1e3:7.1f} us, host enqueue {(t2 - t1) / reps * 1e6:7.1f} us | launches enc+dec {n:.0f}", flush=True)

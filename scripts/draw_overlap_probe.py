"""Next-round experiment (NOT yet run on a B200): does the per-step rotation draw (29 us of the 241 us headline step,
FP64-pipe / latency bound) hide behind the step's own kernels when it is generated 16 matrices ahead on a second
stream?  Same C-ABI calls as bench.py's timed region; (a) the draw of all K rotations enqueued in front of the steps on
one stream (what bench.py times), (b) chunks of 16 drawn on a side stream, one chunk ahead, the steps waiting on
per-chunk events.  If (b) wins, optex_ot_loop gets the same double-buffered draw.  Usage: python scripts/draw_overlap_probe.py [K]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import optimaltextures_b200 as ob  # noqa: F401
from optimaltextures_b200 import _lib
from optimaltextures_b200._runtime import call, ptr, workspace

K = int(sys.argv[1]) if len(sys.argv) > 1 else 192
CHUNK, n, c = 16, 128 * 128, 512
dev = torch.device("cuda", 0)
lib = _lib.lib()
mode = _lib.mode_id("cdf")
g = torch.Generator().manual_seed(0)
sets = [(torch.relu(torch.randn(1, 128, 128, c, generator=g)).to(dev),
         torch.relu(1.3 * torch.randn(1, 128, 128, c, generator=g) + 0.2).to(dev)) for _ in range(4)]
outs = [torch.empty_like(sets[0][0]) for _ in range(2)]
ws = workspace(dev, lib.optex_ot_workspace_bytes(n, n, c, mode))
rots = torch.empty(K, c, c, device=dev)
rot_ws = [torch.empty(lib.optex_rotations_workspace_bytes(c, K), dtype=torch.uint8, device=dev) for _ in range(2)]
main = torch.cuda.current_stream(dev)
side = torch.cuda.Stream(device=dev)


def draw(first, count, stream, slot):
    call("optex_random_rotations", ptr(rots[first]), c, count, 1234, first, None, ptr(rot_ws[slot]),
         rot_ws[slot].numel(), C.c_void_p(stream.cuda_stream))


def step(i):
    p, s = sets[i % 4]
    call("optex_ot_step", ptr(p), ptr(s), ptr(rots[i]), ptr(outs[i % 2]), 1, n, 1, n, c, mode, 1.0, None, 0.0,
         ptr(ws), ws.numel(), C.c_void_p(main.cuda_stream))


def serial():
    draw(0, K, main, 0)
    for i in range(K):
        step(i)


def overlapped():
    ready = {}
    side.wait_stream(main)

    def enqueue_chunk(k):
        first = k * CHUNK
        if first >= K:
            return
        draw(first, min(CHUNK, K - first), side, k & 1)
        ready[k] = side.record_event()

    enqueue_chunk(0)
    for i in range(K):
        k = i // CHUNK
        if i % CHUNK == 0:
            enqueue_chunk(k + 1)             # one chunk ahead (its workspace slot was last used two chunks ago)
            main.wait_event(ready.pop(k))
        step(i)


for name, fn in (("draw in front, one stream", serial), ("draw 16 ahead on a side stream", overlapped)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    fence = torch.zeros(1, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        fn()
    fence.add_(1)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * K)
    print(f"{name}: {us:.1f} us/step = {1e6 / us:.0f} it/s", flush=True)

#!/bin/bash
mkdir -p gpurun_out
for shape in "1024 64" "512 128" "256 256"; do set -- $shape
timeout 300 python bench.py --hw $1 --channels $2 --steps 20 --warmup 3 --sets 1 --no-all-modes --no-layers --no-synthesis --no-cpu-baseline --no-tf32-peak --no-e2e > gpurun_out/r2_layer_$1.json 2>gpurun_out/r2_layer_$1.err; tail -2 gpurun_out/r2_layer_$1.err
python - <<P
import json
d=json.load(open("gpurun_out/r2_layer_$1.json"))
print("$1 x $2:", round(d["value"]), "it/s", round(d["ms_per_step"]*1e3,1), "us/step", {k:(round(v["ms"]*1e3,1),v["launches"],round(v["achieved"]),v["unit"]) for k,v in d["kernels"].items()}, "step hbm frac", round(d["step_roofline"]["hbm_frac"],3))
P
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_conv1.csv python - <<P > /dev/null 2>&1
import sys; sys.path.insert(0, ".")
import torch, optimaltextures_b200 as ob
g = torch.Generator(device="cuda").manual_seed(0)
p = torch.relu(torch.randn(1, 1024, 1024, 64, device="cuda", generator=g)); s = torch.relu(torch.randn(1, 1024, 1024, 64, device="cuda", generator=g))
rots = ob.random_rotations(64, 3, "cuda", seed=1)
for i in range(3): p = ob.optimal_transport(p, s, "cdf", rotation=rots[i])
torch.cuda.synchronize()
P
python scripts/summarize_ncu.py launches cdf 3 gpurun_out/r02_launches_conv1.csv

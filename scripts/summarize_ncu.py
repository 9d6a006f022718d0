"""Turn ncu CSV output into the small summaries committed under profiles/.

  launches <mode> <steps> <launches.csv> [...]   rows for profiles/*_launch_list_summary.csv   (stdout)
      launches.csv = ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <f> python scripts/prof_step.py <mode> auto <steps>
  full <raw.csv>                                 selected metrics per kernel of an `ncu -i x.ncu-rep --page raw --csv` dump
"""
import csv, re, sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def short(name):
    name = re.sub(r"optex::<unnamed>::|optex::\(anonymous namespace\)::|<unnamed>::|unnamed>::|void ", "", name)
    depth = 0
    for i, ch in enumerate(name):          # cut the argument list: the first '(' outside the template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            return name[:i].strip()
    return name.strip()


def launches(mode, steps, path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mn, mv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mn] != "gpu__time_duration.sum":
            continue
        k = short(r[kn])
        if k.startswith("at::") or "elementwise" in k or "reduce_kernel" in k or "distribution" in k:
            continue                      # torch's own kernels of the script (input generation, final sum)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", "")) / 1e3
    step_kernels = {k: v for k, v in agg.items() if not k.startswith("rot_")}
    total = sum(v[1] for v in step_kernels.values())
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = f"{us / total:.3f}" if k in step_kernels else ""
        print(f'{mode},{steps},"{k}",{n},{us:.1f},{us / n:.1f},{share}')


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(c) for c in ("ID", "Kernel Name", "Grid Size", "Block Size")] + [hdr.index(k) for k in KEEP if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in rows[2:]:
        out = [r[c] for c in cols]
        out[1] = short(out[1])
        w.writerow(out)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2])

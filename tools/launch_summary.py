"""Summarise an ncu launch list (gpu__time_duration.sum per launch, CSV) of a bench.py run.
usage: python tools/launch_summary.py gpurun_out/launches.csv "<command line profiled>" > profiles/<name>_summary.txt

'real' launches did work; the others belong to a batch of iterations that was enqueued after the
device-side `done` flag was set and return in a few microseconds (threshold: 20 us)."""
import csv
import statistics
import sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
rows = list(csv.reader(open(path)))
h0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0]


per = defaultdict(list)
for r in rows[h0 + 1:]:
    if len(r) <= vi:
        continue
    per[short(r[ki])].append(float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-3))

print("# ncu --metrics gpu__time_duration.sum --clock-control none: %s" % cmd)
print("# (cold-cache, serialised launches: compare SHARES, not absolutes).  'real' = launches that did work; the others are")
print("# iterations of a batch enqueued after the device-side `done` flag was set: they return in ~3 us.\n")
iter_kernels = OrderedDict((k, v) for k, v in per.items() if any(t in k for t in ("sell_spmv_kernel", "cg_fused_kernel<0, 1", "cg_fused_kernel<1, 1",
                                                                                  "cg_fused_kernel<3, 1", "cg_dir_kernel", "halo_", "cg_finalize")))
real = {k: [t for t in v if t > 20.0] for k, v in iter_kernels.items()}
step = sum(statistics.mean(v) for v in real.values() if v)
print("CG iteration kernels (real launches only):")
for k, v in iter_kernels.items():
    r = real[k]
    early = [t for t in v if t <= 20.0]
    if not r:
        continue
    print("  %-44s real launches %3d  mean %9.1f us  share of CG step %5.1f %%   (early-exit launches: %d, median %.1f us)"
          % (k[:44], len(r), statistics.mean(r), 100.0 * statistics.mean(r) / step, len(early), statistics.median(early) if early else 0.0))
print("\nall kernels of the command (setup included):")
tot = sum(sum(v) for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    print("  %-70s n=%4d total %10.1f us  %5.1f %%" % (k[:70], len(v), sum(v), 100.0 * sum(v) / tot))

"""Where do the warp-stall samples of a kernel fall?  Reads an `ncu --set full --import-source on` report (SASS page) and
prints the instructions that collect the most samples, grouped into runs of neighbouring instructions.

  python tools/ncu_stall_summary.py gpurun_out/r1q_prof_sell_bench.ncu-rep [--min-pct 0.7] > profiles/....txt
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--min-pct", type=float, default=0.7)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    name = rows[heads[0] - 1][1] if heads[0] > 0 else "?"
    h = rows[heads[0]]
    body = rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))]
    cs, ce = h.index("# Samples"), h.index("Instructions Executed")
    samples = [int(r[cs]) if len(r) > cs and r[cs].isdigit() else 0 for r in body]
    tot = sum(samples)
    print("# %s\n# kernel: %s (first captured launch), %d SASS instructions, %d warp-stall samples" % (a.report, name, len(body), tot))
    print("# instructions with >= %.1f %% of the samples (index, SASS, samples, share, times executed per warp-instruction)" % a.min_pct)
    for k, r in enumerate(body):
        if samples[k] >= a.min_pct / 100.0 * tot:
            print("%5d  %-62s %8d  %5.1f %%  exec %s" % (k, r[1].strip()[:62], samples[k], 100.0 * samples[k] / tot, r[ce]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-instruction warp-stall report of one kernel from an .ncu-rep captured with `--set full --import-source on`.

    python tools/stall_report.py gpurun_out/x.ncu-rep [--top 40] [--kernel-index 0]

Prints the stall-reason totals (share of all warp samples) and the instructions that collect the most samples with
their dominant stall reason -- the view that found the scoreboard-aliasing stall of the direct-space kernels and the
barrier stall of the list build in round 1 (profiles/README.md).
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = [i for i, r in enumerate(rows) if "# Samples" in r]
    for n, hi in enumerate(start):
        end = start[n+1]-1 if n+1 < len(start) else len(rows)
        name = rows[hi-1][1] if hi > 0 and len(rows[hi-1]) > 1 else "?"
        h = rows[hi]
        col = {c: i for i, c in enumerate(h)}
        data = [r for r in rows[hi+1:end] if len(r) == len(h)]
        total = sum(int(r[col["# Samples"]] or 0) for r in data)
        print("== %s: %d samples over %d instructions" % (name, total, len(data)))
        if not total:
            continue
        stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
        for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
            print("   %-24s %8d  %5.1f%%" % (s, v, 100.0*v/total))
        for r in sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:args.top]:
            st = {s: int(r[col[s]] or 0) for s in stalls}
            reason, count = max(st.items(), key=lambda x: x[1])
            print("   %5d  %-72s %7s  %s=%d  executed %s" % (data.index(r), r[col["Source"]].strip()[:72], r[col["# Samples"]], reason, count,
                                                             r[col["Instructions Executed"]]))


if __name__ == "__main__":
    main()

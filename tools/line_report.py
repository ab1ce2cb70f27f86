#!/usr/bin/env python
"""Executed warp instructions per CUDA source line of one kernel, from an .ncu-rep captured with --import-source on
(kernels compiled with -lineinfo).

    python tools/line_report.py gpurun_out/x.ncu-rep k_buildList [--top 40]
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    func, path, header, acc, total, samples = None, None, None, {}, 0, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            func = r[1]
        elif r and r[0] == "Line No":
            header = {c: i for i, c in enumerate(r)}
        elif header and func and args.kernel in func and r and r[0].isdigit():
            # (a source line containing commas inside a quoted string can confuse the CSV: take the columns from the right)
            off = len(r)-len(header)
            try:
                n = int(r[header["Instructions Executed"]+off] or 0)
                smp = int(r[header["# Samples"]+off] or 0)
            except ValueError:
                continue
            key = (path.split("/")[-1], int(r[0]), r[1].strip()[:100])
            acc[key] = acc.get(key, 0)+n
            samples[key] = samples.get(key, 0)+smp
            total += n
    print("%s: %d warp instructions attributed to source lines" % (args.kernel, total))
    for key, n in sorted(acc.items(), key=lambda x: -x[1])[:args.top]:
        print("%6.2f%% %6.2f%%s  %s:%d  %s" % (100.0*n/max(total, 1), 100.0*samples[key]/max(sum(samples.values()), 1), key[0], key[1], key[2]))


if __name__ == "__main__":
    main()

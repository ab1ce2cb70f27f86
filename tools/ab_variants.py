#!/usr/bin/env python
"""Build tagged variants of the CUDA core and write the script that A/B-benchmarks them in ONE gpurun call.

    python tools/ab_variants.py base: p14:-DPACKED_CTAS_PER_SM=14 "u4:-DDIRECT_UNROLL=4 -DDIRECT_CTAS_PER_SM=20"
    gpurun --timeout 900 -- 'bash scratch/ab_variants.sh'

Every argument is tag:nvcc-flags (an empty tag or flags = the product build).  Each variant is compiled into
openmm-lab_b200/csrc/libslicednb_b200_<tag>.so (build.py SNB_LIB_TAG / SNB_NVCC_EXTRA) and selected at run time through
SNB_B200_LIB (kernel.py), so that all of them travel in one snapshot and are measured on the same GPU, back to back:
numbers from different gpurun calls come from different boxes and differ by a few per cent on their own.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    variants = []
    for arg in sys.argv[1:]:
        tag, _, flags = arg.partition(":")
        variants.append((tag if tag != "base" else "", flags))
    if not variants:
        print(__doc__)
        return
    for tag, flags in variants:
        env = dict(os.environ, SNB_LIB_TAG=tag, SNB_NVCC_EXTRA=flags)
        print("building [%s] %s" % (tag or "base", flags), flush=True)
        subprocess.run([sys.executable, "-m", "openmm-lab_b200.build"], cwd=ROOT, env=env, check=True, stdout=subprocess.DEVNULL)
    lines = ["mkdir -p gpurun_out"]
    for tag, _ in variants:
        suffix = "_"+tag if tag else ""
        lines += ["export SNB_B200_LIB=$PWD/openmm-lab_b200/csrc/libslicednb_b200%s.so" % suffix,
                  "timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/ab%s.json 2> gpurun_out/ab%s.err" % (suffix, suffix),
                  "echo 'variant [%s]'; python tools/show_bench.py gpurun_out/ab%s.json | grep -v kernels" % (tag or "base", suffix)]
    path = os.path.join(ROOT, "scratch", "ab_variants.sh")
    with open(path, "w") as f:
        f.write("\n".join(lines)+"\n")
    print("wrote", path)


if __name__ == "__main__":
    main()

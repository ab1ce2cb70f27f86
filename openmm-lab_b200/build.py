"""Build recipe of the CUDA core: nvcc, sm_100a only, objects compiled in parallel, linked
in-tree into csrc/libslicednb_b200.so (git-ignored, travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libslicednb_b200.so")
SOURCES = ["api.cu", "comm.cu", "sort.cu", "neighbor.cu", "direct.cu", "direct_plain.cu", "direct_rf.cu", "direct_ewald.cu",
           "direct_ljpme.cu", "direct_packed.cu", "bonded.cu", "pme.cu", "pme_brick.cu", "ewald.cu", "finish.cu", "params.cu"]
HEADERS = ["snb_core.cuh", "kernels.cuh", "comm.cuh", "direct_impl.cuh", "../../include/slicednb.h", "../../include/snb_constants.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("SNB_NVCC_EXTRA", "").split()       # tuning experiments, e.g. -DDIRECT_CTAS_PER_SM=16
    tag = os.environ.get("SNB_LIB_TAG", "")                      # experiments build next to the product: libslicednb_b200_<tag>.so
    lib = LIB if not tag else LIB.replace(".so", "_%s.so" % tag)
    headers = [os.path.join(CSRC, h) for h in HEADERS]
    objdir = os.path.join(CSRC, "build"+("_"+tag if tag else ""))
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if force or _stale(obj, [os.path.join(CSRC, src)]+headers):
            jobs.append([nvcc]+NVCC_FLAGS+extra+["-c", os.path.join(CSRC, src), "-o", obj])

    def run(cmd):
        out = subprocess.run(cmd, capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s%s" % (" ".join(cmd), out.stdout, out.stderr))
        if verbose:
            print(" ".join(cmd[-3:]), file=sys.stderr)
        return out

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(lib, objs):
        run([nvcc, "-shared", "-o", lib]+objs+["-lcufft", "-ldl", "-Xlinker", "-rpath", "-Xlinker", "/usr/local/cuda/lib64"])
    return lib


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))

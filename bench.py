#!/usr/bin/env python
"""Benchmark of the SlicedNonbondedForce hot path (BASELINE.json metric: sliced-PME force
evaluations per second).

    python bench.py --gpus N --steps K --warmup W [--workload dhfr|stmv|apoa1|fep|water] [--impl reference]

One "step" = one evaluation (forces + energy + derivatives, direct + reciprocal space) of the
workload, including the full neighbour/tile-list rebuild -- exactly what one `execute` of the
reference kernel does (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:193-261).

  value     evaluations/s with the positions already resident in HBM (snb_execute_device)
  e2e       the same through the reference-facing call with HOST buffers (snb_execute):
            positions host->device, forces + energies device->host inside the timed region
  roofline  direct-space tile kernel: algorithmic flop (67 per interacting pair)
            / its CUDA-event duration, against the FP32 FMA peak (non-tensor path)
  cpu_baseline / --impl reference: the reference's own Reference-platform sources (oracle/_ref,
            compiled unchanged) timed on the host cores -- the one place bench.py executes oracle/.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_PAIR = 67.0          # forces + energy + slice accumulators
FP32_PEAK_TFLOPS = 148*128*2*1.965e9/1e12   # 74.4: SMs x lanes x 2 x max SM clock (MEASURED_PEAKS.json has no FP32 entry)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu="+q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(s[2+k].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz)//2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def make_workload(name):
    from openmmlab_b200 import systems
    return systems.WORKLOADS[name]()


def cpu_reference(w, seconds_budget=20.0, max_evals=3):
    """Time the reference's Reference-platform arithmetic (serial, as the reference is) on one host core."""
    import openmmlab_b200 as mm
    from oracle import oracle_kernel
    ctx = mm.Context(w.system, "Reference")
    ctx.setPositions(w.positions)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    times = []
    t_all = time.time()
    for _ in range(max_evals):
        ctx._forces = np.zeros((w.num_particles, 3))
        ctx._energyParameterDerivatives = {}
        t0 = time.time()
        kernel.execute(ctx, True, True, True, True)
        times.append(time.time()-t0)
        if time.time()-t_all > seconds_budget:
            break
    times.sort()
    med = times[len(times)//2]
    return {"value": 1.0/med, "unit": "evals/s", "cores": 1, "kind": "reference" if kernel.backend() == "reference" else "port",
            "sample": "%d full evaluation(s) of %s (neighbour list + direct + PME + exceptions), median; serial as the reference's Reference platform is"
                      % (len(times), w.name), "seconds_per_eval": med}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload)
    base = cpu_reference(w, seconds_budget=60.0, max_evals=max(1, min(args.steps, 5)))
    line = {"impl": "reference", "metric": "sliced-PME force evaluations/s", "value": base["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*base["seconds_per_eval"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, args), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ns_per_day_2fs": base["value"]*0.1728}
    print(json.dumps(line))


def config_of(w, args):
    f = w.force
    alpha, nx, ny, nz = f.getPMEParameters()
    return {"workload": w.name, "atoms": w.num_particles, "subsets": f.getNumSubsets(), "slices": f.getNumSlices(),
            "method": f.getNonbondedMethodName(), "cutoff_nm": f.getCutoffDistance(), "pme_grid": [nx, ny, nz],
            "scaling_parameters": f.getNumScalingParameters(), "derivatives": f.getNumScalingParameterDerivatives(),
            "evaluation": "forces+energy+derivatives, direct+reciprocal, tile list rebuilt every step",
            "l2": "every step rewrites the force/grid/list buffers; inputs are re-read from HBM after a 256 MiB L2 flush write",
            "precision": "mixed (FP32 pair arithmetic, FP64 / 2^32 fixed-point accumulation)"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import openmmlab_b200 as mm
    from openmmlab_b200 import _abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = make_workload(args.workload)
    n = w.num_particles
    platform = mm.Platform.getPlatformByName("B200")
    platform._properties["DeviceIndex"] = local
    ctx = mm.Context(w.system, platform)
    ctx.setPositions(w.positions)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
    box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
    flags = _abi.EXECUTE_ALL

    # device-resident inputs / outputs for `value`
    posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
    forces_fixed = torch.zeros((3, n), dtype=torch.int64, device="cuda")
    flush = torch.empty(256*1024*1024//4, dtype=torch.float32, device="cuda")

    def step_device():
        kernel.executeDevice(posq.data_ptr(), box, params, flags, forces_fixed.data_ptr())
        return kernel.readEnergies(flags)

    # host buffers for `e2e`: pinned positions in, forces + energies out
    pos_host = torch.from_numpy(w.positions.copy()).pin_memory()
    ctx._positions = pos_host.numpy()

    def step_host():
        ctx._forces = np.zeros((n, 3))
        ctx._energyParameterDerivatives = {}
        return kernel.execute(ctx, True, True, True, True)

    def timed(fn, steps, warmup, flush_l2):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        total = 0.0
        for _ in range(steps):
            if flush_l2:
                flush.zero_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()                      # returns after the stream synchronised (energies are read back)
            total += time.perf_counter()-t0
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    sampler.start()
    t_dev = timed(step_device, args.steps, args.warmup, True)
    # per-phase CUDA-event times (this mode serialises the two streams), averaged over a few more steps
    kernel.setTiming(True)
    step_device()
    acc = {k: 0.0 for k in kernel.getPhaseTimes()}
    reps = min(args.steps, 10)
    for _ in range(reps):
        flush.zero_()
        step_device()
        for k, v in kernel.getPhaseTimes().items():
            acc[k] += v/reps
    kernel.setTiming(False)
    t_e2e = timed(step_host, args.steps, max(3, args.warmup), True)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if rank != 0:
        return
    evals = args.steps*world if args.scaling == "weak" else args.steps
    value = evals/t_dev
    e2e = evals/t_e2e
    pairs = len(kernel.dumpPairs())
    direct_s = acc["direct"]*1e-3
    achieved = pairs*FLOP_PER_PAIR/direct_s/1e12 if direct_s > 0 else 0.0
    peaks = load_peaks()
    line = {"metric": "sliced-PME force evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*t_dev/args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(w, args),
            "ns_per_day_2fs": value*0.1728, "ns_per_day_4fs": value*0.3456,
            "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": int(n*24), "d2h_bytes_per_step": int(n*24+8*72+32),
                    "ms_per_step": 1e3*t_e2e/args.steps},
            "gpu_launches": int(kernel.getCounters()[2])*args.steps,
            "roofline": {"bound": "fp32", "kernel": "k_direct (direct-space tile kernel)", "achieved": achieved,
                         "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved/FP32_PEAK_TFLOPS,
                         "peak_source": "analytic 148 SM x 128 lanes x 2 x 1.965 GHz (FP32 FMA, non-tensor; MEASURED_PEAKS.json carries no FP32 entry)",
                         "interacting_pairs": pairs, "flop_per_pair": FLOP_PER_PAIR, "kernel_ms": acc["direct"], "traffic": None,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "phases_ms": acc, "clocks": sampler.summary()}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_reference(w)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default=None)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="weak")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "dhfr"
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

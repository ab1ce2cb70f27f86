#!/usr/bin/env python
"""Benchmark of the SlicedNonbondedForce hot path (BASELINE.json metric: sliced-PME force
evaluations per second, and ns/day).

    python bench.py --gpus N --steps K --warmup W [--workload dhfr|stmv|apoa1|fep|water] [--impl reference]

One "step" = one evaluation (forces + energy + derivatives, direct + reciprocal space) of the
workload, including the full neighbour/tile-list rebuild -- exactly what one `execute` of the
reference kernel does (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:193-261).

  value     evaluations/s with the positions already resident in HBM (snb_execute_device),
            host wall-clock around each synchronised step (>= the CUDA-event time, which is
            reported next to it as device_ms_per_step), max over ranks
  e2e       the same through the reference-facing call with HOST buffers (snb_execute):
            positions host->device, forces + energies device->host inside the timed region
  roofline  direct-space tile kernel: algorithmic flop (67 per interacting pair)
            / its CUDA-event duration, against the FP32 FMA peak (non-tensor path); the
            memory-bound PME kernels are listed against the measured HBM peak in "kernels"
  N > 1     --multi partition: ONE evaluation is shared by all ranks -- tile list and exceptions
            partitioned, reciprocal space on rank 0, fixed-point forces + slice energies all-reduced
            by NCCL ("scaling": "strong"); the path BASELINE.json asks for on LARGE systems (STMV).
            --multi replicas: every rank evaluates its own copy of the workload (the lambda windows
            of an alchemical calculation, one per GPU), no collective ("scaling": "weak").
            --multi auto (default): partition at >= 200 000 atoms, replicas below -- a 23 558-atom
            evaluation is launch/latency bound (145 us) and sharing it between GPUs gains nothing
            (measured 1.01x at N=2); the line then also carries the partitioned number
            ("partitioned") so that both are on record.
  stmv      the same measurement on the STMV-size workload (BASELINE.json configs[4]), a few steps,
            always partitioned for N > 1
  cpu_baseline / --impl reference: the reference's own Reference-platform sources (oracle/_ref,
            compiled unchanged) timed on the host cores -- the one place bench.py executes oracle/.
"""
import argparse
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
PACKED_MIN_ATOMS = 65536                    # include/slicednb.h SNB_PACKED_MIN_ATOMS: the packed-FP32 kernel takes over from here


def measured_fp32_peak():
    """FFMA throughput measured on a B200 of this pool with scratch/ubench/fp32peak.cu (profiles/r1_fp32_peak.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_fp32_peak.json")) as f:
            return json.load(f)["ffma_tflops"]
    except Exception:
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def load_traffic(workload):
    """DRAM bytes per launch of k_direct from the committed ncu --set full capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)[workload]["dram_bytes_per_launch"]
    except Exception:
        return None


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
    from oracle import oracle_kernel  # noqa: F401
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
                      % (len(times), w.name), "seconds_per_eval": med, "host_cores_available": os.cpu_count()}


def _reference_worker(name, evals, queue):
    w = make_workload(name)
    base = cpu_reference(w, seconds_budget=1e9, max_evals=evals)
    queue.put((base["seconds_per_eval"], base["kind"]))


def cpu_reference_all_cores(name, evals):
    """The Reference platform is serial; a host with C cores can still run C independent evaluations at once (the
    lambda windows of an alchemical calculation).  C worker processes, `evals` evaluations each, aggregate rate."""
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 64))
    ctx = mp.get_context("fork")
    queue = ctx.Queue()
    t0 = time.time()
    procs = [ctx.Process(target=_reference_worker, args=(name, evals, queue)) for _ in range(cores)]
    for p in procs:
        p.start()
    got = [queue.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.time()-t0
    med = sorted(g[0] for g in got)[len(got)//2]
    # the rate while all workers evaluate (workload construction excluded: each worker's own median time per evaluation)
    return {"value": cores/med, "unit": "evals/s", "cores": cores, "kind": got[0][1],
            "sample": "%d worker processes x %d full evaluation(s) of %s each (neighbour list + direct + PME + exceptions), one per core; "
                      "rate = workers / median seconds per evaluation under that load" % (cores, evals, name),
            "seconds_per_eval": med, "wall_seconds": wall, "host_cores_available": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_workload(args.workload)
    partition = args.gpus > 1 and (args.multi == "partition" or (args.multi == "auto" and w.num_particles >= PARTITION_MIN_ATOMS))
    if w.num_particles >= PARTITION_MIN_ATOMS:
        base = cpu_reference(w, seconds_budget=60.0, max_evals=max(1, min(args.steps, 2)))      # one evaluation takes most of a minute
    else:
        base = cpu_reference_all_cores(args.workload, max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": "sliced-PME force evaluations/s", "value": base["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*base["seconds_per_eval"],
            "higher_is_better": True, "scaling": "strong" if partition else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, args.gpus, partition), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ns_per_day_2fs": base["value"]*0.1728}
    print(json.dumps(line), flush=True)


PARTITION_MIN_ATOMS = 200000


def config_of(w, world=1, partition=True):
    f = w.force
    alpha, nx, ny, nz = f.getPMEParameters()
    return {"workload": w.name, "atoms": w.num_particles, "subsets": f.getNumSubsets(), "slices": f.getNumSlices(),
            "method": f.getNonbondedMethodName(), "cutoff_nm": f.getCutoffDistance(), "pme_grid": [nx, ny, nz],
            "scaling_parameters": f.getNumScalingParameters(), "derivatives": f.getNumScalingParameterDerivatives(),
            "evaluation": "forces+energy+derivatives, direct+reciprocal, tile list rebuilt every step",
            "parallelism": "1 GPU" if world == 1 else
                           ("%d ranks share each evaluation: tile list + exceptions partitioned, reciprocal space on rank 0, "
                            "positions replicated by the caller, one ncclAllReduce(int64) of fixed-point forces + slice energies per evaluation" % world
                            if partition else
                            "%d replicas: every rank evaluates its own copy of the workload, no collective (a %d-atom evaluation is "
                            "latency bound and does not shard profitably; the partitioned number is reported beside it)" % (world, w.num_particles)),
            "l2": "256 MiB flush write between timed steps; every step rewrites the force/grid/list buffers",
            "timing": "host wall-clock around each synchronised step, max over ranks (CUDA-event time reported as device_ms_per_step)",
            "precision": "mixed (FP32 pair arithmetic, FP64 / 2^32 fixed-point accumulation)"}


class Runner:
    """One workload on this rank's GPU (and, for world > 1, its share of every evaluation)."""

    def __init__(self, name, local, rank, world, dist, partition=True):
        import torch
        import openmmlab_b200 as mm
        from openmmlab_b200 import _abi
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        w = self.w = make_workload(name)
        n = self.n = w.num_particles
        platform = mm.Platform.getPlatformByName("B200")
        platform._properties["DeviceIndex"] = local
        self.ctx = ctx = mm.Context(w.system, platform)
        ctx.setPositions(w.positions)
        for k, v in w.parameters.items():
            ctx.setParameter(k, v)
        self.kernel = kernel = ctx._kernels[0][1]
        self.partition = partition and world > 1
        if self.partition:
            ident = [kernel.ncclUniqueId() if rank == 0 else None]
            dist.broadcast_object_list(ident, src=0)
            kernel.commInit(ident[0], rank, world)
            kernel.setCommOptions(broadcast_positions=False)    # every rank already holds the same positions (config says so)
        self.params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
        self.box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
        self.flags = _abi.EXECUTE_ALL
        # device-resident inputs / outputs for `value`
        self.posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        self.posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
        self.forces_fixed = torch.zeros((3, n), dtype=torch.int64, device="cuda")
        # host buffers for `e2e`: pinned positions in, forces + energies out
        self.pos_host = torch.from_numpy(w.positions.copy()).pin_memory()
        ctx._positions = self.pos_host.numpy()

    def step_device(self):
        self.kernel.executeDevice(self.posq.data_ptr(), self.box, self.params, self.flags, self.forces_fixed.data_ptr())
        return self.kernel.readEnergies(self.flags)

    def step_host(self):
        # the caller's force vector: kept and zeroed every step, as OpenMM's ContextImpl does with its own (the kernel adds into it)
        if getattr(self, "_forces_host", None) is None:
            self._forces_host = np.zeros((self.n, 3))
        self._forces_host.fill(0.0)
        self.ctx._forces = self._forces_host
        self.ctx._energyParameterDerivatives = {}
        return self.kernel.execute(self.ctx, True, True, True, True)

    def timed(self, fn, steps, warmup, flush):
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        total, dev = 0.0, 0.0
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()     # the ranks enter every shared evaluation together
            t0 = time.perf_counter()
            fn()                      # returns after the stream synchronised (energies are read back)
            total += time.perf_counter()-t0
            dev += self.kernel.getLastDeviceMs()*1e-3
        t = torch.tensor([total, dev], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    def phases(self, reps, flush):
        """Per-phase CUDA-event times (this mode serialises the two streams), averaged."""
        k = self.kernel
        k.setTiming(True)
        self.step_device()
        acc = {name: 0.0 for name in k.getPhaseTimes()}
        for _ in range(reps):
            flush.zero_()
            self.step_device()
            for name, v in k.getPhaseTimes().items():
                acc[name] += v/reps
        k.setTiming(False)
        return acc

    def measure(self, steps, warmup, flush, peaks, e2e=True):
        torch, dist = self.torch, self.dist
        if self.partition:
            # a shared evaluation measures its load balance for eight evaluations, freezes the partition, then runs one
            # more eagerly and captures the CUDA graph (NCCL inside): all of that belongs to the warm-up
            warmup = max(warmup, 12)
        t_dev, t_evt = self.timed(self.step_device, steps, warmup, flush)
        acc = self.phases(min(steps, 10), flush)
        # evaluations finished per step of the job: one when the ranks share it, one per rank when they run replicas
        units = 1 if (self.partition or self.world == 1) else self.world
        out = {"value": units*steps/t_dev, "ms_per_step": 1e3*t_dev/steps, "device_ms_per_step": 1e3*t_evt/steps}
        if e2e:
            t_e2e, _ = self.timed(self.step_host, steps, max(3, warmup), flush)
            n = self.n
            out["e2e"] = {"value": units*steps/t_e2e, "unit": "evals/s", "h2d_bytes_per_step": int(units*n*24),
                          "d2h_bytes_per_step": int(units*(n*24+8*72+32)), "ms_per_step": 1e3*t_e2e/steps}
        # roofline of the dominant kernel on the slowest rank: its own pairs / its own kernel time
        pairs = self.kernel.countPairs()
        mine = torch.tensor([acc["direct"], float(pairs)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            gathered = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine)
            rows = [g.cpu().numpy() for g in gathered]
        else:
            rows = [mine.cpu().numpy()]
        slow = max(range(len(rows)), key=lambda r: rows[r][0])
        kernel_ms, kernel_pairs = float(rows[slow][0]), int(rows[slow][1])
        achieved = kernel_pairs*FLOP_PER_PAIR/(kernel_ms*1e-3)/1e12 if kernel_ms > 0 else 0.0
        name = "k_directPacked (direct-space tile kernel, packed FP32)" if self.n >= PACKED_MIN_ATOMS else "k_direct (direct-space tile kernel)"
        measured = measured_fp32_peak()
        out["roofline"] = {"bound": "fp32", "kernel": name, "achieved": achieved,
                           "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved/FP32_PEAK_TFLOPS,
                           "peak_source": "analytic 148 SM x 128 lanes x 2 x 1.965 GHz (FP32 FMA, non-tensor; MEASURED_PEAKS.json carries no FP32 entry)",
                           "peak_measured_ffma": measured, "frac_of_measured": (achieved/measured if measured else None),
                           "interacting_pairs": kernel_pairs, "interacting_pairs_all_ranks": int(sum(r[1] for r in rows)),
                           "flop_per_pair": FLOP_PER_PAIR, "kernel_ms": kernel_ms, "rank": slow, "traffic": load_traffic(self.w.name)}
        # the memory-bound kernels against the measured HBM peak (algorithmic bytes)
        f = self.w.force
        _, nx, ny, nz = f.getPMEParameters()
        S, G, n = f.getNumSubsets(), nx*ny*nz, self.n
        hbm = peaks.get("hbm_gbs")
        bytes_of = {"spread": n*28+S*G*4, "conv": 2*S*nx*ny*(nz//2+1)*8, "interp": n*40+S*G*4}
        out["kernels"] = {k: {"ms": acc[k], "algorithmic_bytes": b, "gbs": b/(acc[k]*1e-3)/1e9 if acc[k] > 0 else None,
                              "frac_of_hbm_peak": (b/(acc[k]*1e-3)/1e9/hbm) if (hbm and acc[k] > 0) else None}
                          for k, b in bytes_of.items()}
        out["phases_ms"] = acc
        out["gpu_launches"] = int(self.kernel.getCounters()[2])*steps*units
        return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = load_peaks()
    flush = torch.empty(256*1024*1024//4, dtype=torch.float32, device="cuda")
    sampler = ClockSampler(local)
    sampler.start()
    from openmmlab_b200 import systems
    natoms = systems.WORKLOAD_ATOMS.get(args.workload, 0)
    partition = world > 1 and (args.multi == "partition" or (args.multi == "auto" and natoms >= PARTITION_MIN_ATOMS))
    main = Runner(args.workload, local, rank, world, dist, partition=partition)
    res = main.measure(args.steps, args.warmup, flush, peaks)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    shared = None
    if world > 1 and not partition:
        # on record beside the replicas: the same evaluation shared by all ranks
        del main
        main = Runner(args.workload, local, rank, world, dist, partition=True)
        shared = main.measure(args.steps, args.warmup, flush, peaks)
        shared = {k: shared[k] for k in ("value", "ms_per_step", "device_ms_per_step", "e2e", "phases_ms")}
        shared["scaling"] = "strong"
        shared["parallelism"] = config_of(main.w, world, True)["parallelism"]
    extra = None
    if not args.no_stmv and args.workload != "stmv":
        w = main.w
        del main
        stmv = Runner("stmv", local, rank, world, dist)
        extra = stmv.measure(min(args.steps, 10), 3, flush, peaks)
        extra["config"] = config_of(stmv.w, world)
        del stmv
        main_w = w
    else:
        main_w = main.w
    if rank == 0:
        value = res["value"]
        line = {"metric": "sliced-PME force evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "device_ms_per_step": res["device_ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if partition else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(main_w, world, partition), "ns_per_day_2fs": value*0.1728, "ns_per_day_4fs": value*0.3456,
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "kernels": res["kernels"],
                "phases_ms": res["phases_ms"], "clocks": sampler.summary()}
        if shared is not None:
            line["partitioned"] = shared
        if extra is not None:
            extra["ns_per_day_2fs"] = extra["value"]*0.1728
            extra["scaling"] = "strong" if world > 1 else "weak"
            line["stmv"] = extra
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_reference(main_w)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries the ONE JSON line and nothing else: whatever libraries print there (NCCL's version banner, ...)
    # goes to stderr instead -- file descriptor 1 is pointed at stderr, the JSON line is written to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="dhfr")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-stmv", action="store_true")
    ap.add_argument("--multi", default="auto", choices=["auto", "replicas", "partition"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

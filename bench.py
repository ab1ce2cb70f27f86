#!/usr/bin/env python
"""Benchmark of the SlicedNonbondedForce hot path (BASELINE.json metric: sliced-PME force
evaluations per second, and ns/day).

    python bench.py --gpus N --steps K --warmup W [--workload dhfr|stmv|apoa1|fep|water] [--impl reference]

One "step" = one evaluation (forces + energy + derivatives, direct + reciprocal space) of the
workload, including the full neighbour/tile-list rebuild -- exactly what one `execute` of the
reference kernel does (platforms/reference/src/ReferenceOpenMMLabKernels.cpp:193-261).

Workload of the main line: DHFR-size (BASELINE.json configs[1]) on one GPU; STMV-size (configs[4]), ONE evaluation
shared by all ranks ("scaling": "strong"), on N > 1 GPUs -- the configuration north_star names for multi-GPU.
--workload overrides either.

  value         evaluations/s with the positions already resident in HBM (snb_execute_device),
                host wall-clock around each synchronised step (>= the CUDA-event time, which is
                reported next to it as device_ms_per_step), max over ranks
  e2e           the same through the reference-facing call with HOST buffers (snb_execute):
                positions host->device, forces + energies device->host inside the timed region
                (N > 1: ONE host->device copy on rank 0 + ncclBroadcast, forces read back on rank 0 only)
  parity        residuals of the workload that is timed, in the mode that is timed, against the reference's own
                sources (oracle/_ref; fixture generated from it for the STMV-size workload) by the reference's
                metric (AssertionUtilities.h:7-37): total energy, worst slice, worst derivative, per-atom force
                metric, rms / max force error over the rms force, pair-set equality (count + SHA-256); the same
                for Precision=double, whose throughput is value_double
  roofline      direct-space tile kernel: algorithmic flop per interacting pair (67 for PME Coulomb+LJ
                with energies and derivatives, 105 for LJPME) / its CUDA-event duration, against the FP32 FMA peak
                (non-tensor path); the memory-bound PME kernels are listed against the measured HBM peak in "kernels"
  md            (N = 1) the timed workload along a trajectory -- positions rewritten on the device before every step --
                with the tile list rebuilt every evaluation (as on the main line) and with platform property ListSkin
                (list built for cutoff + skin, kept until an atom has moved skin/2; rebuild count reported)
  N > 1         tile list and exceptions partitioned, reciprocal space on rank 0 overlapping the ncclReduce of the
                other ranks' fixed-point forces.  "replicas" = every rank evaluates its own DHFR-size copy (the lambda
                windows of an alchemical calculation), no collective -- reported as a sub-object.
  stmv / apoa1 / fep   (N = 1) the other BASELINE.json configurations as short sub-objects
  cpu_baseline / --impl reference: the reference's own Reference-platform sources (oracle/_ref,
                compiled unchanged) timed on the host cores -- the one place bench.py executes oracle/.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FP32_PEAK_TFLOPS = 148*128*2*1.965e9/1e12   # 74.4: SMs x lanes x 2 x max SM clock (MEASURED_PEAKS.json has no FP32 entry)
PARTITION_MIN_ATOMS = 200000
SAMPLE_STRIDE_DEFAULT = 1


def flop_per_pair(force):
    """Algorithmic flop per interacting pair of the direct-space kernel: geometry 10 + Coulomb
    (Ewald erfc 25, reaction field / plain 8) + LJ 11 + apply 12 + energies and slice accumulators 9; LJPME adds 38,
    the switching function 14."""
    method = force.getNonbondedMethodName()
    flop = 10+11+12+9+(25 if method in ("Ewald", "PME", "LJPME") else 8)
    if method == "LJPME":
        flop += 38
    elif force.getUseSwitchingFunction():
        flop += 14
    return float(flop)


def measured_fp32_peak():
    """FFMA throughput measured on a B200 of this pool with tools/ubench/fp32peak.cu (profiles/r1_fp32_peak.json), or None."""
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
    """DRAM bytes per launch of the direct-space kernel from the committed ncu --set full capture (profiles/traffic.json), or None."""
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


# ---------------------------------------------------------------------------------------------------------------------
# parity: the reference's metric (openmmapi/include/internal/AssertionUtilities.h:7-37)

def digest(a):
    h = hashlib.sha256()
    flat = np.ascontiguousarray(a).reshape(-1)
    for k in range(0, flat.size, 1 << 24):
        h.update(flat[k:k+(1 << 24)].tobytes())
    return h.hexdigest()


def residuals(ref, got, stride=1):
    """ref / got: dicts with energy, slices [L][2], derivs {name: value}, forces ([N][3]; ref may hold a strided sample)."""
    rel = lambda a, b: abs(a-b)/max(abs(a), 1.0)    # noqa: E731
    fr = np.asarray(ref["forces"], dtype=float)
    fg = np.asarray(got["forces"], dtype=float)[::stride]
    scale = np.maximum(np.sqrt((fr**2).sum(axis=1)), 1.0)
    dF = fr-fg
    rms_f = ref.get("rms_force") or max(float(np.sqrt((fr**2).sum(axis=1).mean())), 1.0)
    L = ref["slices"].shape[0]
    worst = max(((rel(ref["slices"][s, t], got["slices"][s, t]), s, t) for s in range(L) for t in range(2)))
    return {"energy_rel": rel(ref["energy"], got["energy"]),
            "worst_slice_rel": worst[0],
            "worst_slice": {"slice": worst[1], "term": "Coulomb" if worst[2] == 0 else "vdW", "reference": float(ref["slices"][worst[1], worst[2]]),
                            "abs_error": float(abs(ref["slices"][worst[1], worst[2]]-got["slices"][worst[1], worst[2]]))},
            "worst_derivative_rel": max([rel(ref["derivs"][k], got["derivs"][k]) for k in ref["derivs"]] or [0.0]),
            "force_per_atom_metric": float((np.abs(dF).max(axis=1)/scale).max()),
            "force_rms_over_rms_force": float(np.sqrt((dF**2).sum(axis=1).mean())/rms_f),
            "force_max_over_rms_force": float(np.abs(dF).max()/rms_f),
            "atoms_compared": int(fr.shape[0])}


def evaluate_host(platform, w, properties=None, pairs=False, count=False):
    """One kernel.execute through the reference-facing call with host buffers."""
    import openmmlab_b200 as mm
    ctx = mm.Context(w.system, platform, properties)
    ctx.setPositions(w.positions)
    for k, v in w.parameters.items():
        ctx.setParameter(k, v)
    kernel = ctx._kernels[0][1]
    ctx._forces = np.zeros((w.num_particles, 3))
    ctx._energyParameterDerivatives = {}
    e = kernel.execute(ctx, True, True, True, True)
    out = {"energy": e, "forces": ctx._forces.copy(), "slices": kernel.sliceEnergies.copy(), "derivs": dict(ctx._energyParameterDerivatives)}
    if pairs:
        p = kernel.dumpPairs()
        out["pair_count"], out["pair_sha256"] = int(len(p)), digest(p.astype(np.int32, copy=False))
    elif count:
        out["pair_count"] = int(kernel.countPairs())
    if platform == "B200":
        out["packed"] = bool(kernel.getCounters()[5])
    else:
        out["backend"] = kernel.backend()
    return out


def load_fixture(w):
    """Large fixture generated from oracle/_ref in the build container (tests/golden/make_golden_large.py), or None."""
    path = os.path.join(ROOT, "tests", "golden", w.name+"_large.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    if str(z["positions_sha256"]) != digest(w.positions):
        return None
    names = [str(s) for s in z["deriv_names"]]
    return {"energy": float(z["total_energy"]), "slices": z["total_slices"], "forces": z["total_forces_sample"],
            "derivs": dict(zip(names, z["total_derivs"])), "rms_force": float(z["total_rms_force"]),
            "pair_count": int(z["pair_count"]), "pair_sha256": str(z["pair_sha256"]), "stride": int(z["sample_stride"]),
            "backend": "reference (fixture tests/golden/%s_large.npz, generated from oracle/_ref)" % w.name}


def parity_report(w, local, live_max_atoms=200000):
    """Residuals of the CUDA path on workload `w` against the reference's own arithmetic: mixed (the mode that is
    timed, default kernel choice) and Precision=double."""
    from oracle import oracle_kernel  # noqa: F401  (registers the CPU platforms; checker only)
    if w.num_particles <= live_max_atoms:
        ref = evaluate_host("Reference", w, pairs=True)
        ref["stride"] = 1
        source = "oracle/_ref run here (the reference's Reference-platform sources compiled unchanged)" if ref["backend"] == "reference" \
            else "oracle port run here (oracle/_ref is not built on this box)"
    else:
        ref = load_fixture(w)
        if ref is None:
            return {"unavailable": "no fixture for %s and too large for a live CPU run" % w.name}
        source = ref["backend"]
    out = {"against": source, "metric": "AssertionUtilities.h:7-37: |x-ref|/max(|ref|,1); forces per atom |dF|_inf/max(|F_ref|,1)",
           "tolerance_north_star": 1e-5}
    for mode, props in (("mixed", {"DeviceIndex": local}), ("double", {"DeviceIndex": local, "Precision": "double"})):
        got = evaluate_host("B200", w, props, pairs=True)
        r = residuals(ref, got, ref["stride"])
        r["pair_count"] = got["pair_count"]
        r["pair_set_equal"] = bool(got["pair_count"] == ref["pair_count"] and got["pair_sha256"] == ref["pair_sha256"])
        r["direct_kernel"] = "k_directPacked" if got["packed"] else "k_direct"
        r["within_1e-5"] = bool(max(r["energy_rel"], r["worst_slice_rel"], r["worst_derivative_rel"], r["force_per_atom_metric"]) <= 1e-5)
        out[mode] = r
    return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference arm

def cpu_reference(w, evals=3, warmup=0, seconds_budget=20.0):
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
    for it in range(warmup+evals):
        ctx._forces = np.zeros((w.num_particles, 3))
        ctx._energyParameterDerivatives = {}
        t0 = time.time()
        kernel.execute(ctx, True, True, True, True)
        if it >= warmup:
            times.append(time.time()-t0)
        if times and time.time()-t_all > seconds_budget:
            break
    total = sum(times)
    times.sort()
    med = times[len(times)//2]
    return {"value": 1.0/med, "unit": "evals/s", "cores": 1, "kind": "reference" if kernel.backend() == "reference" else "port",
            "sample": "%d full evaluation(s) of %s (neighbour list + direct + PME + exceptions), median; serial as the reference's Reference platform is"
                      % (len(times), w.name), "seconds_per_eval": med, "evals_timed": len(times), "seconds_timed": total,
            "host_cores_available": os.cpu_count()}


def _reference_worker(name, evals, warmup, budget, queue):
    w = make_workload(name)
    base = cpu_reference(w, evals=evals, warmup=warmup, seconds_budget=budget)
    queue.put((base["seconds_per_eval"], base["kind"], base["evals_timed"]))


def cpu_reference_all_cores(name, evals, warmup, workers, budget):
    """The Reference platform is serial; a host with C cores can still run C independent evaluations at once (the
    lambda windows of an alchemical calculation).  `workers` processes, `evals` evaluations each, aggregate rate."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    queue = ctx.Queue()
    t0 = time.time()
    procs = [ctx.Process(target=_reference_worker, args=(name, evals, warmup, budget, queue)) for _ in range(workers)]
    for p in procs:
        p.start()
    got = [queue.get() for _ in procs]
    for p in procs:
        p.join()
    wall = time.time()-t0
    med = sorted(g[0] for g in got)[len(got)//2]
    timed = min(g[2] for g in got)
    # the rate while all workers evaluate (workload construction excluded: each worker's own median time per evaluation)
    return {"value": workers/med, "unit": "evals/s", "cores": workers, "kind": got[0][1],
            "sample": "%d worker processes x %d timed evaluation(s) of %s each after %d warm-up (neighbour list + direct + PME + exceptions), "
                      "one process per core; rate = workers / median seconds per evaluation under that load" % (workers, timed, name, warmup),
            "seconds_per_eval": med, "evals_timed": timed, "wall_seconds": wall, "host_cores_available": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload or ("dhfr" if args.gpus == 1 else "stmv")
    w = make_workload(name)
    partition = args.gpus > 1 and w.num_particles >= PARTITION_MIN_ATOMS
    cores = max(1, min(os.cpu_count() or 1, 64))
    # a bounded sample: the requested steps, cut short by a time budget (the counts that really ran are what is printed)
    if w.num_particles >= PARTITION_MIN_ATOMS:
        workers = max(1, min(cores, 4))          # ~6 GB and ~40 s per evaluation
        base = cpu_reference_all_cores(name, max(1, min(args.steps, 2)), 0, workers, 90.0)
    else:
        base = cpu_reference_all_cores(name, max(1, args.steps), max(0, min(args.warmup, 3)), cores, 120.0)
    line = {"impl": "reference", "metric": "sliced-PME force evaluations/s", "value": base["value"], "unit": "evals/s",
            "n_gpus": args.gpus, "steps": base["evals_timed"], "steps_requested": args.steps, "warmup": min(args.warmup, 3) if w.num_particles < PARTITION_MIN_ATOMS else 0,
            "ms_per_step": 1e3/base["value"], "ms_per_step_per_worker": 1e3*base["seconds_per_eval"],
            "higher_is_better": True, "scaling": "strong" if partition else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(w, args.gpus, partition), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "ns_per_day_2fs": base["value"]*0.1728}
    print(json.dumps(line), flush=True)


def config_of(w, world=1, partition=True):
    f = w.force
    alpha, nx, ny, nz = f.getPMEParameters()
    return {"workload": w.name, "atoms": w.num_particles, "subsets": f.getNumSubsets(), "slices": f.getNumSlices(),
            "method": f.getNonbondedMethodName(), "cutoff_nm": f.getCutoffDistance(), "pme_grid": [nx, ny, nz],
            "scaling_parameters": f.getNumScalingParameters(), "derivatives": f.getNumScalingParameterDerivatives(),
            "evaluation": "forces+energy+derivatives, direct+reciprocal, tile list rebuilt every step",
            "parallelism": "1 GPU" if world == 1 else
                           ("%d ranks share each evaluation: tile list + exceptions partitioned, reciprocal space on rank 0 overlapping "
                            "one grouped ncclReduce(int64 fixed-point forces -> rank 0) + ncclAllReduce(slice energies, flags); results on rank 0" % world
                            if partition else
                            "%d replicas: every rank evaluates its own copy of the workload, no collective" % world),
            "l2": "256 MiB flush write between timed steps; every step rewrites the force/grid/list buffers",
            "timing": "host wall-clock around each synchronised step, max over ranks (CUDA-event time reported as device_ms_per_step)",
            "precision": "mixed (FP32 pair arithmetic, FP64 / 2^32 fixed-point accumulation)"}


class Runner:
    """One workload on this rank's GPU (and, for world > 1, its share of every evaluation)."""

    def __init__(self, name, local, rank, world, dist, partition=True, properties=None, workload=None):
        import torch
        import openmmlab_b200 as mm
        from openmmlab_b200 import _abi
        self.torch, self.dist, self.rank, self.world = torch, dist, rank, world
        w = self.w = workload if workload is not None else make_workload(name)
        n = self.n = w.num_particles
        props = {"DeviceIndex": local}
        props.update(properties or {})
        self.ctx = ctx = mm.Context(w.system, "B200", props)
        ctx.setPositions(w.positions)
        for k, v in w.parameters.items():
            ctx.setParameter(k, v)
        self.kernel = kernel = ctx._kernels[0][1]
        self.partition = partition and world > 1
        if self.partition:
            ident = [kernel.ncclUniqueId() if rank == 0 else None]
            dist.broadcast_object_list(ident, src=0)
            kernel.commInit(ident[0], rank, world)
        self.params = np.array([ctx.getParameter(name) for name in kernel.globalNames], dtype=np.float64)
        self.box = np.ascontiguousarray(w.box, dtype=np.float64).reshape(9)
        self.flags = _abi.EXECUTE_ALL
        # device-resident inputs / outputs for `value` (N > 1: already replicated on every rank, no broadcast)
        self.posq = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
        self.posq[:, :3] = torch.from_numpy(w.positions.astype(np.float32)).cuda()
        self.forces_fixed = torch.zeros((3, n), dtype=torch.int64, device="cuda")
        # host buffers for `e2e`: pinned positions in (N > 1: rank 0's are the ones that count), forces + energies out
        self.pos_host = torch.from_numpy(w.positions.copy()).pin_memory()
        ctx._positions = self.pos_host.numpy()

    def step_device(self):
        self.kernel.executeDevice(self.posq.data_ptr(), self.box, self.params, self.flags, self.forces_fixed.data_ptr())
        return self.kernel.readEnergies(self.flags)

    def step_host(self):
        # the caller's force vector: kept and zeroed every step, as OpenMM's ContextImpl does with its own (the kernel adds into it)
        if getattr(self, "_forces_host", None) is None:
            self._forces_host = np.zeros((self.n, 3))
        if self.rank == 0 or not self.partition:
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

    def measure_md(self, steps, warmup, flush, sigma=0.001, period=40):
        """Device-resident evaluations/s along a trajectory: the positions change EVERY step, as between the steps of a
        simulation (ballistic motion with a per-component velocity of `sigma` nm/step, reversed every `period` steps so
        that atoms never overlap), positions written by a torch kernel outside the timed region."""
        torch = self.torch
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234)
        vel = torch.randn((self.n, 3), generator=gen, device="cuda", dtype=torch.float32)*sigma
        base = self.posq[:, :3].clone()

        def place(k):
            phase = k % (2*period)
            self.posq[:, :3] = base+float(phase if phase <= period else 2*period-phase)*vel

        for k in range(warmup):
            place(k)
            self.step_device()
        c0 = self.kernel.getCounters()
        total, dev = 0.0, 0.0
        for k in range(warmup, warmup+steps):
            place(k)
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            self.step_device()
            total += time.perf_counter()-t0
            dev += self.kernel.getLastDeviceMs()*1e-3
        c1 = self.kernel.getCounters()
        self.posq[:, :3] = base
        return {"value": steps/total, "ms_per_step": 1e3*total/steps, "device_ms_per_step": 1e3*dev/steps,
                "rebuilds": int(c1[6]-c0[6]), "steps": steps}

    def phases(self, reps, flush):
        """Per-phase CUDA-event times (this mode serialises the streams), averaged."""
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

    def measure(self, steps, warmup, flush, peaks, e2e=True, detail=True):
        torch, dist = self.torch, self.dist
        if self.partition:
            # a shared evaluation measures its load balance for eight evaluations, freezes the partition, then runs one
            # more eagerly and captures the CUDA graph (NCCL inside): all of that belongs to the warm-up
            warmup = max(warmup, 12)
            self.kernel.setCommOptions(broadcast_positions=False)    # device-resident inputs are already on every rank
        t_dev, t_evt = self.timed(self.step_device, steps, warmup, flush)
        # evaluations finished per step of the job: one when the ranks share it, one per rank when they run replicas
        units = 1 if (self.partition or self.world == 1) else self.world
        out = {"value": units*steps/t_dev, "ms_per_step": 1e3*t_dev/steps, "device_ms_per_step": 1e3*t_evt/steps}
        if e2e:
            if self.partition:
                self.kernel.setCommOptions(broadcast_positions=True)     # host positions: rank 0 uploads, NVLink broadcast
            t_e2e, _ = self.timed(self.step_host, steps, max(3, warmup), flush)
            n = self.n
            copies = 1 if self.partition else units         # shared evaluation: rank 0 alone copies in and out
            out["e2e"] = {"value": units*steps/t_e2e, "unit": "evals/s", "h2d_bytes_per_step": int(copies*n*24),
                          "d2h_bytes_per_step": int(copies*n*24+units*(8*72+32)), "ms_per_step": 1e3*t_e2e/steps}
            if self.partition:
                self.kernel.setCommOptions(broadcast_positions=False)
        acc = self.phases(min(steps, 10), flush)
        # roofline of the dominant kernel on the slowest rank: its own pairs / its own kernel time
        pairs = self.kernel.countPairs()
        packed = bool(self.kernel.getCounters()[5])
        mine = torch.tensor([acc["direct"], float(pairs)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            gathered = [torch.zeros_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine)
            rows = [g.cpu().numpy() for g in gathered]
        else:
            rows = [mine.cpu().numpy()]
        slow = max(range(len(rows)), key=lambda r: rows[r][0])
        kernel_ms, kernel_pairs = float(rows[slow][0]), int(rows[slow][1])
        fpp = flop_per_pair(self.w.force)
        achieved = kernel_pairs*fpp/(kernel_ms*1e-3)/1e12 if kernel_ms > 0 else 0.0
        name = "k_directPacked (direct-space tile kernel, packed FP32)" if packed else "k_direct (direct-space tile kernel)"
        measured = measured_fp32_peak()
        out["roofline"] = {"bound": "fp32", "kernel": name, "achieved": achieved,
                           "peak": FP32_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved/FP32_PEAK_TFLOPS,
                           "peak_source": "analytic 148 SM x 128 lanes x 2 x 1.965 GHz (FP32 FMA, non-tensor; MEASURED_PEAKS.json carries no FP32 entry)",
                           "peak_measured_ffma": measured, "frac_of_measured": (achieved/measured if measured else None),
                           "interacting_pairs": kernel_pairs, "interacting_pairs_all_ranks": int(sum(r[1] for r in rows)),
                           "flop_per_pair": fpp, "kernel_ms": kernel_ms, "rank": slow, "traffic": load_traffic(self.w.name),
                           "list_fill": kernel_pairs/(256.0*self.kernel.getCounters()[1]) if self.kernel.getCounters()[1] else None}
        out["gpu_launches"] = int(self.kernel.getCounters()[2])*steps*units
        if detail:
            # the memory-bound kernels against the measured HBM peak (algorithmic bytes)
            f = self.w.force
            _, nx, ny, nz = f.getPMEParameters()
            S, G, n = f.getNumSubsets(), nx*ny*nz, self.n
            hbm = peaks.get("hbm_gbs")
            bytes_of = {"spread": n*28+S*G*4, "conv": 2*S*nx*ny*(nz//2+1)*8, "interp": n*40+S*G*4}
            out["kernels"] = {k: {"ms": round(acc[k], 5), "algorithmic_bytes": b, "gbs": round(b/(acc[k]*1e-3)/1e9, 1) if acc[k] > 0 else None,
                                  "frac_of_hbm_peak": round(b/(acc[k]*1e-3)/1e9/hbm, 4) if (hbm and acc[k] > 0) else None}
                              for k, b in bytes_of.items()}
            out["phases_ms"] = {k: round(v, 5) for k, v in acc.items()}
        return out


def md_report(name, w, local, rank, world, dist, flush, steps, skin):
    """The same workload along a trajectory (positions move every step): tile list rebuilt every evaluation (the main
    line's mode, Reference-platform behaviour) against platform property ListSkin (list kept until an atom has moved
    skin/2; the exact predicate still decides the pair set every step) -- include/slicednb.h snb_set_list_skin."""
    out = {"trajectory": "ballistic, 0.001 nm per step and component (thermal speed of a 10 u particle at 300 K over 2 fs), reversed every 40 steps; "
                         "positions rewritten on the device before every step, outside the timed region",
           "list_skin_nm": skin}
    for key, props in (("rebuild_every_step", {}), ("list_reuse", {"ListSkin": str(skin)})):
        r = Runner(name, local, rank, world, dist, properties=props, workload=w)
        m = r.measure_md(steps, 8, flush)
        if key == "list_reuse":
            m["steps_per_rebuild"] = round(m["steps"]/max(1, m["rebuilds"]), 1)
        else:
            m.pop("rebuilds")
        out[key] = m
        del r
    out["speedup"] = out["list_reuse"]["value"]/out["rebuild_every_step"]["value"]
    return out


def short(res, keys=("value", "ms_per_step", "device_ms_per_step", "e2e", "roofline", "gpu_launches", "kernels", "phases_ms")):
    out = {k: res[k] for k in keys if k in res}
    if "roofline" in out:
        out["roofline"] = {k: out["roofline"][k] for k in ("kernel", "achieved", "frac", "kernel_ms", "interacting_pairs", "flop_per_pair", "list_fill", "traffic")}
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
    from openmmlab_b200 import systems
    name = args.workload or ("dhfr" if world == 1 else "stmv")
    natoms = systems.WORKLOAD_ATOMS.get(name, 0)
    partition = world > 1 and (args.multi == "partition" or (args.multi == "auto" and natoms >= PARTITION_MIN_ATOMS))
    big = natoms >= PARTITION_MIN_ATOMS
    steps, warmup = (min(args.steps, 30), min(args.warmup, 5)) if big else (args.steps, args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    main = Runner(name, local, rank, world, dist, partition=partition, properties={"DirectKernel": args.direct_kernel})
    res = main.measure(steps, warmup, flush, peaks)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    main_w = main.w
    del main
    line = None
    if rank == 0:
        value = res["value"]
        line = {"metric": "sliced-PME force evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world,
                "steps": steps, "warmup": max(warmup, 12) if partition else warmup, "ms_per_step": res["ms_per_step"],
                "device_ms_per_step": res["device_ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if partition else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(main_w, world, partition), "ns_per_day_2fs": value*0.1728, "ns_per_day_4fs": value*0.3456,
                "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "roofline": res["roofline"], "clocks": sampler.summary()}
    extras = {}
    if world == 1:
        if not args.no_parity:
            extras["parity"] = parity_report(main_w, local)
            # the mode that meets the north-star tolerance on every quantity, timed as well
            dbl = Runner(name, local, rank, world, dist, properties={"Precision": "double"}, workload=main_w)
            r = dbl.measure(min(steps, 50), min(warmup, 5), flush, peaks, e2e=False, detail=False)
            extras["value_double"] = {"value": r["value"], "unit": "evals/s", "ms_per_step": r["ms_per_step"], "dtype": "f64",
                                      "direct_kernel_ms": r["roofline"]["kernel_ms"]}
            del dbl
        if not args.no_cpu:
            extras["cpu_baseline"] = cpu_reference(main_w)
        if not args.no_md:
            extras["md"] = md_report(name, main_w, local, rank, world, dist, flush, min(steps, 200), args.list_skin)
        others = [] if args.no_others else [x for x in ("stmv", "apoa1", "fep") if x != name]
        for other in others:
            r = Runner(other, local, rank, world, dist)
            m = r.measure(min(args.steps, 30 if other == "stmv" else 100), 3, flush, peaks, detail=other == "stmv")
            sub = short(m)
            sub["config"] = {k: v for k, v in config_of(r.w, 1).items() if k in ("workload", "atoms", "subsets", "slices", "method", "pme_grid")}
            sub["ns_per_day_2fs"] = m["value"]*0.1728
            if other == "stmv" and not args.no_parity:
                sub["parity"] = parity_report(r.w, local)
            extras[other] = sub
            del r
    else:
        # on record beside the shared evaluation: independent DHFR-size replicas, one per GPU (no collective)
        if not args.no_others and name != "dhfr":
            r = Runner("dhfr", local, rank, world, dist, partition=False)
            m = r.measure(args.steps, args.warmup, flush, peaks, detail=False)
            extras["replicas_dhfr"] = dict(short(m), scaling="weak", parallelism=config_of(r.w, world, False)["parallelism"])
            del r
    if rank == 0:
        # short objects first, the bulky ones last (a truncated tail must not cost the headline sub-objects)
        for k in ("parity", "value_double", "cpu_baseline", "md", "stmv", "apoa1", "fep", "replicas_dhfr"):
            if k in extras:
                line[k] = extras[k]
        line["kernels"] = res.get("kernels")
        line["phases_ms"] = res.get("phases_ms")
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
    ap.add_argument("--workload", default=None, help="default: dhfr on one GPU, stmv on several")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-md", action="store_true", help="skip the trajectory sub-object (list reuse)")
    ap.add_argument("--list-skin", type=float, default=0.1, help="ListSkin (nm) of the trajectory sub-object")
    ap.add_argument("--no-others", action="store_true", help="skip the sub-objects of the other workloads")
    ap.add_argument("--no-stmv", action="store_true", help="(kept for old command lines) same as --no-others")
    ap.add_argument("--multi", default="auto", choices=["auto", "replicas", "partition"])
    ap.add_argument("--direct-kernel", default="auto", choices=["auto", "packed", "general"],
                    help="platform property DirectKernel of the main workload (A/B measurements)")
    args = ap.parse_args()
    args.no_others = args.no_others or args.no_stmv
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

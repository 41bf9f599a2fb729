#!/usr/bin/env python
"""bench.py — GLUPS of the fused D2Q9 collide-stream step (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--dtype f32|f64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of lbm.rs on the host cores

A "step" is one State::step (stream -> bounce-back -> collide) over the whole lattice.
Workloads:
  config2  (default) 4096x4096 per GPU, periodic, no solids, smooth analytic init, tau=0.8;
           N GPUs stack N such slabs in y (weak scaling)
  config3  8192x2048 channel with walls + cylinder mask (per GPU)
  strong   32768x32768 global lattice, y-slab sharded over the N GPUs (config 4)
  weak16k  16384x16384 per GPU (config 5)
Multi-GPU runs shard the lattice into y-slabs, one process per GPU; `--halo p2p` (default) uses the
fused peer-memory halo (falls back to NCCL when the neighbours cannot be mapped), `--halo nccl` the
NCCL send/recv exchange.

Timing (one JSON line, printed by rank 0).  The K-step batch is timed `reps` times; every batch
is bracketed by a host barrier + synchronize on both sides, starts behind a DEVICE-side barrier
(chemsim_lbm_barrier: a one-element all-reduce on the lattice's stream, so the ranks' start events
are aligned to a collective's latency instead of to host skew) and is timed with CUDA events on
the lattice's stream; a batch's time is the MAX over ranks, `ms_per_step`/`value` are the MEDIAN
over the batches (`reps`, `stat`, `batch_ms` say so).  Besides the headline the line carries
sub-records under `extras` (skip with --no-extras): BASELINE config 4 (`strong_32768`), config 5
(`weak16k`, with the mass drift against its prediction) and, for N > 1, `parity_sharded` — a
sharded run bit-compared with an unsharded one and with the CPU restatement.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_CELL = {"f32": 72, "f64": 144}      # 9 loads + 9 stores of one population value
NP_DTYPE = {"f32": np.float32, "f64": np.float64}
TAU = 0.8
N1_RECORD = "/tmp/chemsim_b200_bench_n1.json"  # single-GPU rates of this box, for the N>1 runs' efficiencies


def workload_shape(name: str, n_gpus: int):
    """-> (width, global_height, scaling)"""
    if name == "config2":
        return 4096, 4096 * n_gpus, "weak"
    if name == "config3":
        return 8192, 2048 * n_gpus, "weak"
    if name == "weak16k":
        return 16384, 16384 * n_gpus, "weak"
    if name == "strong":
        return 32768, 32768, "strong"
    raise SystemExit(f"unknown workload {name}")


def workload_config(name: str, n_gpus: int, dtype: str, collision: str):
    """The `config` object — identical in the GPU arm and the --impl reference arm."""
    w, hg, _ = workload_shape(name, n_gpus)
    hl = hg // n_gpus
    return {"workload": name, "lattice": f"{w}x{hg}", "per_gpu": f"{w}x{hl}",
            "collision": "BGK tau=0.8" if collision == "bgk" else f"{collision} nu=0.1",
            "edge": "periodic", "sharding": f"y-slabs x{n_gpus}",
            "l2": "working set %.2f GiB per GPU >> 126 MB L2 (no flush needed)"
                  % (2 * w * hl * BYTES_PER_CELL[dtype] / 2 / 2**30)}


def slab_fields(name: str, w: int, hg: int, y0: int, y1: int, dtype):
    from chemsim_b200 import scenarios
    if name == "config3":
        # channel + cylinder, replicated per 2048-row slab so every GPU has the same mask work
        rho, vx, vy, solid = scenarios.channel_cylinder(w, 2048, dtype)
        reps = (y1 - y0) // 2048
        tile = lambda a: np.tile(a, (reps, 1))
        return tile(rho), tile(vx), tile(vy), tile(solid)
    return scenarios.smooth_periodic_rows(w, hg, y0, y1, dtype)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_bytes(workload: str, dtype: str):
    """dram read+write bytes per launch of the step kernel from the committed ncu summary."""
    path = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(f"{workload}_{dtype}")
    except Exception:
        return None


def f32_weight_excess():
    """sum of the reference's f32 weights 16/36, 4/36 x4, 1/36 x4 minus 1 (src/lbm.rs:209-219):
    BGK creates this much relative mass per step, divided by tau."""
    w32 = [np.float32(n) / np.float32(36.0) for n in (16, 4, 4, 4, 4, 1, 1, 1, 1)]
    return float(np.sum(np.array(w32, dtype=np.float64))) - 1.0


def predicted_mass_drift(dtype: str, steps: int, collision: str):
    if collision != "bgk":
        return None
    return (f32_weight_excess() if dtype == "f32" else 0.0) / TAU * steps


class ClockSampler:
    """Samples SM clock + throttle reasons with NVML while the timed region runs.  NVML is
    initialised in the constructor — construct it BEFORE the barrier that precedes the timed
    region (several processes initialising NVML at once take milliseconds, rank-dependent)."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.005):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period_s = period_s
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if visible:
                index = int(visible.split(",")[index])
            self._dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._dev, pynvml.NVML_CLOCK_SM)
            self.sample_now()          # first call pays NVML's lazy set-up; not part of the record
            self.samples.clear()
            self.reasons.clear()
        except Exception:
            self._nv = None

    def sample_now(self):
        nv = self._nv
        if not nv:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._dev, nv.NVML_CLOCK_SM))
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._dev)
            for bit, name in self.REASONS.items():
                if mask & bit and name != "gpu_idle":
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self.sample_now()
            self._stop.wait(self.period_s)

    def __enter__(self):
        if self._nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------
# CPU arm: the oracle's fused OpenMP restatement of lbm.rs, timed on the host cores
# ------------------------------------------------------------------------------------

def cpu_stepper(w, h, dtype):
    """Steady-state CPU stepping: persistent A/B buffers, first-touched by the OpenMP threads,
    nothing allocated or copied inside the timed calls (oracle.FusedStepper)."""
    from chemsim_b200 import scenarios
    from oracle import lbm_oracle as O
    rho, vx, vy, _ = scenarios.smooth_periodic(w, h, dtype)
    f = O.compute_equilibrium(rho, vx, vy)
    del rho, vx, vy
    return O.FusedStepper(f, None, TAU, O.EDGE_PERIODIC)


def cpu_time_steps(stepper, steps, warmup):
    if warmup:
        stepper.step(warmup)
    t0 = time.perf_counter()
    f = stepper.step(steps)
    dt = time.perf_counter() - t0
    assert np.isfinite(f[0, 0, 0])
    return dt


def cpu_baseline(w, dtype_name, budget_s=20.0):
    """Bounded sample of the workload on the host cores -> dict for the JSON line."""
    from oracle import lbm_oracle as O
    O.use_all_cores()
    dtype = NP_DTYPE[dtype_name]
    rows = 1024                                    # a (w x 1024) band of the lattice, periodic
    stepper = cpu_stepper(w, rows, dtype)
    t5 = cpu_time_steps(stepper, 5, 3) / 5.0
    steps = int(max(5, min(20000, budget_s / max(t5, 1e-4))))
    dt = cpu_time_steps(stepper, steps, 0)
    glups = w * rows * steps / dt / 1e9
    out = {"value": glups, "unit": "GLUPS", "cores": O.max_threads(), "kind": "port",
           "sample": f"{w}x{rows} band of the workload, periodic, {steps} steps, fused OpenMP restatement of lbm.rs "
                     f"(oracle/lbm_oracle.c, not ArrayFire), persistent buffers, {dt:.2f} s"}
    try:
        out["reference_structured_1thread"] = cpu_reference_structured(dtype)
    except Exception as e:                      # informational only
        out["reference_structured_1thread"] = {"error": str(e)}
    return out


def cpu_reference_structured(dtype, size=1024, steps=3):
    """The oracle in the reference's own structure: three separate passes
    (stream / bounce-back / collide) with full-array temporaries, single thread, mirroring
    lbm.rs's array-at-a-time calls.  Informational: what the restatement costs before fusing."""
    from chemsim_b200 import scenarios
    from oracle import lbm_oracle as O
    rho, vx, vy, solid = scenarios.smooth_periodic(size, size, dtype)
    f = O.compute_equilibrium(rho, vx, vy)
    col = O.collision(O.BGK, tau=TAU)
    f = O.step_ref(f, solid, 1, col, O.EDGE_PERIODIC)
    t0 = time.perf_counter()
    f = O.step_ref(f, solid, steps, col, O.EDGE_PERIODIC)
    dt = time.perf_counter() - t0
    return {"value": size * size * steps / dt / 1e9, "unit": "GLUPS", "cores": 1,
            "sample": f"{size}x{size}, {steps} steps, three-pass array-at-a-time restatement"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import lbm_oracle as O
    O.use_all_cores()
    w, hg, scaling = workload_shape(args.workload, args.gpus)
    dtype = NP_DTYPE[args.dtype]
    # size each step so that the whole run fits in ~2 minutes of CPU time
    probe_rows = 256
    probe = cpu_stepper(w, probe_rows, dtype)
    per_row = cpu_time_steps(probe, 3, 1) / 3.0 / probe_rows
    del probe
    budget_per_step = 100.0 / (args.steps + args.warmup + 2)
    rows = int(min(hg, max(64, (budget_per_step / per_row) // 64 * 64)))
    stepper = cpu_stepper(w, rows, dtype)          # includes two untimed steps (first touch of both buffers)
    dt = cpu_time_steps(stepper, args.steps, args.warmup)
    glups = w * rows * args.steps / dt / 1e9
    sample = (f"each step updates a {w}x{rows} periodic band of the {w}x{hg} lattice, persistent buffers "
              f"(fused OpenMP restatement of lbm.rs, oracle/lbm_oracle.c, {O.max_threads()} threads)")
    line = {
        "impl": "reference", "metric": "GLUPS D2Q9 fused collide-stream", "value": glups, "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, args.dtype, "bgk"),
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": O.max_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------

class Ctx:
    """Process-group plumbing shared by the headline leg and the extras."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run")
        if not torch.cuda.is_available():
            raise SystemExit("no CUDA device: the D2Q9 path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        """element-wise MAX over the ranks of a list of floats"""
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu().tolist()]

    def nccl_id(self):
        """a fresh ncclUniqueId for one sharded lattice, from rank 0"""
        from chemsim_b200 import lbm
        if self.world == 1:
            return None
        ident = self.torch.zeros(128, dtype=self.torch.uint8, device=self.dev)
        if self.rank == 0:
            ident = self.torch.frombuffer(bytearray(lbm.nccl_unique_id()), dtype=self.torch.uint8).to(self.dev)
        self.dist.broadcast(ident, 0)
        return bytes(ident.cpu().numpy().tobytes())

    def make_state(self, w, hg, dtype, collision, halo, edge=None):
        from chemsim_b200 import lbm
        state = lbm.State.create((w, hg), collision, lbm.Discretization(1.0, 1.0), dtype=dtype,
                                 edge=lbm.EDGE_PERIODIC if edge is None else edge, device=self.local_rank,
                                 rank=self.rank, nranks=self.world, nccl_id=self.nccl_id())
        if halo == "p2p" and self.world > 1:
            state.enable_p2p_halo()
        return state

    def init_workload(self, state, workload, w, hg, dtype):
        """initialise in row chunks so that host memory stays bounded for the 32768^2 / 16384^2 workloads"""
        hl, y0 = state.local_height, state.row_offset
        chunk = 2048
        for r in range(0, hl, chunk):
            rows = min(chunk, hl - r)
            rho, vx, vy, solid = slab_fields(workload, w, hg, y0 + r, y0 + r + rows, dtype)
            state.init_equilibrium_rows(r, rho, vx, vy)
            if solid.any():
                state.set_geometry_rows(r, solid)

    def timed_batches(self, state, steps, reps):
        """`reps` batches of exactly `steps` steps -> per-batch device time in ms, MAX over ranks."""
        torch = self.torch
        stream = torch.cuda.ExternalStream(state.cuda_stream(), device=self.dev)
        events = []
        for _ in range(reps):
            self.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            state.barrier()                  # device-side: every rank's ev0 follows the same collective
            ev0.record(stream)
            state.step(steps)
            ev1.record(stream)
            state.synchronize()
            self.barrier()
            events.append((ev0, ev1))
        return self.max_over_ranks([a.elapsed_time(b) for a, b in events])


def pick_reps(ctx, args, batch_ms):
    if args.reps > 0:
        return args.reps
    # about one second of timed stepping, 5..40 batches; every rank must pick the same number
    est = ctx.max_over_ranks([batch_ms])[0]
    return int(min(40, max(5, round(1000.0 / max(est, 0.05)))))


def collision_of(name, dtype):
    from chemsim_b200 import lbm
    disc = lbm.Discretization(1.0, 1.0)
    return {"bgk": lambda: lbm.BGK(TAU),                                   # nu = (tau - 1/2)/3 = 0.1
            "trt": lambda: lbm.TRT.new(0.25, 0.1, disc, dtype),           # same viscosity, magic lambda 1/4
            "regularized": lambda: lbm.Regularized.new(lbm.KBC.new(0.1)),  # main.rs:198-199's operator
            "kbc": lambda: lbm.KBC.new(0.1)}[name]()


def run_gpu_arm(args):
    ctx = Ctx(args)
    torch = ctx.torch
    world, rank = ctx.world, ctx.rank
    clocks = ClockSampler(ctx.local_rank)          # NVML set-up happens here, far from any timed region
    dtype = NP_DTYPE[args.dtype]
    w, hg, scaling = workload_shape(args.workload, world)
    state = ctx.make_state(w, hg, dtype, collision_of(args.collision, dtype), args.halo)
    hl = state.local_height
    ctx.init_workload(state, args.workload, w, hg, dtype)
    mass0 = state.total_mass(global_=True)
    cells_global = w * hg

    # ---- device-resident throughput (`value`) -------------------------------------
    state.step(args.warmup)
    state.synchronize()
    calib = ctx.timed_batches(state, args.steps, 1)[0]       # one more untimed batch: calibrates `reps`
    reps = pick_reps(ctx, args, calib)
    launches0 = state.kernel_launches()
    with clocks:
        batch_ms = ctx.timed_batches(state, args.steps, reps)
    # kernels of ONE timed K-step batch (minus the start barrier's all-reduce, which precedes ev0)
    launches = (state.kernel_launches() - launches0) // reps - (1 if world > 1 else 0)
    ms = float(np.median(batch_ms))
    glups = cells_global * args.steps / (ms * 1e-3) / 1e9
    steps_taken = args.warmup + args.steps * (reps + 1)
    mass1 = state.total_mass(global_=True)

    # ---- end to end through the C ABI with host buffers (`e2e`) --------------------
    # One frame of the reference's loop per step (main.rs:66-91, :128-177): upload the
    # (possibly edited) geometry, State::step, read the density field back for rendering.
    big = w * hl > 2 ** 28
    e2e_steps = 4 if big else max(4, min(args.steps, 30))
    mask_host = torch.from_numpy(state.geometry.astype(np.uint8)).pin_memory()
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    rho_host = [torch.empty((hl, w), dtype=tdt).pin_memory() for _ in range(2)]   # the caller double-buffers

    def frame_upload(i, nsteps=1):
        state.set_geometry_async(mask_host.data_ptr(), mask_host.numel())
        state.step(nsteps)
        state.density_async(rho_host[i & 1].data_ptr(), rho_host[i & 1].numel())

    def frame_paint(i, nsteps=2):
        state.paint_brush((w // 2 + (i % 7), hl // 2))        # main.rs:71-91 on the device: no H2D at all
        state.step(nsteps)
        state.density_async(rho_host[i & 1].data_ptr(), rho_host[i & 1].numel())

    def time_frames(frame, frames, **kw):
        for i in range(2):
            frame(i, **kw)
        state.synchronize()
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(frames):
            frame(i, **kw)
        state.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ctx.barrier()
        return ctx.max_over_ranks([wall_ms])[0]

    e2e_ms = time_frames(frame_upload, e2e_steps)
    e2e_glups = cells_global * e2e_steps / (e2e_ms * 1e-3) / 1e9
    assert abs(float(rho_host[0].mean()) - 1.0) < 0.05 and abs(float(rho_host[1].mean()) - 1.0) < 0.05
    d2h_bytes = int(rho_host[0].numel() * rho_host[0].element_size()) * world
    e2e = {"value": e2e_glups, "unit": "GLUPS", "h2d_bytes_per_step": int(mask_host.numel()) * world,
           "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "d2h_gb_per_s_per_gpu": d2h_bytes / world / (e2e_ms / e2e_steps * 1e-3) / 1e9,
           "what": "one frame of the reference's loop per STEP through the C ABI with pinned HOST buffers: "
                   "chemsim_lbm_set_geometry_async(mask) + chemsim_lbm_step(1) + chemsim_lbm_get_density_async(rho); "
                   "host wall clock incl. the final synchronize.  PCIe-bound by the density read-back (see "
                   "d2h_gb_per_s_per_gpu); a floor: main.rs steps speed_factor = 2 times per frame (main.rs:324)"}
    if not args.no_extras:
        ms2 = time_frames(frame_upload, e2e_steps, nsteps=2)
        ms3 = time_frames(frame_paint, e2e_steps, nsteps=2)
        e2e["variants"] = {
            "speed_factor_2": {"value": cells_global * 2 * e2e_steps / (ms2 * 1e-3) / 1e9, "unit": "GLUPS",
                               "what": "same frame with main.rs's two steps per frame"},
            "paint_rect_speed_factor_2": {"value": cells_global * 2 * e2e_steps / (ms3 * 1e-3) / 1e9, "unit": "GLUPS",
                                          "h2d_bytes_per_step": 0,
                                          "what": "the mouse handler's 9x9 brush painted on the device "
                                                  "(chemsim_lbm_fill_geometry + chemsim_lbm_paint_rect) instead of "
                                                  "a mask upload, two steps, density read-back"}}
        state.fill_geometry(False)
    kernel_name, halo_mode = state.step_kernel_name(), (state.halo_mode() if world > 1 else "none")
    state.close()
    del state, rho_host, mask_host

    def emit(extras):
        if rank != 0:
            return
        peak, peak_src = measured_peak_gbs()
        bpc = BYTES_PER_CELL[args.dtype]
        # The dominant kernel is the fused step kernel.  One launch reads every population once and
        # writes it once = 72 B (f32) / 144 B (f64) per cell of ALGORITHMIC traffic — whether the
        # launch advances the lattice by one step (step_vec_kernel) or by two (step2_kernel, which
        # keeps the intermediate lattice in shared memory).  achieved = that / the launch duration.
        cells_per_launch = w * hl
        # passes over the slab per batch: the library pairs the steps of one chemsim_lbm_step call
        step_launches = (args.steps // 2 + args.steps % 2) if kernel_name.startswith("step2") else args.steps
        steps_per_launch = args.steps / step_launches
        launch_s = ms * 1e-3 / step_launches
        achieved = bpc * cells_per_launch / launch_s / 1e9
        drift = (mass1 - mass0) / mass0
        line = {
            "metric": "GLUPS D2Q9 fused collide-stream", "value": glups, "unit": "GLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args.workload, world, args.dtype, args.collision),
            "reps": reps, "stat": "median over reps of (max over ranks of the CUDA-event time of one K-step batch)",
            "batch_ms": {"min": min(batch_ms), "median": ms, "max": max(batch_ms)},
            "run": {"halo": halo_mode, "kernel": kernel_name, "steps_taken": steps_taken,
                    "mass_drift_rel": drift,
                    "mass_drift_predicted_rel": predicted_mass_drift(args.dtype, steps_taken, args.collision),
                    "start_alignment": "chemsim_lbm_barrier (1-element ncclAllReduce on the lattice stream) before ev0"
                                       if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(args.workload, args.dtype),
                         "algorithmic_bytes_per_launch": bpc * cells_per_launch, "peak_source": peak_src,
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "steps_per_launch": steps_per_launch, "launch_us": launch_s * 1e6,
                         "per_step_accounting": {
                             "bytes_per_cell_per_step": bpc, "GBps": bpc * cells_per_launch * steps_per_launch / launch_s / 1e9,
                             "frac_of_peak": bpc * cells_per_launch * steps_per_launch / launch_s / 1e9 / peak,
                             "note": "BASELINE.json's 72 B (f32) / 144 B (f64) per lattice update, i.e. what a "
                                     "one-step-per-pass kernel would have to move; above 1.0 because the two-step "
                                     "kernel moves half of it (temporal blocking), not because work is skipped: "
                                     "results are bit-identical to single steps (tests/, extras.parity_sharded)"}},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "gpu_launches_what": "kernels of this rank inside one timed batch of `steps` steps",
            "clocks": clocks.summary(),
        }
        if extras is not None:
            line["extras"] = extras
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(w, args.dtype)
        if world == 1 and args.workload == "config2" and args.collision == "bgk":
            remember_n1(args.dtype, "config2", glups)
        print(json.dumps(line), flush=True)

    extras = None
    if not args.no_extras:
        # The sub-records must never cost the headline: if a leg fails or stalls (a rank that left a
        # collective cannot be waited for), the line is printed with what exists and the process ends.
        partial = {}

        def bail(reason):
            partial["aborted"] = reason
            emit(partial)
            sys.stdout.flush()
            os._exit(0)

        watchdog = threading.Timer(args.extras_deadline, bail, args=(f"extras exceeded {args.extras_deadline:.0f} s",))
        watchdog.daemon = True
        watchdog.start()
        try:
            run_extras(ctx, args, partial)
        except Exception as e:
            bail(f"{type(e).__name__}: {e}")
        watchdog.cancel()
        extras = partial
    emit(extras)
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


# ------------------------------------------------------------------------------------
# extras: BASELINE configs 4 and 5 and the sharded parity check, as sub-records
# ------------------------------------------------------------------------------------

def remember_n1(dtype, key, glups):
    """The N=1 run leaves its rates for the N>1 runs of the same box (efficiencies)."""
    try:
        rec = {}
        if os.path.exists(N1_RECORD):
            with open(N1_RECORD) as fh:
                rec = json.load(fh)
        rec[f"{key}_{dtype}"] = glups
        with open(N1_RECORD, "w") as fh:
            json.dump(rec, fh)
    except Exception:
        pass


def recall_n1(dtype, key):
    try:
        with open(N1_RECORD) as fh:
            return json.load(fh).get(f"{key}_{dtype}")
    except Exception:
        return None


def throughput_record(ctx, args, workload, steps, reps, warmup=3, mass_check=False):
    """GLUPS of a BASELINE workload on the current world, BGK, `reps` batches of `steps` steps."""
    from chemsim_b200 import lbm
    dtype = NP_DTYPE[args.dtype]
    world = ctx.world
    w, hg, scaling = workload_shape(workload, world)
    state = ctx.make_state(w, hg, dtype, lbm.BGK(TAU), args.halo)
    ctx.init_workload(state, workload, w, hg, dtype)
    mass0 = state.total_mass(global_=True) if mass_check else None
    state.step(warmup)
    state.synchronize()
    batch_ms = ctx.timed_batches(state, steps, reps)
    ms = float(np.median(batch_ms))
    glups = w * hg * steps / (ms * 1e-3) / 1e9
    rec = {"workload": workload, "lattice": f"{w}x{hg}", "per_gpu": f"{w}x{state.local_height}", "scaling": scaling,
           "n_gpus": world, "steps": steps, "reps": reps, "value": glups, "unit": "GLUPS", "ms_per_step": ms / steps,
           "batch_ms": {"min": min(batch_ms), "median": ms, "max": max(batch_ms)},
           "halo": state.halo_mode() if world > 1 else "none", "kernel": state.step_kernel_name()}
    peak, _ = measured_peak_gbs()
    rec["roofline_frac_per_gpu"] = BYTES_PER_CELL[args.dtype] * w * state.local_height / (ms * 1e-3 / steps) / 1e9 / peak
    if mass_check:
        taken = warmup + steps * reps
        mass1 = state.total_mass(global_=True)
        rec["steps_taken"] = taken
        rec["mass_drift_rel"] = (mass1 - mass0) / mass0
        rec["mass_drift_predicted_rel"] = predicted_mass_drift(args.dtype, taken, "bgk")
        rec["mass_drift_note"] = ("f32: the reference's weights sum to 1 + 7.45e-9, BGK creates (sum w - 1)/tau of "
                                  "relative mass per step (DESIGN.md §1); the residual against that prediction is "
                                  "the conservation check")
        if rec["mass_drift_predicted_rel"] is not None:
            rec["mass_drift_residual_rel"] = rec["mass_drift_rel"] - rec["mass_drift_predicted_rel"]
    state.close()
    n1 = glups if world == 1 else recall_n1(args.dtype, workload)
    if world == 1:
        remember_n1(args.dtype, workload, glups)
    if n1:
        rec["n1_value"] = n1
        rec["efficiency"] = glups / (n1 * world)          # strong and weak alike: GLUPS_N / (N * GLUPS_1)
        rec["efficiency_basis"] = "same-box N=1 run of this bench (%s)" % N1_RECORD
    return rec


def parity_sharded(ctx, args):
    """A 4096 x (64*N) lattice with random solids, 12 steps: the sharded run (p2p and nccl halo) is
    gathered on rank 0 and compared BIT FOR BIT with an unsharded run on rank 0's GPU and with the
    CPU restatement (oracle/ — used here only as the checker)."""
    from chemsim_b200 import lbm, scenarios
    dist = ctx.dist
    dtype = NP_DTYPE[args.dtype]
    w, hg, steps = 4096, 64 * ctx.world, 12
    rho, vx, vy, solid = scenarios.random_state(w, hg, dtype, seed=23, solid_fraction=0.02)
    out = {"lattice": f"{w}x{hg}", "steps": steps, "collision": "BGK tau=0.8", "edge": "periodic"}
    ref_single = ref_oracle = None
    u = np.uint32 if dtype == np.float32 else np.uint64
    if ctx.rank == 0:
        single = lbm.State.create((w, hg), lbm.BGK(TAU), dtype=dtype, edge=lbm.EDGE_PERIODIC, device=ctx.local_rank)
        single.init_equilibrium(rho, vx, vy)
        single.geometry = solid
        single.step(steps)
        ref_single = single.populations_array().view(u)
        single.close()
        try:
            from oracle import lbm_oracle as O
            O.use_all_cores()
            ref_oracle = O.step_fused(O.compute_equilibrium(rho, vx, vy), solid, steps, TAU, O.EDGE_PERIODIC).view(u)
        except Exception as e:          # the checker is optional here; tests/ hold the gate
            out["oracle_error"] = str(e)
    for halo in ("p2p", "nccl"):
        state = ctx.make_state(w, hg, dtype, lbm.BGK(TAU), halo)
        r0, h = state.row_offset, state.local_height
        state.init_equilibrium(rho[r0:r0 + h], vx[r0:r0 + h], vy[r0:r0 + h])
        state.geometry = solid[r0:r0 + h]
        for n in (1, 2, steps - 3):      # single steps and batches: every dependency edge of the pipeline
            state.step(n)
        state.synchronize()
        mine = state.populations_array()
        mode = state.halo_mode()
        state.close()
        gathered = [None] * ctx.world
        dist.gather_object(mine, gathered if ctx.rank == 0 else None, dst=0)
        if ctx.rank == 0:
            got = np.concatenate(gathered, axis=1).view(u)
            rec = {"halo": mode, "bit_identical_to_unsharded": bool((got == ref_single).all())}
            if ref_oracle is not None:
                rec["bit_identical_to_cpu_restatement"] = bool((got == ref_oracle).all())
            out[halo] = rec
    return out


def run_extras(ctx, args, extras):
    legs = [("strong_32768", lambda: throughput_record(ctx, args, "strong", 50, 3)),
            ("weak16k", lambda: throughput_record(ctx, args, "weak16k", args.extras_weak_steps, 1, mass_check=True))]
    if ctx.world > 1:
        legs.append(("parity_sharded", lambda: parity_sharded(ctx, args)))
    for name, leg in legs:
        t0 = time.perf_counter()
        rec = leg()
        rec["wall_s"] = time.perf_counter() - t0
        extras[name] = rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--reps", type=int, default=0, help="timed batches of --steps steps (default: ~1 s worth, 5..40)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "strong", "weak16k"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records (strong_32768, weak16k, parity_sharded, e2e variants)")
    ap.add_argument("--extras-weak-steps", type=int, default=1000, help="steps of the weak16k sub-record")
    ap.add_argument("--extras-deadline", type=float, default=300.0,
                    help="seconds after which the sub-records are abandoned and the line is printed without them")
    ap.add_argument("--collision", default="bgk", choices=["bgk", "trt", "regularized", "kbc"],
                    help="collision operator of the step (default: BGK, BASELINE.json's metric)")
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p"],
                    help="multi-GPU halo: the fused peer-memory face kernel (default; falls back to NCCL if the "
                         "neighbours cannot be mapped) or NCCL send/recv")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload != "config2" or args.collision != "bgk":
        args.no_extras = True             # the sub-records belong to the headline configuration
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

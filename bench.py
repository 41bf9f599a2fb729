#!/usr/bin/env python
"""bench.py — GLUPS of the fused D2Q9 collide-stream step (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--dtype f32|f64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the CPU restatement of lbm.rs on the host cores

A "step" is one State::step (stream -> bounce-back -> BGK collide) over the whole
lattice.  Workloads:
  config2  (default) 4096x4096 per GPU, periodic, no solids, smooth analytic init, tau=0.8;
           N GPUs stack N such slabs in y (weak scaling)
  config3  8192x2048 channel with walls + cylinder mask (per GPU)
  strong   32768x32768 global lattice, y-slab sharded over the N GPUs (config 4)
  weak16k  16384x16384 per GPU (config 5)
Multi-GPU runs shard the lattice into y-slabs, one process per GPU; `--halo p2p` (default) uses the
fused peer-memory halo (falls back to NCCL when the neighbours cannot be mapped), `--halo nccl` the
NCCL send/recv exchange.  One JSON line is printed by rank 0.
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


class ClockSampler:
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
        except Exception:
            self._nv = None

    def sample_now(self):
        """One sample, taken by the caller while the GPU is busy (guarantees at least one
        sample under load even when the timed region is shorter than the sampling period)."""
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
            self._stop.wait(0.02)

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

def cpu_time_steps(w, h, dtype, steps, warmup):
    from chemsim_b200 import scenarios
    from oracle import lbm_oracle as O
    rho, vx, vy, _ = scenarios.smooth_periodic(w, h, dtype)
    f = O.compute_equilibrium(rho, vx, vy)
    if warmup:
        f = O.step_fused(f, None, warmup, TAU, O.EDGE_PERIODIC)
    t0 = time.perf_counter()
    f = O.step_fused(f, None, steps, TAU, O.EDGE_PERIODIC)
    dt = time.perf_counter() - t0
    assert np.isfinite(f[0, 0, 0])
    return dt


def cpu_baseline(w, dtype_name, budget_s=25.0):
    """Bounded sample of the workload on the host cores -> dict for the JSON line."""
    from oracle import lbm_oracle as O
    O.use_all_cores()
    dtype = NP_DTYPE[dtype_name]
    rows = 1024                                    # a (w x 1024) band of the lattice, periodic
    t5 = cpu_time_steps(w, rows, dtype, 5, 3) / 5.0            # calibrate after the pages are touched
    steps = int(max(5, min(20000, budget_s / max(t5, 1e-4))))
    dt = cpu_time_steps(w, rows, dtype, steps, 3)
    glups = w * rows * steps / dt / 1e9
    out = {"value": glups, "unit": "GLUPS", "cores": O.max_threads(), "kind": "port",
           "sample": f"{w}x{rows} band of the workload, periodic, {steps} steps, fused OpenMP restatement of lbm.rs "
                     f"(oracle/lbm_oracle.c, not ArrayFire), {dt:.2f} s"}
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
    t_probe = cpu_time_steps(w, probe_rows, dtype, 1, 1)
    per_row = t_probe / probe_rows
    budget_per_step = 120.0 / (args.steps + args.warmup)
    rows = int(min(hg, max(64, (budget_per_step / per_row) // 64 * 64)))
    dt = cpu_time_steps(w, rows, dtype, args.steps, args.warmup)
    glups = w * rows * args.steps / dt / 1e9
    sample = (f"each step updates a {w}x{rows} periodic band of the {w}x{hg} lattice "
              f"(fused OpenMP restatement of lbm.rs, {O.max_threads()} threads)")
    line = {
        "impl": "reference", "metric": "GLUPS D2Q9 fused collide-stream", "value": glups, "unit": "GLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": args.workload, "lattice": f"{w}x{hg}", "collision": "BGK tau=0.8",
                   "edge": "periodic", "sample": sample},
        "cpu_baseline": {"value": glups, "unit": "GLUPS", "cores": O.max_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": glups, "unit": "GLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from chemsim_b200 import lbm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the D2Q9 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    nccl_id = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        ident = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            ident = torch.frombuffer(bytearray(lbm.nccl_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(ident, 0)
        nccl_id = bytes(ident.cpu().numpy().tobytes())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dtype = NP_DTYPE[args.dtype]
    w, hg, scaling = workload_shape(args.workload, world)
    disc = lbm.Discretization(1.0, 1.0)
    collision = {"bgk": lambda: lbm.BGK(TAU),                                   # nu = (tau - 1/2)/3 = 0.1
                 "trt": lambda: lbm.TRT.new(0.25, 0.1, disc, dtype),           # same viscosity, magic lambda 1/4
                 "regularized": lambda: lbm.Regularized.new(lbm.KBC.new(0.1)),  # main.rs:198-199's operator
                 "kbc": lambda: lbm.KBC.new(0.1)}[args.collision]()
    state = lbm.State.create((w, hg), collision, disc, dtype=dtype,
                             edge=lbm.EDGE_PERIODIC, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    if args.halo == "p2p" and world > 1:
        state.enable_p2p_halo()
    hl, y0 = state.local_height, state.row_offset
    # initialise in row chunks so that host memory stays bounded for the 32768^2 / 16384^2 workloads
    chunk = 2048
    solid_any = False
    for r in range(0, hl, chunk):
        rows = min(chunk, hl - r)
        rho, vx, vy, solid = slab_fields(args.workload, w, hg, y0 + r, y0 + r + rows, dtype)
        state.init_equilibrium_rows(r, rho, vx, vy)
        if solid.any():
            state.set_geometry_rows(r, solid)
            solid_any = True
    del rho, vx, vy, solid
    mass0 = state.total_mass(global_=True)
    stream = torch.cuda.ExternalStream(state.cuda_stream(), device=dev)
    cells_global = w * hg

    # ---- device-resident throughput (`value`) -------------------------------------
    state.step(args.warmup)
    state.synchronize()
    barrier()
    launches0 = state.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        state.step(args.steps)
        ev1.record(stream)
        clocks.sample_now()              # the steps are queued and running: a sample under load
        state.synchronize()
        barrier()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = state.kernel_launches() - launches0
    glups = cells_global * args.steps / (ms * 1e-3) / 1e9
    mass1 = state.total_mass(global_=True)

    # ---- end to end through the C ABI with host buffers (`e2e`) --------------------
    # One frame of the reference's loop per step (main.rs:66-91, :128-177): upload the
    # (possibly edited) geometry, State::step, read the density field back for rendering.
    big = w * hl > 2 ** 28
    e2e_steps = 4 if big else max(4, min(args.steps, 30))
    mask_host = torch.from_numpy(state.geometry.astype(np.uint8)).pin_memory()
    tdt = torch.float32 if args.dtype == "f32" else torch.float64
    rho_host = [torch.empty((hl, w), dtype=tdt).pin_memory() for _ in range(2)]   # the caller double-buffers

    def frame(i):
        state.set_geometry_async(mask_host.data_ptr(), mask_host.numel())
        state.step(1)
        state.density_async(rho_host[i & 1].data_ptr(), rho_host[i & 1].numel())

    for i in range(2):
        frame(i)
    state.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        frame(i)
    state.synchronize()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_wall_ms)
    e2e_glups = cells_global * e2e_steps / (e2e_ms * 1e-3) / 1e9
    assert abs(float(rho_host[0].mean()) - 1.0) < 0.05 and abs(float(rho_host[1].mean()) - 1.0) < 0.05

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bpc = BYTES_PER_CELL[args.dtype]
        # the dominant kernel is the fused step kernel: one launch per step per GPU
        # (sharded runs add two one-row launches and one NCCL send/recv kernel per step)
        cells_per_launch = w * hl
        launch_s = ms * 1e-3 / args.steps
        achieved = bpc * cells_per_launch / launch_s / 1e9
        line = {
            "metric": "GLUPS D2Q9 fused collide-stream", "value": glups, "unit": "GLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": args.workload, "lattice": f"{w}x{hg}", "per_gpu": f"{w}x{hl}",
                       "collision": "BGK tau=0.8" if args.collision == "bgk" else f"{args.collision} nu=0.1",
                       "edge": "periodic", "sharding": f"y-slabs x{world}",
                       "halo": state.halo_mode() if world > 1 else "none",
                       "l2": "working set %.2f GiB per GPU >> 126 MB L2 (no flush needed)" % (2 * 9 * w * hl * (bpc / 18) / 2**30),
                       "kernel": state.step_kernel_name(),
                       "mass_drift_rel": abs(mass1 - mass0) / mass0},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic_bytes(args.workload, args.dtype),
                         "algorithmic_bytes_per_launch": bpc * cells_per_launch, "peak_source": peak_src,
                         "frac_of_nominal_8TBps": achieved / 8000.0},
            "e2e": {"value": e2e_glups, "unit": "GLUPS", "h2d_bytes_per_step": int(mask_host.numel()) * world,
                    "d2h_bytes_per_step": int(rho_host[0].numel() * rho_host[0].element_size()) * world,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "what": "one frame of the reference's loop per step through the C ABI with pinned HOST buffers: "
                            "chemsim_lbm_set_geometry_async(mask) + chemsim_lbm_step(1) + "
                            "chemsim_lbm_get_density_async(rho); host wall clock incl. the final synchronize; "
                            "PCIe-bound (D2H of the density field)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(w, args.dtype)
        print(json.dumps(line), flush=True)
    state.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=["config2", "config3", "strong", "weak16k"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--collision", default="bgk", choices=["bgk", "trt", "regularized", "kbc"],
                    help="collision operator of the step (default: BGK, BASELINE.json's metric)")
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p"],
                    help="multi-GPU halo: the fused peer-memory face kernel (default; falls back to NCCL if the "
                         "neighbours cannot be mapped) or NCCL send/recv")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

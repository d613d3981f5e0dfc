#!/usr/bin/env python
"""Benchmark of the propagation hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): the configuration the north-star target is quoted on — a batch
of independent 1024 x 1024 complex64 beams, one stateful ``Beam2D.propagate(dz)`` per step
(spectrum resident, spectrum *= H, field = ifft2(spectrum)), per-GPU batch fixed (weak
scaling), beams sharded across ranks with no collective on the data path.

  value      beam-propagations per second, inputs resident in HBM, whole job (all ranks)
  e2e        the same count through the public ensemble API with HOST (pinned) buffers:
             every step copies its input fields H2D, runs fft2 -> H -> ifft2 -> |E|^2 and
             copies the intensities D2H, copies inside the timed region
  roofline   dominant kernel timed alone with CUDA events on the launching stream
  cpu_baseline / --impl reference
             the numpy oracle port of the reference path on the box's host cores
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GRID = 1024
BITS = 64
BATCH_PER_GPU = 256          # 2 GiB spectrum + 2 GiB field per GPU: far larger than the 126 MB L2
E2E_BEAMS = 64               # beams per end-to-end step (512 MiB H2D + 256 MiB D2H)
AREA, WL, DZ = 4e-1, 532e-7, 0.02     # tests/gaussbeam_performance.py:16-23 geometry
METRIC = "propagate steps/sec at 1024^2 grid (complex64), batched Beam2D.propagate"
UNIT = "propagations/s"
WORKLOAD = (f"batched {N_GRID}x{N_GRID} complex64 stateful Beam2D.propagate(dz), "
            f"{BATCH_PER_GPU} beams per GPU, area {AREA} cm, wl {WL} cm, dz {DZ} cm")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """One oracle process: `steps` stateful propagate steps on its own 1024^2 complex64 beam."""
    seed, steps, n, bits = args
    import numpy as np
    from oracle import beam_oracle as orc
    rng = np.random.default_rng(seed)
    ct = np.complex64 if bits == 64 else np.complex128
    f = (rng.uniform(size=(n, n)) * np.exp(2j * np.pi * rng.uniform(size=(n, n)))).astype(ct)
    b = orc.OracleBeam(AREA, n, WL, init_field=f, complex_bits=bits)
    t0 = time.perf_counter()
    for _ in range(steps):
        b.propagate(DZ)
    return time.perf_counter() - t0


def cpu_arm(steps, warmup, procs):
    """numpy oracle port of the reference path, one process per host core; every step = one
    propagate on `procs` independent beams (a bounded sample of the GPU workload)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(i, max(1, warmup), N_GRID, BITS) for i in range(procs)])
        # every worker times its own `steps` propagate calls (construction excluded); the
        # processes run concurrently, the slowest one bounds the job
        dt = max(pool.map(_cpu_worker, [(i, steps, N_GRID, BITS) for i in range(procs)]))
    return procs * steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    # keep the run bounded: ~0.07 s per 1024^2 complex64 step per core
    steps = min(steps, 40)
    warmup = min(warmup, 3)
    value, dt = cpu_arm(steps, warmup, procs)
    sample = (f"{procs} processes x {steps} propagate steps, one {N_GRID}^2 complex64 beam each "
              f"(numpy pocketfft is single-threaded: one process per host core)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 transforms, f64 phase", "data": "synthetic",
        "config": {"workload": WORKLOAD, "impl_note": "numpy oracle port of lightprop2d.Beam2D.propagate on host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import lightprop2d_b200 as lp
    from lightprop2d_b200 import _lib
    from lightprop2d_b200.ensemble import EnsemblePropagator, shard_range

    def barrier():
        if world > 1:
            dist.barrier()

    K, W = max(1, args.steps), max(3, args.warmup)
    n, B = N_GRID, BATCH_PER_GPU
    total_beams = B * world
    lo, hi = shard_range(total_beams, rank, world)           # contiguous block of beams per rank
    assert hi - lo == B

    # synthetic ensemble, generated on the device from the global beam index (seeded per beam block)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + lo)
    amp = torch.rand(B, n, n, generator=g, device=dev)
    ph = torch.rand(B, n, n, generator=g, device=dev) * (2 * np.pi)
    fields = torch.polar(amp, ph).to(torch.complex64)
    del amp, ph
    beam = lp.Beam2D(AREA, n, WL, init_field=fields, complex_bits=BITS, use_gpu=True, numpy_output=False,
                     device=local_rank)
    beam.spectrum                                            # materialise the spectrum state once
    torch.cuda.synchronize()

    for _ in range(W):
        beam.propagate(DZ)
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(K):
        beam.propagate(DZ)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = total_beams * K / (ms * 1e-3)

    # ---- roofline: the two kernels of the step, each timed alone on the launching stream
    peak, peak_src = measured_peaks()
    lib = _lib.load()
    plan = beam._plan.handle
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    spec, fld = beam._s, beam._f
    PB = n * n * 8

    def time_pass(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / iters

    # chunked exactly like the library does inside lp2d_propagate is not possible per pass; the
    # per-kernel figures below are for whole-batch launches (inputs >> L2, so cold)
    t_row = time_pass(lambda: lib.lp2d_pass(plan, _lib.AXIS_ROWS, _lib.OP_H_FFT, spec.data_ptr(), fld.data_ptr(),
                                            spec.data_ptr(), None, B, 1, DZ, st))
    t_col = time_pass(lambda: lib.lp2d_pass(plan, _lib.AXIS_COLS, _lib.OP_FFT, fld.data_ptr(), fld.data_ptr(),
                                            None, None, B, 1, 0.0, st))
    row_bytes, col_bytes = 3 * PB * B, 2 * PB * B
    step_ms = ms / K
    step_alg = 3 * PB * B                                    # read S, write S, write field
    name_buf = C.create_string_buffer(200)
    fused = lib.lp2d_plan_step_kernel(plan, name_buf, 200) == 1
    per_pass = {"row_pass_ms": t_row, "col_pass_ms": t_col,
                "row_pass_frac": row_bytes / (t_row * 1e-3) / 1e9 / peak, "col_pass_frac": col_bytes / (t_col * 1e-3) / 1e9 / peak,
                "note": "one-kernel-per-pass line kernels timed alone (row: 3*P*B, column: 2*P*B of their own traffic); "
                        "they serve plain fft2 and the fresh chain, and every step when LP2D_PIPE=0"}
    if fused:
        # ONE launch per step: the dominant kernel is the step itself; time it alone
        t_step = time_pass(lambda: lib.lp2d_propagate(plan, spec.data_ptr(), fld.data_ptr(), None, B, DZ, st))
        dominant, dom_t, dom_bytes = "step", t_step, step_alg
        kernel_name = name_buf.value.decode()
    else:
        dominant = "row" if t_row >= t_col else "col"
        dom_t, dom_bytes = (t_row, row_bytes) if dominant == "row" else (t_col, col_bytes)
        kernel_name = ("line_kernel<row, PRE_H> (spectrum*=H, store spectrum, inverse row FFT)" if dominant == "row"
                       else "line_kernel<col> (inverse column FFT in place)")
    achieved = dom_bytes / (dom_t * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "kernel": kernel_name,
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": None,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes,
        "kernel_ms": dom_t,
        "per_pass_kernels": per_pass,
        "step": {"algorithmic_bytes": step_alg, "achieved": step_alg / (step_ms * 1e-3) / 1e9,
                 "frac": step_alg / (step_ms * 1e-3) / 1e9 / peak,
                 "note": "whole propagate step in the timed loop against 3*N^2*8 B per beam (read spectrum, write spectrum, write field)"},
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tj = json.load(f)
            roofline["traffic"] = tj.get(dominant + "_bytes_per_beam", tj.get(dominant + "_pass_bytes_per_beam", 0)) * B or None
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- end to end: host (pinned) fields in, intensities out, through the public ensemble API
    del beam, spec, fld
    torch.cuda.empty_cache()
    prop = EnsemblePropagator(AREA, n, WL, complex_bits=BITS, device=local_rank)
    host_in = torch.empty(E2E_BEAMS, n, n, dtype=torch.complex64, pin_memory=True)
    host_in.copy_(fields[:E2E_BEAMS])
    host_out = torch.empty(E2E_BEAMS, n, n, dtype=torch.float32, pin_memory=True)
    Ke = max(3, min(K, 10))
    for _ in range(2):
        prop.intensity_from_host(host_in, DZ, chunk=8, out=host_out)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(Ke):
        prop.intensity_from_host(host_in, DZ, chunk=8, out=host_out)
    b_.record()
    torch.cuda.synchronize()
    e2e_ms = max(a.elapsed_time(b_), (time.perf_counter() - t0) * 1e3)
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": E2E_BEAMS * world * Ke / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": E2E_BEAMS * n * n * 8, "d2h_bytes_per_step": E2E_BEAMS * n * n * 4,
           "steps": Ke, "beams_per_step_per_gpu": E2E_BEAMS,
           "api": "EnsemblePropagator.intensity_from_host (fft2 -> H -> ifft2 -> |E|^2 per beam)"}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        procs = max(1, min(cores, 64))
        cpu_steps = 40          # ~2 s of wall clock, ~30 core-seconds of CPU work on a 16-core box
        cv, cdt = cpu_arm(steps=cpu_steps, warmup=2, procs=procs)
        cpu = {"value": cv, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"{procs} oracle processes x {cpu_steps} propagate steps on one {n}^2 complex64 beam each ({cdt:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 transforms, f64 phase", "data": "synthetic",
            "config": {"workload": WORKLOAD, "npoints": n, "complex_bits": BITS, "beams_per_gpu": B,
                       "total_beams": total_beams, "sharding": "contiguous beam blocks, no collective",
                       "l2": "inputs (2 GiB spectrum + 2 GiB field per GPU) are larger than the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

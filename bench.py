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
             the reference's OWN lightprop2d.Beam2D (numpy path; installed unmodified under oracle/_ref by
             oracle/build_ref.py and imported through the cupy stub of oracle/ref_shim.py) on the box's host
             cores, one process per core; the oracle port only if oracle/_ref did not travel
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
def _ref_kind():
    """'reference' when the unmodified reference package can be imported here, else 'port' (the oracle)."""
    try:
        from oracle import ref_shim
        return "reference" if ref_shim.available() else "port"
    except Exception:
        return "port"


def _make_beam(kind, field, bits):
    if kind == "reference":
        from oracle import ref_shim
        ref = ref_shim.load()
        return ref.Beam2D(AREA, field.shape[-1], WL, init_field=field, complex_bits=bits)
    from oracle import beam_oracle as orc
    return orc.OracleBeam(AREA, field.shape[-1], WL, init_field=field, complex_bits=bits)


def _cpu_worker(args):
    """One host process.  mode 'state': `warm` + `steps` stateful propagate(DZ) steps on its own beam (the metric);
    mode 'e2e': per step a NEW field is seeded (`_update_obj`: fft2), propagated and its iprofile read -- the work
    the GPU arm's e2e does per beam (fft2 -> H -> ifft2 -> |E|^2)."""
    seed, warm, steps, n, bits, kind, mode = args
    import numpy as np
    rng = np.random.default_rng(seed)
    ct = np.complex64 if bits == 64 else np.complex128
    f = (rng.uniform(size=(n, n)) * np.exp(2j * np.pi * rng.uniform(size=(n, n)))).astype(ct)
    b = _make_beam(kind, f.copy(), bits)
    acc = 0.0
    for i in range(warm + steps):
        if i == warm:
            t0 = time.perf_counter()
        if mode == "state":
            b.propagate(DZ)
        else:
            if kind == "reference":
                b._update_obj(f.copy())
            else:
                b.update(f.copy())
            b.propagate(DZ)
            acc += float(b.iprofile[0, 0])
    return time.perf_counter() - t0


def cpu_arm(steps, warmup, procs, kind, mode="state", n=None, bits=None):
    """One process per host core (numpy's pocketfft is single-threaded); every step = one propagate on `procs`
    independent beams, a bounded sample of the GPU workload.  Returns (beam-propagations / s, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    n = N_GRID if n is None else n
    bits = BITS if bits is None else bits
    with ctx.Pool(procs) as pool:
        # every worker times its own `steps` calls after `warmup` untimed ones (construction excluded); the
        # processes run concurrently, the slowest one bounds the job
        dt = max(pool.map(_cpu_worker, [(i, warmup, steps, n, bits, kind, mode) for i in range(procs)]))
    return procs * steps / dt, dt


def secondary_baselines():
    """north_star's secondary references: the reference's pyfftw CPU path and its cupy / cuFFT GPU path.  Neither
    package is in this image (no network): reported as unavailable rather than substituted."""
    out = {}
    for mod, what in (("pyfftw", "reference CPU path with pyfftw (beam2d.py:25-34)"),
                      ("cupy", "reference GPU path, cupy + cuFFT (beam2d.py:17-23, 106-109)")):
        try:
            __import__(mod)
            out[mod] = {"available": True, "what": what}
        except Exception as e:                                   # noqa: BLE001
            out[mod] = {"available": False, "what": what, "unavailable": f"{type(e).__name__}: {e}"}
    return out


def time_secondary(sec, procs):
    """Time whichever secondary reference paths are importable (bounded); fills `value` in place."""
    if sec.get("pyfftw", {}).get("available") and _ref_kind() == "reference":
        # with pyfftw importable the reference selects it by itself (beam2d.py:25-34), threads = -1: one process
        v, dt = cpu_arm(steps=10, warmup=2, procs=1, kind="reference")
        sec["pyfftw"].update({"value": v, "unit": UNIT, "sample": f"1 process x 10 propagate steps ({dt:.1f} s), FFTW threads=-1"})
    if sec.get("cupy", {}).get("available") and _ref_kind() == "reference":
        try:
            import cupy as cp
            import numpy as np
            from oracle import ref_shim
            ref = ref_shim.load()
            f = cp.asarray(np.random.default_rng(0).uniform(size=(N_GRID, N_GRID)).astype(np.complex64))
            b = ref.Beam2D(AREA, N_GRID, WL, init_field=f, complex_bits=BITS, use_gpu=True)
            for _ in range(3):
                b.propagate(DZ)
            cp.cuda.Stream.null.synchronize()
            t0 = time.perf_counter()
            for _ in range(20):
                b.propagate(DZ)
            cp.cuda.Stream.null.synchronize()
            sec["cupy"].update({"value": 20 / (time.perf_counter() - t0), "unit": UNIT, "sample": "1 beam x 20 propagate steps"})
        except Exception as e:                                   # noqa: BLE001
            sec["cupy"].update({"available": False, "unavailable": f"{type(e).__name__}: {e}"})
    return sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    kind = _ref_kind()
    # the driver's K and W are honoured; one step costs ~0.07 s per core, so even K = 1000 stays within minutes
    steps, warmup = max(1, min(args.steps, 1000)), max(0, min(args.warmup, 100))
    value, dt = cpu_arm(steps, warmup, procs, kind)
    e2e_steps = max(3, min(steps, 10))
    e2e_value, e2e_dt = cpu_arm(e2e_steps, 1, procs, kind, mode="e2e")
    who = ("lightprop2d.Beam2D of the unmodified reference (oracle/_ref), numpy path" if kind == "reference"
           else "numpy oracle port of lightprop2d.Beam2D (oracle/_ref absent)")
    sample = (f"{procs} processes x {steps} propagate steps (+{warmup} warm-up), one {N_GRID}^2 complex64 beam each "
              f"(numpy pocketfft is single-threaded: one process per host core)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 transforms, f64 phase", "data": "synthetic",
        "config": {"workload": WORKLOAD, "impl_note": who + " on the host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        # same work per beam as the GPU arm's e2e (a new host field each step: fft2 -> H -> ifft2 -> |E|^2);
        # host only, so no bytes cross a link
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "steps": e2e_steps, "work": "per beam: _update_obj(field) [fft2] -> propagate(dz) -> iprofile",
                "stateful_value": value},
        "secondary_baselines": time_secondary(secondary_baselines(), procs),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm
def ctypes_name(lib, plan):
    import ctypes as C
    buf = C.create_string_buffer(240)
    lib.lp2d_plan_step_kernel(plan, buf, 240)
    return buf.value.decode()


def slab_record(rank, world, dev, dist):
    """BASELINE config 4 on the ranks of this run: ONE complex64 grid (32768^2 = 8 GiB on 8 GPUs, 16384^2 on 2 / 4)
    decomposed in row slabs, propagate = local H + column pass, NCCL all-to-all transpose over NVLink, local row
    pass.  Parity at full size against the plane-wave closed form (tools/planewave_check.py); CUDA events, max
    over ranks."""
    import torch
    from lightprop2d_b200.slab import SlabBeam2D, slab_rows
    from tools import planewave_check as pw
    n = 32768 if world >= 8 else 16384
    rec = {"npoints": n, "complex_bits": 64, "ranks": world}
    try:
        rec["planewave_rel_l2"] = pw.check(n, 64, AREA * n / 1024, WL, DZ, rank, world)
        rec["parity_ok"] = max(rec["planewave_rel_l2"].values()) <= 1e-5
        torch.cuda.empty_cache()
        r0, r1 = slab_rows(n, rank, world)
        g = torch.Generator(device=dev)
        g.manual_seed(7 + rank)
        f = torch.polar(torch.rand(r1 - r0, n, generator=g, device=dev),
                        torch.rand(r1 - r0, n, generator=g, device=dev) * 6.2831853).to(torch.complex64)
        b = SlabBeam2D(AREA * n / 1024, n, WL, init_field=f, complex_bits=64)
        del f
        for _ in range(3):
            b.propagate(DZ)
        torch.cuda.synchronize()
        dist.barrier()
        Ks = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Ks):
            b.propagate(DZ)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / Ks], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        sent = n * n * 8 * (world - 1) // (world * world)
        rec.update({"ms_per_propagate": ms, "propagations_per_s": 1e3 / ms,
                    "nvlink_bytes_sent_per_rank": sent, "nvlink_gbs_per_rank": sent / (ms * 1e-3) / 1e9,
                    "wire_time_ms_at_770gbs": sent / 770e9 * 1e3,
                    "hbm_alg_gbs_per_gpu": 3 * n * n * 8 / world / (ms * 1e-3) / 1e9})
        del b
        torch.cuda.empty_cache()
    except Exception as e:                                    # noqa: BLE001  (reported, never hidden)
        rec["error"] = f"{type(e).__name__}: {e}"
    return rec


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
        # ONE launch per step: the dominant kernel IS the step, so its average launch duration over the timed region is
        # the step time of that region (CUDA events on the launching stream, launches back to back); the same kernel
        # timed alone afterwards (same buffers every launch, later in the run) is reported next to it
        t_alone = time_pass(lambda: lib.lp2d_propagate(plan, spec.data_ptr(), fld.data_ptr(), None, B, DZ, st))
        dominant, dom_t, dom_bytes = "step", step_ms, step_alg
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
        "kernel_ms_alone": t_alone if fused else dom_t,
        "kernel_ms_source": ("timed region: K back-to-back launches of the one kernel of the step" if fused
                             else "the kernel timed alone, 10 launches after 3 warm-up launches"),
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
        prop.intensity_from_host(host_in, DZ, chunk=8, out=host_out)      # returns with the result on the host
    b_.record()
    torch.cuda.synchronize()
    e2e_ms = max(a.elapsed_time(b_), (time.perf_counter() - t0) * 1e3)
    barrier()
    # host-link roofline of that step: the same bytes (H2D of the fields, D2H of the intensities) as plain
    # concurrent copies on two streams, all ranks at the same time, no kernel
    dev_in = torch.empty(E2E_BEAMS, n, n, dtype=torch.complex64, device=dev)
    dev_out = torch.empty(E2E_BEAMS, n, n, dtype=torch.float32, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def link_pass():
        with torch.cuda.stream(s_in):
            dev_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            host_out.copy_(dev_out, non_blocking=True)
    for _ in range(2):
        link_pass()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        link_pass()
    torch.cuda.synchronize()
    link_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    del dev_in, dev_out
    if world > 1:
        t = torch.tensor([e2e_ms, link_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, link_ms = float(t[0].item()), float(t[1].item())
    h2d_b, d2h_b = E2E_BEAMS * n * n * 8, E2E_BEAMS * n * n * 4
    e2e_value = E2E_BEAMS * world * Ke / (e2e_ms * 1e-3)
    link_value = E2E_BEAMS * world * Ke / (link_ms * 1e-3)
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
           "steps": Ke, "beams_per_step_per_gpu": E2E_BEAMS,
           "api": "EnsemblePropagator.intensity_from_host (fft2 -> H -> ifft2 -> |E|^2 per beam; result on the host at return)",
           "host_link_roofline": {
               "value": link_value, "unit": UNIT, "frac": e2e_value / link_value,
               "h2d_gbs": h2d_b * world * Ke / (link_ms * 1e-3) / 1e9, "d2h_gbs": d2h_b * world * Ke / (link_ms * 1e-3) / 1e9,
               "how": "the step's H2D + D2H bytes as bare concurrent cudaMemcpyAsync on two streams from the same pinned "
                      "buffers, all ranks at once, max over ranks (aggregate GB/s over the N GPUs)"}}

    # ---- complex128 (BASELINE metric: c64 / c128): the same stateful step on 1024^2 complex128 beams
    c128 = None
    try:
        B128 = 128                                            # 2 GiB spectrum + 2 GiB field
        g.manual_seed(4321 + lo)
        f128 = torch.polar(torch.rand(B128, n, n, generator=g, device=dev, dtype=torch.float64),
                           torch.rand(B128, n, n, generator=g, device=dev, dtype=torch.float64) * (2 * np.pi))
        beam128 = lp.Beam2D(AREA, n, WL, init_field=f128, complex_bits=128, use_gpu=True, numpy_output=False, device=local_rank)
        del f128
        beam128.spectrum
        for _ in range(3):
            beam128.propagate(DZ)
        torch.cuda.synchronize()
        barrier()
        K128 = max(3, min(K, 20))
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K128):
            beam128.propagate(DZ)
        b_.record()
        torch.cuda.synchronize()
        ms128 = a.elapsed_time(b_)
        if world > 1:
            t = torch.tensor([ms128], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms128 = float(t.item())
        nb = ctypes_name(lib, beam128._plan.handle)
        alg128 = 3 * n * n * 16 * B128
        c128 = {"value": B128 * world * K128 / (ms128 * 1e-3), "unit": UNIT, "ms_per_step": ms128 / K128, "steps": K128,
                "workload": f"batched {n}x{n} complex128 stateful Beam2D.propagate(dz), {B128} beams per GPU",
                "kernel": nb, "roofline_frac": alg128 / (ms128 / K128 * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_step": alg128}
        del beam128
        torch.cuda.empty_cache()
    except Exception as e:                                    # noqa: BLE001  (reported, never hidden)
        c128 = {"error": f"{type(e).__name__}: {e}"}

    # ---- the other BASELINE configs, device-side, one GPU (rank 0 of a one-GPU run): config 2 (single complex128 field
    # z-sweep step), config 3 (512^2 complex64 ensemble: fft2 -> H -> ifft2 -> |E|^2 per beam), config 5 (mode projection)
    configs = None
    if world == 1:
        configs = {}

        def ev_ms(fn, iters, warm=3):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a_, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(iters):
                fn()
            b2.record()
            torch.cuda.synchronize()
            return a_.elapsed_time(b2) / iters

        try:
            for nn in (1024, 2048, 4096):
                bm = lp.Beam2D(AREA, nn, WL, init_field_gen=lp.gaussian_beam, init_gen_args=(1, 0.05), complex_bits=128,
                               use_gpu=True, numpy_output=False, device=local_rank)
                bm.lens(10.0)
                t_ms = ev_ms(lambda: bm.propagate(DZ), 20)
                configs[f"config2_single_field_c128_{nn}"] = {
                    "us_per_propagate": t_ms * 1e3, "roofline_frac": 3 * nn * nn * 16 / (t_ms * 1e-3) / 1e9 / peak,
                    "kernel": ctypes_name(lib, bm._plan.handle)}
                del bm
            torch.cuda.empty_cache()
            nb3 = 1024
            prop3 = EnsemblePropagator(200e-4 * 2, 512, 632e-7, complex_bits=64, device=local_rank)
            f3 = prop3.random_fields(nb3, 150e-4, seed=11, phase=True)
            out3 = torch.empty(nb3, 512, 512, dtype=torch.float32, device=dev)
            t_ms = ev_ms(lambda: prop3.intensity(f3, 100e-4, out=out3), 10)
            configs["config3_ensemble_c64_512"] = {
                "beams_per_s": nb3 / (t_ms * 1e-3), "us_per_beam": t_ms * 1e3 / nb3,
                "roofline_frac": nb3 * 512 * 512 * 12 / (t_ms * 1e-3) / 1e9 / peak,
                "work": "device-generated random_round_hole_phase fields -> fft2 -> H -> ifft2 -> |E|^2, one fused launch per batch"}
            del prop3, f3, out3
            torch.cuda.empty_cache()
            for bits5, bt in ((64, 64), (128, 16)):
                ct5 = torch.complex64 if bits5 == 64 else torch.complex128
                modes = torch.randn(256, 512 * 512, dtype=ct5, device=dev)
                fld5 = torch.randn(bt, 512, 512, dtype=ct5, device=dev)
                b5 = lp.Beam2D(200e-4 * 2, 512, 632e-7, init_field=fld5, complex_bits=bits5, use_gpu=True, numpy_output=False,
                               device=local_rank)
                c5 = b5.deconstruct_by_modes(modes)
                t_p = ev_ms(lambda: b5.deconstruct_by_modes(modes), 5, 2)
                t_s = ev_ms(lambda: b5.construct_by_modes(modes, c5), 5, 2)
                byt = (256 * 512 * 512 + bt * 512 * 512) * (8 if bits5 == 64 else 16)
                configs[f"config5_modes_c{bits5}_256x512x512_{bt}beams"] = {
                    "project_ms": t_p, "synthesize_ms": t_s, "project_algorithmic_gbs": byt / (t_p * 1e-3) / 1e9,
                    "project_tflops": 8.0 * bt * 256 * 512 * 512 / (t_p * 1e-3) / 1e12,
                    "kernel": "tcgen05 3xTF32 (complex64)" if bits5 == 64 else "register-tiled FP64 GEMM, split k (complex128)"}
                del b5, modes, fld5, c5
            torch.cuda.empty_cache()
        except Exception as e:                                # noqa: BLE001  (reported, never hidden)
            configs["error"] = f"{type(e).__name__}: {e}"

    # ---- one grid over all the ranks (BASELINE config 4): slab decomposition + all-to-all exchange, N >= 2
    slab = None
    if world > 1:
        slab = slab_record(rank, world, dev, dist)

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload on host cores
    cpu = None
    sec = secondary_baselines()
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        procs = max(1, min(cores, 64))
        kind = _ref_kind()
        cpu_steps = 40          # ~3 s of wall clock, ~45 core-seconds of CPU work on a 16-core box
        cv, cdt = cpu_arm(steps=cpu_steps, warmup=2, procs=procs, kind=kind)
        cpu = {"value": cv, "unit": UNIT, "cores": procs, "kind": kind,
               "sample": f"{procs} processes x {cpu_steps} propagate steps of "
                         f"{'the reference lightprop2d.Beam2D (oracle/_ref)' if kind == 'reference' else 'the oracle port'} "
                         f"on one {n}^2 complex64 beam each ({cdt:.1f} s)"}
        sec = time_secondary(sec, procs)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 transforms, f64 phase", "data": "synthetic",
            "config": {"workload": WORKLOAD, "npoints": n, "complex_bits": BITS, "beams_per_gpu": B,
                       "total_beams": total_beams, "sharding": "contiguous beam blocks, no collective",
                       "l2": "inputs (2 GiB spectrum + 2 GiB field per GPU) are larger than the 126 MB L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "configs": configs,
            "c128": c128, "secondary_baselines": sec,
        }
        if slab is not None:
            line["slab"] = slab
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

#!/usr/bin/env python
"""bench.py — the headline benchmark of the hot path (BASELINE.json): LUT precompute ms at default dims, 4 orders.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full Atmosphere::build command stream (transmittance, direct irradiance, single scattering and three
density / indirect-irradiance / multiple-scattering passes) replayed from the pre-recorded CUDA graph on pre-allocated
images — what /root/reference/benches/precompute.rs:138-148 times (queue_submit of a pre-recorded command buffer +
device_wait_idle), at the default dims of BASELINE.json configs[1] instead of the bench's reduced ones.
At N > 1 every rank builds its own atmosphere per step (independent atmospheres shard with no data-path collective,
SURVEY.md §8e) and `value` is the time per atmosphere over the whole job: weak scaling.

Prints ONE JSON line (rank 0).  Extra objects, at every N:
  `render`  the second half of BASELINE.json's metric: sky evaluation Mpixel/s at 3840x2160 over the 256-view sweep
            (configs[4]); at N > 1 the views are split rank::N, the 8.25 MiB tables are built per rank, no collective;
  `batch`   BASELINE.json configs[3]: 1024 distinct atmospheres (fresh builds, allocation included) split over the ranks;
  `hires`   BASELINE.json configs[2]: ONE high-resolution atmosphere (2 GiB per 3-D table, 8 orders) built by all N
            ranks through fb_pending_run_sharded -- the configuration that communicates (one all-gather of
            scattering_density per order, one-slice halos, irradiance rows over NCCL): strong scaling, with the bytes;
and at N = 1: `roofline` (dominant kernel: scattering_density), `cpu_baseline` (the CPU oracle port: ONE full
default-dims precompute, every texel, measured), `e2e` (host-buffer path through the C ABI), `clocks`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LUT precompute ms (4 orders, default dims)"
UNIT = "ms"
WORKLOAD = "default Earth atmosphere, T 256x64, S 32x128x32x8, E 64x16, 4 orders (BASELINE.json configs[1])"

# Algorithmic work of the dominant kernel, SURVEY.md §8(d): quadrature samples per launch and the per-sample fp32
# flop tallies (as written in scattering_density.comp vs with every loop-invariant hoisted).
DENSITY_SAMPLES = 32 * 128 * 256 * 512
FLOP_AS_WRITTEN = {2: 223.0, 3: 184.0}
FLOP_HOISTED = {2: 92.0, 3: 63.0}


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # the median over samples taken while kernels were in flight: drop idle-clock samples below half of max
        busy = [c for c in sm if mx and c >= 0.5 * mx[0]] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port (fp32 mode = the shaders as written) on the host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_precompute_ms(prefer_reference: bool = True):
    """ONE full default-dims 4-order precompute on the host cores: real tables, every texel of every stage, OpenMP over
    all cores.  A measurement, not an extrapolation.  Runs THE REFERENCE'S OWN SHADERS compiled as C++
    (oracle/_ref/libfb_glsl_ref.so, kind "reference") when that library was built (it needs the reference checkout at
    build time and travels with the snapshot), else the oracle port (kind "port").  Returns (ms, sample text, kind)."""
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core.  libgomp reads
    # the variable when the first OpenMP library is loaded.
    if "oracle.oracle" not in sys.modules:
        os.environ["OMP_NUM_THREADS"] = str(host_cores())
    from oracle import oracle as O
    from oracle import ref_glsl as R
    use_ref = prefer_reference and R.available()
    if use_ref:
        R.set_threads(host_cores())
    O.set_threads(host_cores())
    t0 = time.perf_counter()
    tables = R.precompute(O.Params()) if use_ref else O.precompute(O.Params(), O.F32)
    dt = time.perf_counter() - t0
    ok = bool((tables.scattering[..., :3] >= 0).all()) and float(tables.irradiance.max()) > 0
    what = ("the reference's own GLSL shaders compiled as C++ (oracle/glsl_ref; not lavapipe)" if use_ref
            else "oracle fp32 port (not lavapipe)")
    sample = (f"{what}: one FULL default-dims 4-order precompute, every texel of every stage, "
              f"{dt:.1f} s wall on {host_cores()} cores ({cpu_model()}); result sane: {ok}")
    return dt * 1e3, sample, ("reference" if use_ref else "port")


def ncu_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of k_density_main per launch, from the committed `ncu --set full`
    capture (profiles/r2_precompute_full_raw.csv); None if the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", "r2_precompute_full_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        name, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            if "k_density_main<0>" in r[name]:    # <0>: the order >= 3 instantiation
                return (float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]],
                        "bytes per launch, profiles/r2_precompute_full_raw.csv (order >= 3 launch; tables are L2-resident)")
    except (OSError, ValueError, KeyError, IndexError):
        pass
    return (None, "no ncu capture found")


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  NCCL (its version banner, whatever NCCL_DEBUG asks for) and other native
    libraries write to file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the line is
    written to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank: int):
    """--impl reference: the reference's own CPU-runnable implementation of the path on the host cores.  The crate itself
    (Rust + GLSL on Vulkan) cannot run here -- no cargo, no shaderc, no Vulkan loader/ICD (lavapipe or NVIDIA) in the
    image -- but its SHADERS can: oracle/_ref/libfb_glsl_ref.so is /root/reference/shaders compiled as C++ (kind
    "reference"); the oracle port (kind "port") is the fall-back where that library was not built.  OpenMP over all host
    cores.  Every step is one FULL default-dims 4-order precompute (real tables, every texel); as many of the
    requested steps run as fit in ~100 s, and `steps` reports what ran."""
    if rank != 0:
        return
    budget_s = 100.0
    t_start = time.perf_counter()
    # One warm-up pass at most (it pages the library and the tables in; a CPU pass of 8-16 s has nothing else to warm),
    # then as many timed passes as fit the budget -- at least one.
    warm = min(args.warmup, 1)
    runs, sample = [], ""
    while len(runs) < warm + args.steps:
        spent = time.perf_counter() - t_start
        if len(runs) > warm and spent + spent / len(runs) > budget_s:
            break
        ms, sample, kind = cpu_precompute_ms()
        runs.append(ms)
        if len(runs) == 1 and warm and ms > 30e3:
            warm = 0                       # few host cores: a pass this long is its own warm-up, keep it as a timed step
    vals = runs[warm:]
    v = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": warm, "requested": {"steps": args.steps, "warmup": args.warmup}, "ms_per_step": v,
            "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": host_cores(), "cpu": cpu_model(), "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
            "reference_unavailable": "the Rust crate and its Vulkan pipeline cannot be built or run in this image (no cargo, "
                                     "shaderc, Vulkan loader or ICD): lavapipe and B200-Vulkan baselines are unavailable; "
                                     "kind=reference means the crate's own GLSL shaders compiled as C++ and run on the host cores"}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import api, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — fuzzyblue_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    builder = fb.Builder(local_rank)
    stream = torch.cuda.Stream(device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 2x the 126 MB L2

    def flush_l2():
        with torch.cuda.stream(stream):
            flush_buf.fill_(rank & 0xFF)

    # rank r builds its own atmosphere (rank 0: the default Earth; others: seeded variations of it, same dims)
    params = fb.Parameters() if rank == 0 else synthetic.random_atmospheres(rank, seed=20260)[-1]
    pending = fb.Atmosphere.build(builder, stream, params)
    stream.synchronize()
    launches = pending.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K replays of the recorded command stream --------------------------------------
    for _ in range(max(args.warmup, 3)):
        flush_l2()
        pending.resubmit(stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = []
    for _ in range(args.steps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pending.resubmit(stream)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = ms_per_step / world          # ms per atmosphere over the whole job

    # ---- end to end through the host-buffer path: params from host memory in, the three tables out to pinned host --
    atm = pending.atmosphere()
    P = params
    hT = torch.empty((P.transmittance_r_size, P.transmittance_mu_size, 4), dtype=torch.float32).pin_memory()
    hE = torch.empty((P.irradiance_r_size, P.irradiance_mu_s_size, 4), dtype=torch.float32).pin_memory()
    hS = torch.empty((P.scattering_r_size, P.scattering_mu_size, P.scattering_nu_size * P.scattering_mu_s_size, 4),
                     dtype=torch.float16).pin_memory()
    # the read-backs are recorded into the command stream (fb_pending_set_readback): one call submits the precompute
    # and the copies of the three tables, each leaving as soon as its last writer has run
    pending.set_readback(hT.data_ptr(), hS.data_ptr(), hE.data_ptr())

    def e2e_once():
        pending.resubmit(stream)    # the 320-byte parameter block rides in the kernel arguments of the recorded stream
        stream.synchronize()

    for _ in range(3):
        e2e_once()
    barrier()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush_l2()
        stream.synchronize()        # the L2 flush is not part of the step (as in the device-timed loop above)
        t0 = time.perf_counter()
        e2e_once()
        e2e_s += time.perf_counter() - t0
    barrier()
    pending.set_readback(None, None, None)
    e2e_ms = torch.tensor([e2e_s * 1e3 / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    d2h = hT.numel() * 4 + hS.numel() * 2 + hE.numel() * 4
    clocks = sampler.stop() if rank == 0 else None

    # ---- legs every rank takes part in: the 256-view sweep split over the ranks, the sharded high-resolution build ---
    fma_tflops, sfu_gops = builder.measure_peaks() if rank == 0 else (None, None)
    render = None if args.no_render else render_leg(args, builder, pending, stream, dev, rank, world, fma_tflops, sfu_gops)
    del flush_buf
    batch = None if args.no_batch else batch_leg(args, builder, rank, local_rank, world, dev)
    hires = None if args.no_hires else hires_leg(args, builder, rank, local_rank, world, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rank 0 extras: roofline of the dominant kernel, CPU baseline -------------------------------------------
    roof = {}
    for order in (2, 3):
        n = 10
        for _ in range(3):
            pending.run_stage(api.STAGE_SCATTERING_DENSITY, order=order, stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            pending.run_stage(api.STAGE_SCATTERING_DENSITY, order=order, stream=stream)
        e1.record(stream)
        stream.synchronize()
        roof[order] = e0.elapsed_time(e1) / n
    # one step launches the order-2 variant once and the order>=3 variant twice
    dens_ms = (roof[2] + 2 * roof[3]) / 3
    hoisted = DENSITY_SAMPLES * (FLOP_HOISTED[2] + 2 * FLOP_HOISTED[3]) / 3
    written = DENSITY_SAMPLES * (FLOP_AS_WRITTEN[2] + 2 * FLOP_AS_WRITTEN[3]) / 3
    achieved = hoisted / (dens_ms * 1e-3) / 1e12
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak = json.load(open(peaks_file)).get("hbm_gbs") if os.path.exists(peaks_file) else 6650.0
    roofline = {"kernel": "scattering_density", "bound": "fp32", "achieved": achieved, "peak": fma_tflops, "unit": "TFLOP/s",
                "frac": achieved / fma_tflops if fma_tflops else None, "traffic": ncu_dram_traffic()[0], "traffic_source": ncu_dram_traffic()[1],
                "tally": "hoisted-minimal fp32 flops/sample (92 order 2, 63 order>=3; SURVEY.md §8d), 5.37e8 samples/launch",
                "achieved_as_written_tflops": written / (dens_ms * 1e-3) / 1e12,
                "peak_source": "fb_builder_measure_peaks: FFMA issue microbenchmark on this device (measured); "
                               "MEASURED_PEAKS.json has no FP32 figure",
                "sfu_peak_gops": sfu_gops, "ms_per_launch": {"order2": roof[2], "order3": roof[3]},
                "share_of_step": 3 * dens_ms / ms_per_step,
                "hbm": {"algorithmic_bytes_per_launch": 3 * 8 * 32 * 128 * 256,
                        "achieved_gbs": 3 * 8 * 32 * 128 * 256 / (dens_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if os.path.exists(peaks_file) else "fallback"}}

    # the unit that actually binds the kernel (ncu: l1tex__data_pipe_lsu_wavefronts 80-86 %): shared-memory words returned
    # to the register file, 32 per SM and clock, 6 (12 at order 2) table words per sample
    sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
    lds_peak = builder.sm_count() * 32 * sm_clk * 1e6
    roofline["lds"] = {"words_per_sample": {"order2": 12, "order3": 6}, "peak_words_per_s": lds_peak,
                       "frac": {"order2": DENSITY_SAMPLES * 12 / (roof[2] * 1e-3) / lds_peak,
                                "order3": DENSITY_SAMPLES * 6 / (roof[3] * 1e-3) / lds_peak},
                       "note": "table words only; shuffles, ground rows and the prologue share the same pipe"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, sample, kind = cpu_precompute_ms()
        cpu = {"value": v, "unit": UNIT, "cores": host_cores(), "cpu": cpu_model(), "kind": kind, "sample": sample}

    # the driver's record keeps `config`, `roofline`, `e2e`, `cpu_baseline` of this line and only the NAMES of other
    # keys: the two extra legs are summarised inside `config` (and carried in full under `render` / `hires`)
    legs = {}
    if render:
        legs["render"] = {"metric": render["metric"], "mpixel_s": render["value"], "n_gpus": render["n_gpus"], "views": render["views"],
                          "ms_per_frame_per_gpu": render["ms_per_frame"], "roofline_frac": (render["roofline"] or {}).get("frac"),
                          "roofline_bound": (render["roofline"] or {}).get("bound"),
                          "e2e_mpixel_s": render["e2e"]["value"], "e2e_h2d_bytes": render["e2e"]["h2d_bytes_per_step"],
                          "e2e_d2h_bytes": render["e2e"]["d2h_bytes_per_step"]}
    if batch:
        legs["batch"] = {"metric": batch["metric"], "ms_per_atmosphere": batch["value"], "atmospheres": args.batch_atmospheres,
                         "n_gpus": batch["n_gpus"], "scaling": "strong", "atmospheres_per_second": batch["atmospheres_per_second"]}
    if hires:
        legs["hires"] = {"metric": hires["metric"], "ms": hires["value"], "n_gpus": hires["n_gpus"], "scaling": "strong",
                         "bytes_received_per_rank": hires["collective"]["bytes_received_per_rank_per_step"],
                         "exchanges": hires["collective"]["exchanges_per_step"]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "atmospheres_per_step": world, "timing": "CUDA events per step on the launch "
                       "stream, 256 MiB L2 flush between steps (outside the events), max over ranks",
                       "kernels": "FAST", "legs": legs},
            "e2e": {"value": float(e2e_ms.item()) / world, "unit": UNIT, "h2d_bytes_per_step": 320, "d2h_bytes_per_step": d2h,
                    "note": "one graph replay that carries the precompute and the read-back of transmittance, scattering, irradiance into "
                            "pinned host memory (fb_pending_set_readback), host wall clock from submit to stream sync"},
            "gpu_launches": launches * args.steps, "launches_per_step": launches, "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "render": render, "hires": hires, "batch": batch}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def hires_params(scale: int = 1):
    import fuzzyblue_b200 as fb
    return fb.Parameters(order=8, transmittance_mu_size=1024, transmittance_r_size=256, scattering_r_size=128 // scale,
                         scattering_mu_size=512 // scale, scattering_mu_s_size=128 // scale, scattering_nu_size=32 // scale)


def hires_leg(args, builder, rank: int, local_rank: int, world: int, dev):
    """BASELINE.json configs[2] -- scattering 128x512x128x32 (2 GiB per 3-D table), transmittance 1024x256, 8 orders: ONE
    atmosphere split into r-slabs across the ranks, built by fb_pending_run_sharded (fuzzyblue_b200/csrc/fb_sharded.cu):
    per order one all-gather of scattering_density, one-slice halos of the previous order's table, the irradiance rows,
    all over NCCL.  Strong scaling: the job is fixed, `value` is the device time of one whole precompute (max over ranks).
    Returns the record on rank 0."""
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import sharded
    scale = args.hires_scale          # 1 = the full config; 2 halves every scattering axis (1/16 of the work)
    p = hires_params(scale)
    flags = sharded.GATHER_RESULT | (sharded.NO_PIPELINE if args.hires_no_pipeline else 0)
    stream = torch.cuda.Stream(device=dev)
    comm = sharded.NcclComm(local_rank, rank, world) if world > 1 else None
    pend = fb.Atmosphere.allocate(builder, p)
    sampler = ClockSampler(local_rank)
    warm, steps = 1, max(1, args.hires_steps)
    times = []
    for i in range(warm + steps):
        if i == warm and rank == 0:
            sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        pend.run_sharded(comm, rank, world, flags, stream)
        e1.record(stream)
        stream.synchronize()
        if i >= warm:
            times.append(e0.elapsed_time(e1))
    launches = pend.launch_count()
    slow = pend.slow_stages()
    finite = bool(torch.isfinite(torch.from_numpy(pend.atmosphere().read_irradiance())).all())
    total = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    pend.close()
    if comm is not None:
        comm.close()
    builder.trim()
    if rank != 0:
        return None
    ms = float(total.item()) / steps
    rx = sharded.bytes_received(p, min(1, world - 1), world, flags)
    n_steps = len(sharded.plan(p, 0, world, flags))
    return {"metric": "LUT precompute ms (8 orders, high-resolution dims)", "value": ms, "unit": "ms", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"default Earth physics, T 1024x256, S r{p.scattering_r_size} x mu{p.scattering_mu_size} x "
                                   f"mu_s{p.scattering_mu_s_size} x nu{p.scattering_nu_size}, 8 orders, r-slab sharded "
                                   f"(BASELINE.json configs[2], scale 1/{scale})",
                       "inputs": "40 bytes of table per texel, far larger than L2", "kernels": "FAST",
                       "entry_point": "fb_pending_run_sharded (C ABI, NCCL resolved at run time)"},
            "collective": {"exchanges_per_step": rx["exchanges"], "plan_steps": n_steps,
                           "bytes_received_per_rank_per_step": rx["total"], "all_gather_bytes": rx["all_gather"],
                           "halo_bytes": rx["halo"], "irradiance_row_bytes": rx["rows"],
                           "limiting": "the all-gather of scattering_density before each multiple_scattering pass "
                                       "(7 of them + the final table: (N-1)/N x 2 GiB received per rank each)",
                           "overlap": "none (whole-slab ncclAllGather)" if args.hires_no_pipeline else
                                      "sub-slab exchanges on a communication stream behind the next sub-slab's density kernels"},
            "clocks": clocks, "gpu_launches": launches * steps, "launches_per_step": launches, "slow_stages": slow,
            "results_finite": finite}


def run_hires(args, rank: int, local_rank: int, world: int):
    """--workload hires: the `hires` leg alone, as the run's one JSON line."""
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.hires_steps = args.steps
    line = hires_leg(args, fb.Builder(local_rank), rank, local_rank, world, dev)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def batch_leg(args, builder, rank: int, local_rank: int, world: int, dev):
    """BASELINE.json configs[3] -- 1024 distinct atmospheres (randomised Rayleigh / Mie / ozone, Earth-to-Mars radii,
    default dims, 4 orders) sharded across the ranks with no communication.  Every atmosphere is a FRESH
    fb_atmosphere_build (allocation included, served from the builder's block cache), 32 in flight per GPU on forked
    streams; a rank's result sets stay resident in HBM (8.3 MiB each) until the pass is over.  Returns the record on rank 0."""
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import synthetic
    total = args.batch_atmospheres
    mine = synthetic.random_atmospheres(total, seed=20260)[rank::world]
    stream = torch.cuda.Stream(device=dev)
    CH = 32

    def one_pass(params):
        """Chunk i+1 is enqueued before chunk i is waited for, so the GPU never idles on host-side work."""
        out, prev = [], None
        for i in range(0, len(params), CH):
            pend = fb.build_batch(builder, params[i:i + CH], stream)
            done = torch.cuda.Event()
            done.record(stream)
            if prev is not None:
                prev[1].synchronize()
                out += [p.assert_ready(check=False) for p in prev[0]]
            prev = (pend, done)
        prev[1].synchronize()
        out += [p.assert_ready(check=False) for p in prev[0]]
        return out

    for a in one_pass(mine[:CH]):      # warm-up: module load, block cache
        a.close()
    times, finite = [], True
    for _ in range(max(1, args.batch_steps)):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        atms = one_pass(mine)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        finite = finite and all(bool(torch.isfinite(torch.tensor(a.read_irradiance())).all()) for a in atms[:4])
        for a in atms:
            a.close()
    builder.trim()
    t = torch.tensor([sum(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    sec = float(t.item()) / len(times)
    return {"metric": "LUT precompute ms per atmosphere (4 orders, default dims, batch of distinct atmospheres)",
            "value": sec * 1e3 / total, "unit": "ms", "n_gpus": world, "steps": len(times), "warmup": 1,
            "ms_per_step": sec * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{total} distinct atmospheres (BASELINE.json configs[3]), {len(mine)} per GPU, 32 in flight",
                       "timing": "wall clock around the whole pass incl. allocation, max over ranks", "kernels": "FAST"},
            "atmospheres_per_second": total / sec, "results_finite": finite, "gpu_launches": 16 * len(mine) * len(times),
            "pass_seconds_this_rank": times,
            "note": "the first pass allocates every block with cudaMalloc, later passes are served by the builder's block cache"}


def run_batch(args, rank: int, local_rank: int, world: int):
    """--workload batch: the `batch` leg alone, as the run's one JSON line."""
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.batch_steps = args.steps
    line = batch_leg(args, fb.Builder(local_rank), rank, local_rank, world, dev)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# SURVEY.md §8(a) row a24, as-written tallies per pixel: fp32 flops (fma = 2) and SFU-class operations
RENDER_TALLY = {"geometry": (430, 39), "sky": (290, 29)}


def render_roofline(px: int, geometry_px: int, total_ms: float, fma_tflops, sfu_gops):
    """FP32 / SFU roofline of the sky evaluation over the sweep: the as-written tally of render_sky.h:111-191
    (a sky pixel skips the far-point look-up) against the peaks fb_builder_measure_peaks measured on
    this device.  `frac` = max(F / peak_F, S / peak_S) / measured time.  The FAST kernel hoists per-view invariants, so
    the fraction is an as-written-equivalent rate, not a pipe utilisation (ncu: profiles/r1_render_full_raw.csv)."""
    if not fma_tflops or not sfu_gops or total_ms <= 0:
        return None
    sky_px = px - geometry_px
    flops = geometry_px * RENDER_TALLY["geometry"][0] + sky_px * RENDER_TALLY["sky"][0]
    sfu = geometry_px * RENDER_TALLY["geometry"][1] + sky_px * RENDER_TALLY["sky"][1]
    t = total_ms * 1e-3
    f_frac = flops / t / 1e12 / fma_tflops
    s_frac = sfu / t / 1e9 / sfu_gops
    return {"kernel": "k_render_sky", "bound": "sfu" if s_frac >= f_frac else "fp32", "frac": max(f_frac, s_frac),
            "fp32": {"achieved_tflops": flops / t / 1e12, "peak_tflops": fma_tflops, "frac": f_frac},
            "sfu": {"achieved_gops": sfu / t / 1e9, "peak_gops": sfu_gops, "frac": s_frac},
            "geometry_pixel_share": geometry_px / px if px else None,
            "tally": "as written, per pixel: geometry 430 flop + 39 SFU ops, sky 290 + 29 (SURVEY.md §8a a24)"}


def render_leg(args, builder, pending, stream, dev, rank=0, world=1, fma_tflops=None, sfu_gops=None):
    """Second half of BASELINE.json's metric: sky evaluation at 3840x2160 over a 256-view camera sweep
    (config[4]), in chunks of 8 views (depth + two RGBA32F outputs per chunk = 2.4 GB >> L2).  At N > 1 rank r draws
    every N-th view of the altitude-sorted sweep from its own copy of the default-Earth tables (8.25 MiB, built per rank): no
    collective on the data path; `value` = all 256 views' pixels / the slowest rank's device time."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import fuzzyblue_b200 as fb
    from fuzzyblue_b200 import synthetic
    W, H, VIEWS, CHUNK = 3840, 2160, args.render_views, 8
    own = None
    if rank != 0:      # ranks > 0 hold a randomised atmosphere for the precompute leg: the sweep uses the default Earth
        own = fb.Atmosphere.build(builder, stream, fb.Parameters())
        stream.synchronize()
    atm = (own or pending).atmosphere()
    renderer = fb.Renderer(builder)
    draws, extra = synthetic.camera_sweep(VIEWS, W, H)
    depth = torch.empty((CHUNK, H, W), device=dev)
    color = torch.empty((CHUNK, H, W, 4), device=dev)
    transm = torch.empty((CHUNK, H, W, 4), device=dev)

    def make_depth(k):
        """Analytic ground-sphere depth on the device (same construction as synthetic.analytic_depth)."""
        inv = torch.tensor(extra[k][0], device=dev, dtype=torch.float64)
        eye = torch.tensor(extra[k][1], device=dev, dtype=torch.float64)
        xs = (torch.arange(W, device=dev, dtype=torch.float64) + 0.5) / W * 2 - 1
        ys = (torch.arange(H, device=dev, dtype=torch.float64) + 0.5) / H * 2 - 1
        ny, nx = torch.meshgrid(ys, xs, indexing="ij")
        d = inv[:3, 0, None, None] * nx + inv[:3, 1, None, None] * ny + inv[:3, 3, None, None]
        d = d / d.norm(dim=0)
        R = 6360.0e3
        b = (eye[:, None, None] * d).sum(0)
        disc = b * b - (eye @ eye - R * R)
        t = -b - disc.clamp_min(0).sqrt()
        hit = (disc > 0) & (t > 0)
        fwd = inv[:3, 3] / inv[:3, 3].norm()
        z = t * (fwd[:, None, None] * d).sum(0)
        return torch.where(hit, 0.1 / z.clamp_min(1e-9), torch.zeros_like(z)).float()

    total_ms, px, n_launch, geometry_px = 0.0, 0, 0, 0
    # views dealt to the ranks like cards from a deck sorted by camera altitude: a space camera's frame costs half of a
    # low camera's (many of its rays miss the atmosphere), and rank::N over the unsorted sweep left the slowest rank
    # 12 % behind the mean at N = 8
    by_altitude = sorted(range(VIEWS), key=lambda k: draws[k].camera_position[2])
    mine = by_altitude[rank::world]
    for c0 in range(0, len(mine), CHUNK):
        ks = mine[c0:c0 + CHUNK]
        n = len(ks)
        for j, k in enumerate(ks):
            depth[j] = make_depth(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        chunk_draws = [draws[k] for k in ks]
        if c0 == 0:   # warm-up
            renderer.draw_sweep(stream, atm, chunk_draws, depth, color, transm, W, H)
        e0.record(stream)
        renderer.draw_sweep(stream, atm, chunk_draws, depth, color, transm, W, H)
        e1.record(stream)
        stream.synchronize()
        total_ms += e0.elapsed_time(e1)
        px += n * W * H
        n_launch += 1
        geometry_px += int((depth[:n] > 0).sum().item())      # finite-depth pixels take the two-look-up path
    my_ms = total_ms
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cnt = torch.tensor([px, geometry_px, n_launch], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        total_ms, (px, geometry_px, n_launch) = float(t.item()), [int(v) for v in cnt.tolist()]
    if rank != 0:
        return None
    mpx = px / (total_ms * 1e-3) / 1e6
    # end to end for one 4K frame: depth from pinned host memory in, both RGBA32F outputs back to pinned host memory
    hd = depth[0].cpu().pin_memory()
    hc = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    ht = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    renderer.set_depth_buffer(0, depth[0])

    def frame():
        with torch.cuda.stream(stream):
            depth[0].copy_(hd, non_blocking=True)
            renderer.draw(stream, atm, 0, draws[0], color[0], transm[0], W, H)
            hc.copy_(color[0], non_blocking=True)
            ht.copy_(transm[0], non_blocking=True)
        stream.synchronize()

    frame()
    times = []
    for _ in range(9):
        t0 = time.perf_counter()
        frame()
        times.append(time.perf_counter() - t0)
    e2e_s = sorted(times)[len(times) // 2]      # PCIe-bound (33 MB in, 265 MB out per frame): median of 9 frames
    # per-GPU roofline: this rank's pixels over this rank's time (the sweep's aggregate rate is `value`)
    roofline = render_roofline(px // world, geometry_px // world, my_ms, fma_tflops, sfu_gops)
    return {"metric": "sky evaluation Mpixel/s at 3840x2160", "value": mpx, "unit": "Mpixel/s", "views": VIEWS, "n_gpus": world,
            "scaling": "strong", "sharding": "views sorted by camera altitude and dealt rank::N, tables built per rank, no collective",
            "ms_per_frame": total_ms * world / VIEWS, "ms_sweep": total_ms, "gpu_launches": n_launch,
            "inputs": "depth + outputs per 8-view chunk 2.4 GB > L2",
            "hbm": {"bytes_per_pixel": 36, "achieved_gbs_per_gpu": 36 * (px / world) / (my_ms * 1e-3) / 1e9},
            "roofline": roofline,
            "e2e": {"value": W * H / e2e_s / 1e6, "unit": "Mpixel/s", "h2d_bytes_per_step": W * H * 4,
                    "d2h_bytes_per_step": W * H * 32}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--render-views", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-render", action="store_true")
    ap.add_argument("--no-hires", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--batch-steps", type=int, default=2)
    ap.add_argument("--hires-steps", type=int, default=2)
    ap.add_argument("--workload", default="default", choices=["default", "hires", "batch"])
    ap.add_argument("--batch-atmospheres", type=int, default=1024)
    ap.add_argument("--hires-scale", type=int, default=1)
    ap.add_argument("--hires-no-pipeline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    claim_stdout()
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        run_reference(args, rank)
    elif args.workload == "hires":
        run_hires(args, rank, local_rank, world)
    elif args.workload == "batch":
        run_batch(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- rays/s of the sphere-tracing hot path (BASELINE.json metric) on N GPUs of one node.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                   (CPU arm: oracle port of the reference, rank 0 only)

A "step" is one full differentiable render of the workload -- SDFRenderer.render() (depth + normal + silhouette,
50-step recursive march, buffer 5) with gradients enabled w.r.t. the 256-d latent, a scalar loss, and backward()
-- on synthetic inputs (seeded geometric-init 8x512 DeepSDF decoder, seeded latent, fixed camera).  At N GPUs the
image has round(512*sqrt(N))^2 pixels (same view, finer sampling), rows interleaved over ranks, so every GPU traces
~512*512 rays (weak scaling); the per-rank bands and partial gradients are exchanged with ONE all-gather per step.

`value`   : rays/s with inputs resident in HBM (CUDA events, max over ranks).
`e2e`     : same metric through the public API with HOST buffers: H2D of latent/R/T from pinned memory and D2H of
            all four output maps + the latent gradient inside the timed region.
`roofline`: the decoder-row kernel (dominant) timed alone on a 262,144-row batch; achieved = rows/s * F
            (F = 3,146,752 flop per folded row) against the measured dense bf16 peak of MEASURED_PEAKS.json.
"""
import argparse
import gc
import importlib
import json
import math
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

HW_BASE = 512
MARCH_STEP, BUFFER = 50, 5
KIND = "recursive"
CPU_SAMPLE_HW = 128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def loss_of(out):
    depth, normal, mask, min_sdf = out
    return depth[mask.bool()].sum() + min_sdf.sum()


class ClockSampler(object):
    """NVML sampling (background thread, 20 ms period) of SM clocks, power and throttle reasons during the timed
    region.  NVML in-process instead of an `nvidia-smi -lms` subprocess: the subprocess costs ~100 ms per sample and
    measurably slows a 50-80 ms step."""

    def __init__(self, index):
        self.index, self.rows, self.stop, self.th, self.h, self.nv = index, [], False, None, None, None

    def __enter__(self):
        try:
            if self.index < 0:
                raise RuntimeError('sampler disabled')
            import threading
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)

            def run():
                while not self.stop:
                    try:
                        self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                          nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                          nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
                    except Exception:
                        pass
                    time.sleep(0.02)
            self.th = threading.Thread(target=run, daemon=True)
            self.th.start()
        except Exception:
            self.th = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.th is not None:
            self.th.join(timeout=2)

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.rows:
            return out
        nv = self.nv
        out["sm_mhz"] = statistics.median(r[0] for r in self.rows)
        out["sm_max_mhz"] = self.max_sm
        out["power_w_max"] = max(r[1] for r in self.rows)
        out["samples"] = len(self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[2]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        out["reasons"] = [k for k, v in names.items() if bits & v]
        return out


def workload(n_gpus):
    side = int(round(HW_BASE * math.sqrt(n_gpus)))
    return side


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference (oracle/sdf_oracle.py; the reference tree itself is Python and is
    absent on the GPU box) on a bounded sample of the workload: a 128x128 render of the same view."""
    if rank != 0:
        return
    from oracle.sdf_oracle import OracleSDFRenderer
    synth = importlib.import_module("dist-renderer_b200.synth")
    cores = pick_threads(synth)
    dec = synth.make_decoder("B")
    H = W = CPU_SAMPLE_HW
    K = synth.intrinsic(H, W)
    R, T = synth.front_camera()
    ren = OracleSDFRenderer(dec, K, img_hw=(H, W), march_step=MARCH_STEP, buffer_size=BUFFER)
    lat = synth.make_latent()

    def step():
        l = lat.clone().requires_grad_(True)
        out = ren.render(l, R, T, ray_marching_type=KIND)
        loss_of(out).backward()
        return l.grad

    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    value = H * W * args.steps / dt
    side = workload(args.gpus)
    sample = "%dx%d render fwd+bwd of the same view (1/%d of the %dx%d workload's rays) per step" % (
        H, W, (side * side) // (H * W), side, side)
    line = {
        "impl": "reference", "metric": "rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.gpus, side),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def pick_threads(synth):
    """Thread count that gives the oracle's decoder GEMMs the best throughput on this host (many-core boxes lose
    badly to oversubscription at os.cpu_count() threads)."""
    dec = synth.make_decoder("B")
    x = torch.cat([synth.make_latent().expand(20000, -1), torch.rand(20000, 3) - 0.5], 1)
    cores = os.cpu_count() or 1
    best = (None, 1)
    for n in sorted(set(min(cores, c) for c in (8, 16, 32, 64, 128))):
        torch.set_num_threads(n)
        with torch.no_grad():
            dec.inference(x[:2000])
            t0 = time.time()
            dec.inference(x)
            dt = time.time() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, n)
    torch.set_num_threads(best[1])
    return best[1]


def cpu_port_baseline(synth, lat_h, R_h, T_h, side, repeats=1):
    """Oracle port of the reference on the host cores, bounded sample: a 128x128 render fwd+bwd of the same view."""
    from oracle.sdf_oracle import OracleSDFRenderer
    threads = pick_threads(synth)
    dec_c = synth.make_decoder("B")
    Hc = CPU_SAMPLE_HW
    ora = OracleSDFRenderer(dec_c, synth.intrinsic(Hc, Hc), img_hw=(Hc, Hc), march_step=MARCH_STEP, buffer_size=BUFFER)
    best = None
    for _ in range(repeats):
        l = lat_h.clone().requires_grad_(True)
        t0 = time.time()
        o = ora.render(l, R_h, T_h, ray_marching_type=KIND)
        loss_of(o).backward()
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    return {"value": Hc * Hc / best, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": "%dx%d render fwd+bwd of the same view (1/%d of the workload's rays), %d threads of %d host "
                      "cores (best of a thread sweep)" % (Hc, Hc, side * side // (Hc * Hc), threads, os.cpu_count() or 1)}


def config_of(n_gpus, side):
    return {"workload": "%dx%d render(): depth+normal+silhouette, single shape (geometric-init 8x512 DeepSDF, 256-d "
                        "latent), %d-step '%s' march, buffer %d, fwd + backward over the latent"
                        % (side, side, MARCH_STEP, KIND, BUFFER),
            "rays_per_gpu": side * side // n_gpus, "parallelism": "ray-tile (interleaved row bands) x%d" % n_gpus,
            "l2": "flushed between timed iterations (256 MiB write)"}


def ncu_traffic(tc):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on the same 262,144-row batch,
    from the committed `ncu --set full` capture (profiles/r1_*_raw.csv); None when no capture is committed."""
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                        "r1_tc_fwd_raw.csv" if tc else "r1_simt_fwd_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        col = {h: i for i, h in enumerate(rows[0])}
        return sum(float(rows[2][col[k]]) * unit[rows[1][col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:
        return None


def main():
    # NCCL writes its version / INFO lines to stdout by default; stdout carries exactly one JSON line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert world == args.gpus or world == 1, "launch with torch.distributed.run --nproc-per-node N for --gpus N"
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("dist-renderer_b200")
    synth = importlib.import_module("dist-renderer_b200.synth")
    par = importlib.import_module("dist-renderer_b200.parallel")
    abi = importlib.import_module("dist-renderer_b200._abi")
    __import__("__graft_entry__").build()
    lib = abi.lib()

    side = workload(world)
    dec = synth.make_decoder("B").to(dev)
    K = synth.intrinsic(side, side)
    R_h, T_h = synth.front_camera()
    lat_h = synth.make_latent()
    ren = par.ShardedSDFRenderer(dec, K, (side, side), rank=rank, world_size=world, march_step=MARCH_STEP,
                                 buffer_size=BUFFER, engine=args.engine)
    flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)  # 256 MiB > 126 MB L2
    lat_d, R_d, T_d = lat_h.to(dev), R_h.to(dev), T_h.to(dev)
    n_lat = lat_h.numel()

    def step_device(lat_src, R_src, T_src):
        lat = lat_src.detach().requires_grad_(True)
        out = ren.render(lat, R_src, T_src, ray_marching_type=KIND)
        loss_of(out).backward()
        full, extras = ren.gather(out, extra=lat.grad)
        g = extras.sum(0) if extras is not None else lat.grad
        return full, g

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_device(lat_d, R_d, T_d)
        flush.fill_(1.0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ren.local.reset_row_counter()
    gc.collect()
    gc.disable()   # no cyclic-GC pauses inside the timed regions (re-enabled below)
    l0 = lib.dist_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank if not os.environ.get('BENCH_NO_SAMPLER') else -1) as clk:
        torch.cuda.synchronize()
        e0.record()
        marks = []
        for _ in range(args.steps):
            full, g = step_device(lat_d, R_d, T_d)
            flush.fill_(1.0)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append(ev)
        e1.record()
        torch.cuda.synchronize()
    step_ms = [round((e0 if i == 0 else marks[i - 1]).elapsed_time(marks[i]), 2) for i in range(len(marks))]
    launches = lib.dist_launch_count() - l0
    ms = e0.elapsed_time(e1)
    rows_f, rows_g = int(ren.local.rows_evaluated.item()), int(ren.local.rows_grad.item())

    # ---- the same steps once more with every decoder-kernel launch bracketed by CUDA events on its stream
    # (dist_profile_begin/end): launch durations of the dominant kernel inside the running step, GPU still under load
    import ctypes
    n_prof = max(1, min(args.steps, 5))
    ren.local.reset_row_counter()
    torch.cuda.synchronize()
    abi.check(lib.dist_profile_begin())
    p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(n_prof):
        step_device(lat_d, R_d, T_d)
        flush.fill_(1.0)
    p1.record()
    torch.cuda.synchronize()
    k_total_ms, k_launches = ctypes.c_double(0.0), ctypes.c_longlong(0)
    abi.check(lib.dist_profile_end(ctypes.byref(k_total_ms), ctypes.byref(k_launches)))
    prof = {"steps": n_prof, "step_ms": p0.elapsed_time(p1) / n_prof, "kernel_ms": k_total_ms.value / n_prof,
            "launches": k_launches.value / n_prof, "rows_f": int(ren.local.rows_evaluated.item()) / n_prof,
            "rows_g": int(ren.local.rows_grad.item()) / n_prof}
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = side * side * args.steps / (ms * 1e-3)
    clocks = clk.summary()

    # ---- end to end through the public API with host buffers
    pin = lambda x: x.clone().pin_memory()
    lat_p, R_p, T_p = pin(lat_h), pin(R_h), pin(T_h)
    outs_p = [torch.empty(side, side).pin_memory(), torch.empty(side, side, 3).pin_memory(),
              torch.empty(side, side, dtype=torch.uint8).pin_memory(), torch.empty(side, side).pin_memory()]
    g_p = torch.empty(n_lat).pin_memory()

    def step_e2e():
        full, g = step_device(lat_p.to(dev, non_blocking=True), R_p.to(dev, non_blocking=True),
                              T_p.to(dev, non_blocking=True))
        if rank == 0:
            for dst, src in zip(outs_p, full):
                dst.copy_(src, non_blocking=True)
            g_p.copy_(g.reshape(-1), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
        flush.fill_(1.0)
    e1.record()
    torch.cuda.synchronize()
    ms_e = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    t = torch.tensor([ms_e], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e = float(t.item())
    gc.enable()
    e2e_value = side * side * args.steps / (ms_e * 1e-3)
    h2d = (lat_p.numel() + R_p.numel() + T_p.numel()) * 4
    d2h = sum(o.numel() * o.element_size() for o in outs_p) + g_p.numel() * 4

    # ---- roofline of the dominant kernel: decoder rows, timed alone (rank 0)
    F = ren.local.flops_per_row()
    roof = None
    cpu_baseline = None
    if rank == 0:
        peak_burst, peak_sust, _, src = load_peaks()
        n_rows = 262144
        gen = torch.Generator().manual_seed(11)
        pts = ((torch.rand(n_rows, 3, generator=gen) - 0.5) * 1.2).to(dev)
        for _ in range(2):
            pkg.decode_sdf(dec, lat_d, pts, clamp_dist=None, no_grad=True, engine=args.engine)
        reps = 5
        torch.cuda.synchronize()
        time.sleep(1.5)     # "timed alone": let the power-capped clocks of the long timed loops above recover
        e0.record()
        for _ in range(reps):
            pkg.decode_sdf(dec, lat_d, pts, clamp_dist=None, no_grad=True, engine=args.engine)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        isolated = n_rows * F / (k_ms * 1e-3) / 1e12
        in_step = (rows_f * F + rows_g * 2 * F) / (ms * 1e-3) / 1e12
        # dominant kernel inside the step: useful flops of its launches / their summed event-timed durations
        achieved = (prof["rows_f"] * F + prof["rows_g"] * 2 * F) / (prof["kernel_ms"] * 1e-3) / 1e12
        passes = 3 if ren.local.plan.tc is not None else 1
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak_sust, "unit": "TFLOP/s",
                "frac": achieved / peak_sust, "traffic": ncu_traffic(ren.local.plan.tc is not None),
                "peak_source": src + " dense bf16, sustained (kernel timed inside the running step)",
                "kernel": "decoder-row tile kernel: %.0f launches/step, %.1f us average, %.1f %% of the step (CUDA events "
                          "around every launch, %d extra steps after the timed region)"
                          % (prof["launches"], 1e3 * prof["kernel_ms"] / max(prof["launches"], 1),
                             100.0 * prof["kernel_ms"] / prof["step_ms"], prof["steps"]),
                "flop_per_row": F, "launches_per_step": prof["launches"], "kernel_ms_per_step": prof["kernel_ms"],
                "kernel_share_of_step": prof["kernel_ms"] / prof["step_ms"],
                "issued_tflops": passes * achieved, "issued_frac": passes * achieved / peak_sust,
                "isolated": {"tflops": isolated, "rows_per_launch": n_rows, "ms_per_launch": k_ms,
                             "frac_of_burst_peak": isolated / peak_burst, "issued_frac_of_burst_peak": passes * isolated / peak_burst,
                             "burst_peak": peak_burst},
                "traffic_note": "dram bytes of one 262,144-row launch (ncu --set full capture under profiles/)",
                "note": "achieved counts USEFUL flops (F per folded decoder row, 2F per gradient row); the tensor-core engine "
                        "issues 3 fp16 MMA passes per logical GEMM (split-fp16 for fp32-level parity), see issued_*; ncu "
                        "tensor-pipe activity is in profiles/r1_tc_summary.md",
                "whole_step_tflops": in_step, "rows_fwd_per_step": rows_f / args.steps, "rows_grad_per_step": rows_g / args.steps}
        # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
        if world == 1 and not args.no_cpu_baseline:
            cpu_baseline = cpu_port_baseline(synth, lat_h, R_h, T_h, side)
    if rank == 0:
        line = {
            "metric": "rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (split-fp16 tensor-core operands, fp32 accumulate)"
            if ren.local.plan.tc is not None else "f32", "data": "synthetic",
            "config": config_of(world, side), "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e / args.steps},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu_baseline,
            "engine": "tc" if ren.local.plan.tc is not None else "simt", "step_ms": step_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — segmented cumprod fwd+bwd throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c1|c4]

One *step* = one pass of the hot path over one view's element list: grouped_cumprod_forward
followed by grouped_cumprod_backward (the reference ops of cuda_kernel.cpp:17-22) on synthetic
data resident in HBM.  Default workload = BASELINE.json configs[2] ("synthetic 1080p, 1M Gaussians,
heavy-tailed per-pixel segment lengths", scan-only route of SURVEY.md §8d: K = 1920*1080 segments,
lognormal(ln 20, 1) lengths, N ~ 68 M elements) — the config the north star's 70 % target is quoted on.

Prints ONE JSON line (rank 0).  value = Gelem/s fwd+bwd = (elements of all ranks) / max-over-ranks
device time.  N > 1: one process per GPU (torchrun), each rank owns its own view (seed 1080+rank);
views are independent (gs_model.py:402), so there is no data-path collective: scaling = weak.

--impl reference: the CPU arm, rank 0 only.  The reference has no CPU implementation of these ops
(SURVEY.md §8c); the arm BASELINE.json names is "reference PyTorch grouped_cumprod fwd+bwd on CPU": the
pure-PyTorch restatement in oracle/torch_cpu_path.py, all host threads.  The oracle's much faster C/OpenMP
port is reported beside it (cpu_baseline_c_omp).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "segmented cumprod fwd+bwd throughput"
UNIT = "Gelem/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per launch from the committed ncu --set full capture, if any (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:  # noqa: BLE001
            return {}
    return {}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax = float(f[2])
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        except Exception:  # noqa: BLE001
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


L2_BYTES = 126e6


def config_for(e, world: int) -> dict:
    """The `config` object of the JSON line — the same keys and values in both arms (ours / --impl reference)."""
    ws = 4 * e.n * 4   # x, key/inv, grad_out/y in, one array out per op
    l2 = ("inputs exceed L2 (3 x %.0f MB read + %.0f MB written per op), no flush needed" % (e.n * 4 / 1e6, e.n * 4 / 1e6)
          if ws > 2 * L2_BYTES else
          "working set %.0f MB fits the 126 MB L2: a 256 MB buffer is rewritten between timed ops (L2 flush)" % (ws / 1e6))
    return {"workload": e.name, "elements_per_gpu": e.n, "segments_per_gpu": e.k, "l2": l2,
            "parallelism": f"views x{world} (no data-path collective)"}


def make_workload(name: str, device, view: int):
    from simplegaussiansplat_tk71_b200 import workloads as wl

    if name == "c1":
        return wl.c1(device)
    if name == "c4":
        return wl.c4(device)
    return wl.c3(device, view=view)


def cpu_port_run(e_cpu, min_seconds: float, max_reps: int):
    """Time the oracle's C/OpenMP port (fwd + division-free bwd) on host tensors.  Returns (Gelem/s, reps, cores)."""
    import numpy as np

    from oracle import oracle as orc

    x, g, key = e_cpu.x.numpy(), e_cpu.grad_out.numpy(), e_cpu.key.numpy()
    starts = orc.segment_starts(key)
    y = np.empty_like(x)
    gin = np.empty_like(x)
    orc.fwd_bwd_f32_omp(x, g, starts, 0, y, gin)  # warm-up (page faults of the outputs)
    times = []
    t_all = time.perf_counter()
    while len(times) < max_reps and (len(times) < 3 or time.perf_counter() - t_all < min_seconds):
        t0 = time.perf_counter()
        orc.fwd_bwd_f32_omp(x, g, starts, 0, y, gin)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return x.shape[0] / med / 1e9, len(times), orc.max_threads(), med


def torch_cpu_run(e_cpu, reps: int):
    """The 'reference PyTorch path on CPU' (BASELINE.json configs[0] / north_star): pure-PyTorch segmented cumprod
    forward + autograd backward (oracle/torch_cpu_path.py), all host threads.  The bucketing plan (a function of
    the keys only) is built once outside the timed region.  Returns (Gelem/s, median s, threads)."""
    import torch

    from oracle.torch_cpu_path import grouped_cumprod_fwd_bwd

    torch.set_num_threads(os.cpu_count() or 1)
    _, _, plan = grouped_cumprod_fwd_bwd(e_cpu.x, e_cpu.key, e_cpu.grad_out)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        grouped_cumprod_fwd_bwd(e_cpu.x, e_cpu.key, e_cpu.grad_out, plan)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return e_cpu.n / med / 1e9, med, torch.get_num_threads()


def run_reference_arm(args):
    """CPU arm, rank 0 only.  The reference ships no CPU implementation of its ops; the arm
    BASELINE.json names is the pure-PyTorch CPU path, restated in oracle/torch_cpu_path.py.  The much faster
    C/OpenMP port of the oracle is reported beside it (cpu_baseline_c_omp) as a stronger CPU bound."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import oracle as orc
    from oracle.torch_cpu_path import grouped_cumprod_fwd_bwd

    orc.build()
    torch.set_num_threads(os.cpu_count() or 1)
    e = make_workload(args.workload, "cpu", 0)
    _, _, plan = grouped_cumprod_fwd_bwd(e.x, e.key, e.grad_out)
    for _ in range(max(1, min(args.warmup, 2))):
        grouped_cumprod_fwd_bwd(e.x, e.key, e.grad_out, plan)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        grouped_cumprod_fwd_bwd(e.x, e.key, e.grad_out, plan)
    dt = time.perf_counter() - t0
    val = e.n * args.steps / dt / 1e9
    cores = torch.get_num_threads()
    c_val, c_reps, c_cores, c_med = cpu_port_run(e, min_seconds=3.0, max_reps=6)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_for(e, args.gpus),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"whole view ({e.n} elements, {e.k} segments) per step; pure-PyTorch CPU path "
                                   "(length-bucketed torch.cumprod + autograd backward, oracle/torch_cpu_path.py) — "
                                   "the reference has no CPU implementation of its ops"},
        "cpu_baseline_c_omp": {"value": c_val, "unit": UNIT, "cores": c_cores, "kind": "port",
                               "sample": f"same view x {c_reps} reps, median {c_med * 1e3:.1f} ms; C/OpenMP port "
                                         "(oracle/gcp_oracle.c)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_OUT, flush=True)


_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c3", "c4"])
    ap.add_argument("--variant-fwd", type=int, default=-1)
    ap.add_argument("--variant-bwd", type=int, default=-1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="time every kernel variant (stderr table), then exit")
    ap.add_argument("--views", type=int, default=64, help="views per training step for the multi-view leg (C5)")
    ap.add_argument("--no-splat", action="store_true", help="skip the splat-step / multi-view legs")
    ap.add_argument("--no-reference-legs", action="store_true",
                    help="skip timing the reference's CUDA ops and its compositor Function beside ours")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly the JSON line(s) this script prints: everything libraries write to file descriptor 1
    # (NCCL's version banner, for one) is sent to stderr instead
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import ops
    from simplegaussiansplat_tk71_b200 import workloads as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=device)

    e = make_workload(args.workload, device, view=rank)
    n, k = e.n, e.k
    y = torch.empty_like(e.x)
    gin = torch.empty_like(e.x)
    ops.set_variant("fwd", args.variant_fwd)
    ops.set_variant("bwd", args.variant_bwd)

    def step():
        gc.grouped_cumprod_forward(e.x, e.key, y)
        gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.sweep:
        sweep(args, e, y, gin, gc, ops)
        return

    for _ in range(args.warmup):
        step()
    launches_per_step = 0
    gc.grouped_cumprod_forward(e.x, e.key, y)
    launches_per_step += ops.last_launch_count()
    gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
    launches_per_step += ops.last_launch_count()
    assert ops.workspace_status(device) == 0, "watchdog fired during warm-up"

    # ---- timed region: K steps, CUDA events on the launching (current) stream ----
    # L2-resident workloads (C1): a 256 MB buffer is rewritten before every timed op, outside the op's event pair
    flush = torch.empty(64 << 20, dtype=torch.float32, device=device) if 16 * n <= 2 * L2_BYTES else None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * args.steps + 1)]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.1)
    barrier()
    if flush is None:
        # two events per step (step boundary, forward | backward): every record between two kernels costs ~4 us
        # of device idle, so no more of them than the per-kernel durations of the roofline need
        ev[0].record()
        for i in range(args.steps):
            gc.grouped_cumprod_forward(e.x, e.key, y)
            ev[2 * i + 1].record()
            gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
            ev[2 * i + 2].record()
        barrier()
        fwd_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)) / args.steps
        bwd_ms = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)) / args.steps
        total_ms = ev[0].elapsed_time(ev[2 * args.steps])          # the whole bracket: K back-to-back steps
    else:
        ev[0].record()
        for i in range(args.steps):
            flush.fill_(1.0)
            ev[4 * i + 1].record()
            gc.grouped_cumprod_forward(e.x, e.key, y)
            ev[4 * i + 2].record()
            flush.fill_(2.0)
            ev[4 * i + 3].record()
            gc.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end)
            ev[4 * i + 4].record()
        barrier()
        fwd_ms = sum(ev[4 * i + 1].elapsed_time(ev[4 * i + 2]) for i in range(args.steps)) / args.steps
        bwd_ms = sum(ev[4 * i + 3].elapsed_time(ev[4 * i + 4]) for i in range(args.steps)) / args.steps
        total_ms = (fwd_ms + bwd_ms) * args.steps                   # the ops' own event pairs (flush excluded)
    time.sleep(0.05)
    clocks = sampler.stop()
    # The timed region above is K steps = a few milliseconds, too short for more than a few samples of the clocks.
    # The same step run back to back for half a second, no events inside, is reported beside it with its own clock
    # samples (`sustained`): the GPU then touches its software power cap, and the step is still a little FASTER than
    # in the timed region, because nothing separates the kernels.
    sustained = None
    if flush is None:
        s2 = ClockSampler(local)
        s2.start()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(300):
            step()
        torch.cuda.synchronize()
        reps_s = 1500
        a_.record()
        for _ in range(reps_s):
            step()
        b_.record()
        torch.cuda.synchronize()
        ms_s = a_.elapsed_time(b_) / reps_s
        sustained = {"value": n / (ms_s * 1e-3) / 1e9, "unit": UNIT + " per GPU", "ms_per_step": ms_s, "steps": reps_s,
                     "clocks": s2.stop()}
    del flush
    assert ops.workspace_status(device) == 0, "watchdog fired during the timed region"

    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    nn = torch.tensor([float(n)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
    max_ms = float(t.item())
    total_elems = float(nn.item())
    value = total_elems * args.steps / (max_ms * 1e-3) / 1e9

    # ---- e2e: the same step through the public host-buffer API (simplegaussiansplat_tk71_b200.host), pinned HOST
    #      arrays in, pinned HOST arrays out, every copy (and the host-side key packing) inside the timed region.
    #      The streamer uploads x, grad_out and one run-start bit per element of key (the device rebuilds dense
    #      segment ids from the bits), cuts the list at segment boundaries and overlaps H2D, the scan launches and
    #      D2H of neighbouring chunks on three streams. ----
    from simplegaussiansplat_tk71_b200.host import HostStreamer

    e2e_steps = max(1, args.e2e_steps)
    hx, hk, hg, hi, hs = (t_.cpu().pin_memory() for t_ in (e.x, e.key, e.grad_out, e.inv, e.seg_end))
    hy = torch.empty(n, dtype=torch.float32).pin_memory()
    hgin = torch.empty(n, dtype=torch.float32).pin_memory()
    streamer = HostStreamer(device, chunk_elems=8 << 20, depth=3)
    h2d, d2h = streamer.fwd_bwd(hx, hk, hg, hy, hgin)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        # sync=False: the call returns when everything is queued, the current stream waits for the downloads —
        # consecutive steps pipeline (download of step i beside the upload of step i+1), the events bracket it all
        streamer.fwd_bwd(hx, hk, hg, hy, hgin, sync=False)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = total_elems * e2e_steps / (float(t2.item()) * 1e-3) / 1e9
    # the host results of the streamed step equal the resident ones
    # the chunks shift the tile grid, so the fp32 association differs from the resident call's: compared at twice
    # the north star's tolerance (forward) / in relative L2 norm (backward: sums with cancellation)
    yc, gc_ = y.cpu(), gin.cpu()
    assert torch.allclose(hy, yc, rtol=2e-5, atol=2e-6), "streamed forward differs from the resident forward"
    assert float((hgin - gc_).norm() / gc_.norm()) < 1e-5, "streamed backward differs from the resident backward"
    del yc, gc_
    del streamer

    # ---- splat step @1080p and the multi-view step (BASELINE.json configs[4]) ----
    splat = None
    if not args.no_splat:
        del hx, hk, hg, hi, hs, hy, hgin
        splat = splat_legs(args, device, rank, world)
        hx, hk, hi, hs, hg = (t_.cpu() for t_ in (e.x, e.key, e.inv, e.seg_end, e.grad_out))

    # ---- the reference on the same box (rank 0, N == 1 only): its CUDA ops beside ours, and its own compositor
    #      Function with its ops / with the drop-in behind it ----
    ref_ops = ref_fn = None
    if rank == 0 and world == 1 and not args.no_reference_legs:
        try:
            ref_ops = ref_cuda_ops_leg(device, e)
        except Exception as ex:  # noqa: BLE001
            ref_ops = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        if not args.no_splat:
            try:
                ref_fn = reference_function_leg(device)
            except Exception as ex:  # noqa: BLE001
                ref_fn = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
            if splat is not None and isinstance(ref_fn, dict) and "c3_1080p" in ref_fn:
                splat["reference_function_ms"] = ref_fn["c3_1080p"].get("reference_function_ms")
                splat["reference_function_with_dropin_ms"] = ref_fn["c3_1080p"].get("reference_function_with_dropin_ms")

    # ---- CPU baseline (rank 0, N == 1 only): the oracle's C/OpenMP port on the same inputs ----
    cpu = None
    cpu_c = None
    cpu_torch = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as orc

            orc.build()
            e_cpu = wl.ElementList(e.name, hx, hk, hi, hs, hg, e.width, e.height)
            tv, tmed, tthreads = torch_cpu_run(e_cpu, reps=5)
            cpu = {"value": tv, "unit": UNIT, "cores": tthreads, "kind": "port",
                   "sample": f"whole view ({n} elements) x 5 reps, median {tmed * 1e3:.0f} ms; pure-PyTorch CPU path "
                             "(oracle/torch_cpu_path.py), the CPU arm BASELINE.json names"}
            v, reps, cores, med = cpu_port_run(e_cpu, min_seconds=4.0, max_reps=8)
            cpu_c = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                     "sample": f"whole view ({n} elements) x {reps} reps, median {med * 1e3:.1f} ms; C/OpenMP port "
                               "of fwd + division-free bwd, parallel over segments (oracle/gcp_oracle.c)"}
            # BASELINE.json configs[0]: the same PyTorch CPU path on C1 (1 Mi elements / 64 Ki segments)
            c1 = wl.c1("cpu")
            c1v, c1med, _ = torch_cpu_run(c1, reps=5)
            cpu_torch = {"value": c1v, "unit": UNIT, "cores": tthreads, "kind": "port",
                         "sample": f"C1 (1 Mi elements / 64 Ki segments), median of 5 = {c1med * 1e3:.1f} ms"}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        peak, peak_src = _peaks()
        ab = wl.algorithmic_bytes(n, k)
        traffic = _traffic()
        traffic = traffic if args.workload == "c3" else traffic.get(args.workload, {})   # captured per workload
        bwd_gbs = ab["bwd"] / (bwd_ms * 1e-3) / 1e9
        fwd_gbs = ab["fwd"] / (fwd_ms * 1e-3) / 1e9
        both_gbs = ab["fwd_bwd"] / ((fwd_ms + bwd_ms) * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(e, world),
            "variants": {"fwd": ops.variants("fwd")[args.variant_fwd] if args.variant_fwd >= 0 else "default",
                         "bwd": ops.variants("bwd")[args.variant_bwd] if args.variant_bwd >= 0 else "default"},
            "fwd_ms": fwd_ms, "bwd_ms": bwd_ms,
            "roofline": {"bound": "hbm", "kernel": "grouped_cumprod_backward (k_bwd_*)", "achieved": bwd_gbs,
                         "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak, "traffic": traffic.get("bwd"),
                         "peak_source": peak_src, "algorithmic_bytes": ab["bwd"]},
            "roofline_fwd": {"bound": "hbm", "kernel": "grouped_cumprod_forward (k_fwd_*)", "achieved": fwd_gbs,
                             "peak": peak, "unit": "GB/s", "frac": fwd_gbs / peak, "traffic": traffic.get("fwd"),
                             "algorithmic_bytes": ab["fwd"]},
            "roofline_fwd_bwd": {"achieved": both_gbs, "peak": peak, "unit": "GB/s", "frac": both_gbs / peak,
                                 "frac_of_nominal_8TBs": both_gbs / 8000.0},
            "cpu_baseline": cpu, "cpu_baseline_c_omp": cpu_c, "cpu_baseline_torch_c1": cpu_torch,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
            "sustained": sustained,
            "splat_step": splat,
            "ref_cuda_ops": ref_ops,
            "reference_function": ref_fn,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def splat_legs(args, device, rank, world):
    """(i) splat step ms @1080p: compositor forward+backward of ONE view (BASELINE metric, second half);
    (ii) multi-view training step (configs[4]): `--views` views sharded round-robin over the ranks, per-Gaussian
    gradients accumulated into one flat bucket of 38 floats per Gaussian (the reference's parameter set,
    gs_model.py:151-158) and summed with ONE NCCL all-reduce per step."""
    import torch
    import torch.distributed as dist

    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200 import views as vw
    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    mine = vw.views_for_rank(args.views, rank, world)
    distinct = 1                                   # the single-view legs: view 0 of the splat-route workload (numpy rng)
    scenes = [wl.splat_view(1920, 1080, 1_000_000, seed=1080, device=device)]
    n_param = 1_000_000
    bucket = torch.zeros(n_param * vw.PARAM_FLOATS_PER_GAUSSIAN, dtype=torch.float32, device=device)
    leaves = []
    for sc in scenes:
        leaves.append([sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
                       sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True)])
    gI = torch.rand(1081, 1921, 3, device=device) + 0.1
    W, H = torch.tensor(1920), torch.tensor(1080)

    def one_view(i, plan_next=False):
        sc, (m, lam, o, l) = scenes[i % distinct], leaves[i % distinct]
        if plan_next:  # the next view's prologue (offsets, element count) runs on a side stream meanwhile
            nx = scenes[(i + 1) % distinct]
            compositor.plan_view(nx.boxsize, nx.startpoint, nx.endpoint, 1920, 1080)
        for t_ in (m, lam, o, l):
            t_.grad = None
        img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, W, H)
        img.backward(gI)
        # flat bucket laid out like a DDP bucket: one contiguous section per parameter tensor
        # (mean [n,2] | Lambda [n,4] | opacity [n,1] | l [n,3] | the 28 remaining floats per Gaussian)
        k = sc.n
        bucket[0:2 * k] += m.grad.reshape(-1)
        bucket[2 * n_param:2 * n_param + 4 * k] += lam.grad.reshape(-1)
        bucket[6 * n_param:6 * n_param + k] += o.grad.reshape(-1)
        bucket[7 * n_param:7 * n_param + 3 * k] += l.grad.reshape(-1)
        return sc.elements

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for _ in range(2):
        one_view(0)
    torch.cuda.synchronize()
    # (i) single view, through both compositor routes (compositor.ROUTE is the default one and the headline)
    reps = 9
    sc, (m, lam, o, l) = scenes[0], leaves[0]

    def time_single():
        a, bb, c = ev(), ev(), ev()
        tfs, tbs = [], []
        for it in range(reps + 2):
            for t_ in (m, lam, o, l):
                t_.grad = None
            a.record()
            img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, W, H)
            bb.record()
            img.backward(gI)
            c.record()
            torch.cuda.synchronize()
            if it >= 2:
                tfs.append(a.elapsed_time(bb))
                tbs.append(bb.elapsed_time(c))
        tfs.sort()
        tbs.sort()
        # medians: the forward has one host sync and is sensitive to host jitter
        return {"fwd_ms": tfs[reps // 2], "bwd_ms": tbs[reps // 2], "ms": tfs[reps // 2] + tbs[reps // 2]}

    default_route = compositor.ROUTE
    routes = {}
    for r in ("lists", "tiles"):
        compositor.ROUTE = r
        routes[r] = time_single()
    compositor.ROUTE = default_route
    out = {"workload": sc.name, "elements": sc.elements, "gaussians": sc.n, "route": default_route,
           **routes[default_route], "unit": "ms per view (render + backward), median of 9", "routes": routes}
    # the same view through the batch entry point (gcp_views_step, one view): no host wait for the pair count
    nb = vw.NativeViewBatch([sc], 1920, 1080, grad_images=[gI], lanes=1)
    gsec = _sections(bucket, n_param)
    nb.step(*gsec)
    torch.cuda.synchronize()
    assert nb.finish()
    out["native_call_ms"] = _median_ms(lambda: nb.step(*gsec), reps, warm=2)
    out["native_call_launches"] = nb.launches
    del nb
    # roofline of the splat step: the view's kernels are bound by instruction issue (the two walk kernels run at
    # 68-78 % issue-slot utilisation), not by HBM: warp-instructions and DRAM bytes of the committed ncu capture of
    # this very view (profiles/traffic.json: tile_view_1080p) over the live time of the native call
    tv = _traffic().get("tile_view_1080p")
    if tv and default_route == "tiles":
        peak_hbm, _ = _peaks()
        sm_hz = 1.965e9
        issue_peak = 148 * 4 * sm_hz / 1e9           # G warp-instructions / s: 4 schedulers per SM, one per clock
        t_s = out["native_call_ms"] * 1e-3
        out["roofline"] = {"bound": "issue", "achieved": tv["warp_instructions"] / t_s / 1e9, "peak": issue_peak,
                           "unit": "G warp-instructions/s", "frac": tv["warp_instructions"] / t_s / 1e9 / issue_peak,
                           "warp_instructions_per_view": tv["warp_instructions"],
                           "hbm": {"achieved": tv["dram_bytes"] / t_s / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                   "frac": tv["dram_bytes"] / t_s / 1e9 / peak_hbm, "traffic": tv["dram_bytes"]},
                           "time": "native_call_ms (gcp_views_step, one view)",
                           "note": "the walk kernels alone: backward 68 %, render 78 % of their issue slots (ncu)"}
    # end to end from HOST tables to a HOST image and HOST gradients (pinned memory both ways, every copy inside
    # the timed region): what a caller pays who keeps the Gaussians on the host
    out["e2e_host_tables"] = e2e_splat_leg(sc, gI, device)
    # (ii) multi-view training step (BASELINE.json configs[4]): this rank's share of `--views` DISTINCT views (seed
    #      1080 + view id, drawn on the device), one native call per step (views.NativeViewBatch -> gcp_views_step:
    #      render, MSE-loss gradient, backward, scatter-add into the bucket), and the all-reduce of the bucket hidden
    #      behind the last view: the bucket of views 0..V-2 is reduced on a side stream while view V-1 runs into a
    #      small bucket of its own (the 10 floats per Gaussian the compositor produces), which is reduced and added
    #      at the end.
    del scenes, leaves
    torch.cuda.empty_cache()
    out["multi_view"] = multi_view_leg(args, device, rank, world, mine, n_param, bucket)
    if rank == 0 and world == 1 and not os.environ.get("BENCH_SKIP_PY_LOOP"):
        out["multi_view_python_loop"] = python_loop_leg(args, device, mine[: min(len(mine), 16)], n_param, bucket)
    # (iii) C2: the scene bundled with the reference (BASELINE.json configs[1]; poses synthesised), rank 0 only:
    #       splat step per view, and the two scan ops on the element list of view 0 (per-pixel lists of ~350)
    if rank == 0:
        try:
            out["c2_bundled"] = c2_leg(device)
        except Exception as ex:  # noqa: BLE001
            out["c2_bundled"] = {"error": str(ex)}
    return out


def e2e_splat_leg(sc, gI, device):
    import torch

    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    host_in = [t.cpu().pin_memory() for t in (sc.boxsize, sc.startpoint, sc.endpoint, sc.mean.float(), sc.lam, sc.opacity,
                                              sc.l_d, gI)]
    n = sc.n
    host_out = [torch.empty(s_, dtype=torch.float32).pin_memory() for s_ in ((sc.height + 1, sc.width + 1, 3), (n, 2),
                                                                             (n, 2, 2), (n, 1), (n, 3))]
    W, H = torch.tensor(sc.width), torch.tensor(sc.height)

    def once():
        b, sp, ep, m, lam, o, l, g = (t.to(device, non_blocking=True) for t in host_in)
        m, lam, o, l = (t.requires_grad_(True) for t in (m, lam, o, l))
        img = F.apply(b, torch.tensor([n]), sp, ep, m, lam, o, l, W, H)
        img.backward(g)
        for dst, src in zip(host_out, (img.detach(), m.grad, lam.grad, o.grad, l.grad)):
            dst.copy_(src, non_blocking=True)

    ms = _median_ms(once, 7, warm=2)
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    d2h = sum(t.numel() * t.element_size() for t in host_out)
    return {"ms": ms, "h2d_bytes": h2d, "d2h_bytes": d2h,
            "what": "custom_autograd_grouped_cumprod.apply + backward, per-Gaussian tables and dL/dimage from pinned host "
                    "memory, image and the four gradients back to pinned host memory"}


def _sections(bucket, n_param):
    """The four compositor sections of a DDP-like flat bucket (one contiguous section per parameter tensor):
    mean [n,2] | Lambda [n,4] | opacity [n] | l [n,3] | (the 28 remaining floats per Gaussian follow)."""
    return (bucket[0:2 * n_param].view(n_param, 2), bucket[2 * n_param:6 * n_param].view(n_param, 4),
            bucket[6 * n_param:7 * n_param], bucket[7 * n_param:10 * n_param].view(n_param, 3))


def multi_view_leg(args, device, rank, world, mine, n_param, bucket):
    import torch
    import torch.distributed as dist

    from simplegaussiansplat_tk71_b200 import views as vw
    from simplegaussiansplat_tk71_b200 import workloads as wl

    W, H = 1920, 1080
    lanes = int(os.environ.get("BENCH_MV_LANES", "4"))
    views = [wl.splat_view_device(W, H, n_param, seed=1080 + v, device=device) for v in mine]
    elems = sum(v.elements for v in views)
    target = torch.rand(H + 1, W + 1, 3, device=device, generator=torch.Generator(device=device).manual_seed(64))
    # the tail = as many views as there are lanes: the views of the tail run side by side, so the event in front of
    # them fires about one tail's worth of time before the batch ends — enough for the all-reduce of the big bucket
    # (8 GPUs, 8 views per rank: tail 2 / 3 / 4 -> 5.31 / 5.23 / 5.09 ms)
    n_tail = min(int(os.environ.get("BENCH_MV_TAIL", str(lanes))), max(len(views) - 1, 1))
    first_tail = len(views) - n_tail
    batch = vw.NativeViewBatch(views, W, H, targets=[target] * len(views), lanes=lanes)
    small = torch.zeros(10 * n_param, dtype=torch.float32, device=device)     # the tail views' own bucket
    loss = torch.zeros(1, device=device)
    comm = torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    secA, secB = _sections(bucket, n_param), _sections(small, n_param)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    head_done = torch.cuda.Event()
    head_done.record(cur)                  # materialises the CUDA event the native call records in mid-batch

    def step(marks=None):
        bucket.zero_()
        small.zero_()
        loss.zero_()
        # ONE native call for all views of this rank; the views in front of the tail add into the bucket, whose
        # all-reduce starts on `head_done` — recorded inside the batch — beside the tail views
        batch.step(*secA, loss, tail=(first_tail, secB, head_done))
        if world > 1:
            comm.wait_event(head_done)
            with torch.cuda.stream(comm):
                dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
        if marks is not None:
            marks[0].record(cur)
        if world > 1:
            dist.all_reduce(small, op=dist.ReduceOp.SUM)
            cur.wait_stream(comm)
        bucket[:10 * n_param] += small

    def checked_step():
        for _ in range(3):
            step()
            torch.cuda.synchronize()
            if batch.finish():
                return
        raise RuntimeError("the view batch did not fit its arenas after two enlargements")

    checked_step()     # sizes the arenas; the first collective of a process also sets up NCCL's channels (untimed)
    checked_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 3
    t0, t1, t2 = ev(), ev(), ev()
    step_ms = tail_ms = host_ms = 0.0
    for _ in range(steps):
        t0.record()
        h0 = time.perf_counter()
        step(marks=[t1])
        host_ms += (time.perf_counter() - h0) * 1e3
        t2.record()
        torch.cuda.synchronize()
        assert batch.finish()
        step_ms += t0.elapsed_time(t2)
        tail_ms += t1.elapsed_time(t2)
    print(f"multi-view step (native, lanes={lanes}): {step_ms / steps:.2f} ms, host enqueue {host_ms / steps:.2f} ms, "
          f"after the last view {tail_ms / steps:.3f} ms, loss {float(loss):.5f}", file=sys.stderr)
    tot_e, max_ms = vw.aggregate_throughput(elems, step_ms / steps, device)
    _, max_tail = vw.aggregate_throughput(0, tail_ms / steps, device)
    launches = batch.launches
    return {"views": args.views, "views_per_rank": len(mine), "distinct_views_per_rank": len(views), "lanes": lanes,
            "step_ms": max_ms, "host_enqueue_ms": host_ms / steps,
            "exposed_after_last_view_ms": max_tail,
            "tail_views": n_tail,
            "collective": (f"nccl all_reduce(sum): the bucket of the first {first_tail} views on a side stream during the "
                           f"last {n_tail} (event recorded inside the batch), then the 10 floats per Gaussian of those")
            if world > 1 else "none (1 rank)",
            "bucket_bytes": bucket.numel() * 4, "tail_bucket_bytes": small.numel() * 4, "loss": "mean squared error",
            "elements_per_step": tot_e, "Gelem_s": tot_e / (max_ms * 1e-3) / 1e9 if max_ms else None,
            "gpu_launches_per_step": launches}


def python_loop_leg(args, device, mine, n_param, bucket):
    """The same step driven view by view through the drop-in autograd Function (round 1's loop): host-bound."""
    import torch

    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    W, H = 1920, 1080
    views = [wl.splat_view_device(W, H, n_param, seed=1080 + v, device=device) for v in mine]
    leaves = [[sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
               sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True)] for sc in views]
    gI = torch.rand(H + 1, W + 1, 3, device=device) + 0.1
    Wt, Ht = torch.tensor(W), torch.tensor(H)
    sec = _sections(bucket, n_param)

    def one(i):
        sc, (m, lam, o, l) = views[i], leaves[i]
        nx = views[(i + 1) % len(views)]
        compositor.plan_view(nx.boxsize, nx.startpoint, nx.endpoint, W, H)
        for t_ in (m, lam, o, l):
            t_.grad = None
        img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, Wt, Ht)
        img.backward(gI)
        idx = sc.index.long()
        sec[0].index_add_(0, idx, m.grad)
        sec[1].index_add_(0, idx, lam.grad.reshape(-1, 4))
        sec[2].index_add_(0, idx, o.grad.reshape(-1))
        sec[3].index_add_(0, idx, l.grad)

    for i in range(len(views)):
        one(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bucket.zero_()
    a.record()
    h0 = time.perf_counter()
    for i in range(len(views)):
        one(i)
    host_ms = (time.perf_counter() - h0) * 1e3
    b.record()
    torch.cuda.synchronize()
    return {"views": len(views), "step_ms": a.elapsed_time(b), "ms_per_view": a.elapsed_time(b) / len(views),
            "host_enqueue_ms": host_ms}


def c2_leg(device):
    import torch

    import grouped_cumprod as gc
    from simplegaussiansplat_tk71_b200 import compositor
    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    views = wl.bundled_views(device)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    per_view = []
    for sc in views:
        m, lam, o, l = (sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
                        sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True))
        gI = torch.rand(sc.height + 1, sc.width + 1, 3, device=device) + 0.1
        by_route = {}
        default_route = compositor.ROUTE
        for r in ("lists", "tiles"):
            compositor.ROUTE = r
            ts = []
            for i in range(5):
                for t_ in (m, lam, o, l):
                    t_.grad = None
                a, b = ev(), ev()
                a.record()
                img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, sc.width,
                              sc.height)
                img.backward(gI)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts = sorted(ts[1:])
            by_route[r] = ts[len(ts) // 2]
        compositor.ROUTE = default_route
        per_view.append({"view": sc.name, "elements": sc.elements, "gaussians": sc.n, "route": default_route,
                         "splat_ms": by_route[default_route], "splat_ms_by_route": by_route})
        del m, lam, o, l, gI, img
    # the scan ops on view 0's sorted element list
    sc = views[0]
    default_route, compositor.ROUTE = compositor.ROUTE, "lists"
    _, v = compositor._render_forward(sc.boxsize, sc.startpoint, sc.endpoint, sc.mean.float(), sc.lam, sc.opacity,
                                      sc.l_d, sc.width, sc.height)
    compositor.ROUTE = default_route
    x, key = v.x_s, v.key_s
    n = x.numel()
    k = int(torch.unique_consecutive(key).numel())
    y = torch.empty_like(x)
    g = torch.rand_like(x)
    gin = torch.empty_like(x)
    nolen = torch.empty(0, dtype=torch.int32, device=device)
    for _ in range(3):
        gc.grouped_cumprod_forward(x, key, y)
        gc.grouped_cumprod_backward(x, y, g, key, gin, nolen)
    reps = 10
    e0, e1, e2 = ev(), ev(), ev()
    tf = tb = 0.0
    for _ in range(reps):
        e0.record()
        gc.grouped_cumprod_forward(x, key, y)
        e1.record()
        gc.grouped_cumprod_backward(x, y, g, key, gin, nolen)
        e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1)
        tb += e1.elapsed_time(e2)
    tf, tb = tf / reps, tb / reps
    peak, _ = _peaks()
    ab = wl.algorithmic_bytes(n, k)
    return {"views": per_view,
            "scan_on_view0": {"elements": n, "pixel_lists": k, "mean_list_length": n / max(k, 1), "fwd_ms": tf,
                              "bwd_ms": tb, "Gelem_s": n / ((tf + tb) * 1e-3) / 1e9,
                              "frac_of_measured_hbm": ab["fwd_bwd"] / ((tf + tb) * 1e-3) / 1e9 / peak}}


def _median_ms(fn, reps, warm=2, flush=None):
    """Median device time of fn() over `reps` calls (CUDA events on the current stream, one pair per call)."""
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def ref_cuda_ops_leg(device, e_c3):
    """The reference's own CUDA ops (oracle/_ref/grouped_cumprod_ref.so: its four native sources compiled unchanged
    for sm_100a — thrust::inclusive_scan_by_key x2, grouped_cumprod_forward.cu:17-23 / grouped_cumsum_forward.cu:17-23,
    and the O(sum L^2) backward kernel, grouped_cumprod_backward.cu:9-41) timed beside ours on the SAME device
    arrays with the same CUDA-event harness: C1, C3, and C4 with its deep segments capped at 16 Ki elements (the
    reference backward walks every segment tail per element: the uncapped 262 144-element segments alone would
    take minutes).  L2-resident C1 is timed with an L2 flush before every call."""
    import torch

    import grouped_cumprod as ours
    from oracle import ref_function as rf
    from simplegaussiansplat_tk71_b200 import workloads as wl

    ref = rf.reference_ops()
    if ref is None:
        return {"unavailable": "oracle/_ref/grouped_cumprod_ref.so not built"}
    flush = torch.empty(64 << 20, dtype=torch.float32, device=device)
    out = {"harness": "median of per-call CUDA-event pairs on the default stream, same arrays for both; the "
                      "reference's thrust calls include their cudaMalloc/cudaFree + host sync (that is the op)"}

    def one(e, reps, use_flush, ref_bwd_reps):
        y, s, gin = (torch.empty_like(e.x) for _ in range(3))
        fl = flush if use_flush else None
        r = {"elements": e.n, "segments": e.k, "l2_flush": bool(use_flush)}
        for name, mod in (("ours", ours), ("reference", ref)):
            rb = reps if name == "ours" else ref_bwd_reps
            f = _median_ms(lambda: mod.grouped_cumprod_forward(e.x, e.key, y), reps, flush=fl)
            c = _median_ms(lambda: mod.grouped_cumsum_forward(e.grad_out, e.key, s), reps, flush=fl)
            b = _median_ms(lambda: mod.grouped_cumprod_backward(e.x, y, e.grad_out, e.inv, gin, e.seg_end), rb,
                           warm=1, flush=fl)
            r[name] = {"fwd_ms": f, "cumsum_ms": c, "bwd_ms": b, "fwd_bwd_Gelem_s": e.n / ((f + b) * 1e-3) / 1e9}
        r["speedup"] = {k: r["reference"][k] / r["ours"][k] for k in ("fwd_ms", "cumsum_ms", "bwd_ms")}
        return r

    out["c1"] = one(wl.c1(device), 15, True, 15)
    out["c3"] = one(e_c3, 7, False, 3)
    try:
        L, is_deep = wl.lengths_c4(3840 * 2160, 2160, deep=512, deep_lo=8192, deep_hi=16384)
        import numpy as np

        e4 = wl.build("C4 3840x2160 lognormal(ln30,1)+512 deep segments capped at 16Ki", L, 3840, 2160, 2160, device,
                      alpha_scale_per_seg=np.where(is_deep, 1e-3, 1.0))
        out["c4_capped"] = one(e4, 5, False, 2)
        del e4
    except Exception as ex:  # noqa: BLE001
        out["c4_capped"] = {"error": str(ex)[:200]}
    del flush
    torch.cuda.empty_cache()
    return out


def reference_function_leg(device):
    """The reference's own compositor Function (baseline/_ref/gs_model.py:477-820, unmodified; loader
    oracle/ref_function.py) forward + backward on the GPU: (a) with its own CUDA ops, (b) with this repo's drop-in
    `grouped_cumprod` module behind the same Python, (c) the native compositor — same scenes, same harness."""
    import torch

    from oracle import ref_function as rf
    from simplegaussiansplat_tk71_b200 import workloads as wl
    from simplegaussiansplat_tk71_b200.compositor import custom_autograd_grouped_cumprod as F

    if not rf.available():
        return {"unavailable": "baseline/_ref/gs_model.py not shipped"}
    if rf.reference_ops() is None:
        return {"unavailable": "oracle/_ref/grouped_cumprod_ref.so not built"}
    out = {}
    scenes = [("c3_1080p", lambda: wl.splat_view(1920, 1080, 1_000_000, seed=1080, device=device)),
              ("c3_sixteenth", lambda: wl.splat_view(480, 270, 62_500, seed=1080, device=device)),
              ("c2_bundled_view0", lambda: wl.bundled_views(device, n_views=1)[0])]
    for tag, make in scenes:
        try:
            sc = make()
            scene = (sc.boxsize, sc.startpoint, sc.endpoint, sc.mean, sc.lam, sc.opacity, sc.l_d)
            gI = torch.rand(sc.height + 1, sc.width + 1, 3, device=device) + 0.1
            r = {"workload": sc.name, "elements": sc.elements, "gaussians": sc.n}
            res = {}
            for ops_name in ("ref", "dropin"):
                def go(ops_name=ops_name):
                    res[ops_name] = rf.run(ops_name, scene, sc.width, sc.height, gI)
                r["reference_function_ms" if ops_name == "ref" else "reference_function_with_dropin_ms"] = \
                    _median_ms(go, 3, warm=1)
            m, lam, o, l = (sc.mean.float().requires_grad_(True), sc.lam.clone().requires_grad_(True),
                            sc.opacity.clone().requires_grad_(True), sc.l_d.clone().requires_grad_(True))

            def native():
                for t_ in (m, lam, o, l):
                    t_.grad = None
                img = F.apply(sc.boxsize, torch.tensor([sc.n]), sc.startpoint, sc.endpoint, m, lam, o, l, sc.width,
                              sc.height)
                img.backward(gI)
                res["native"] = img.detach()
            r["native_ms"] = _median_ms(native, 7, warm=2)
            r["speedup_vs_reference_function"] = r["reference_function_ms"] / r["native_ms"]
            ia, ib, ic = res["ref"][0], res["dropin"][0], res["native"]
            r["image_sums_ref_dropin_native"] = [float(ia.double().sum()), float(ib.double().sum()), float(ic.double().sum())]
            r["max_abs_image_diff_dropin_vs_ref_ops"] = float((ia - ib).abs().max())
            r["max_abs_image_diff_native_vs_ref_function"] = float((ia - ic).abs().max())
            out[tag] = r
            del res, sc, scene, gI, m, lam, o, l
        except Exception as ex:  # noqa: BLE001
            out[tag] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        torch.cuda.empty_cache()
    return out


def sweep(args, e, y, gin, gc, ops):
    """Times both paths of both ops on the resident workload — aligned, sliced by one element with every array in
    the same 16-byte phase (the alignment peel of the persistent kernels), and sliced with mixed phases (the
    plain-load fallback); prints a table to stderr and JSON to stdout."""
    import torch

    from simplegaussiansplat_tk71_b200 import workloads as wl

    peak, _ = _peaks()
    res = {"workload": e.name, "n": e.n, "rows": []}

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.steps

    n1 = e.n - 4
    layouts = {"aligned": (0, 0, 0, 0, 0), "sliced, same phase (+1 element)": (1, 1, 1, 1, 1),
               "sliced, same phase (+3)": (3, 3, 3, 3, 3), "sliced, mixed phases": (1, 2, 3, 0, 1)}
    gc.grouped_cumprod_forward(e.x, e.key, y)
    for lname, (ox, ok, og, oi, oo) in layouts.items():
        x, key, g, inv = e.x[ox:ox + n1], e.key[ok:ok + n1], e.grad_out[og:og + n1], e.inv[oi:oi + n1]
        yy, gg = y[oo:oo + n1], gin[oo:oo + n1]
        ab = wl.algorithmic_bytes(n1, e.k)
        for v, name in enumerate(ops.variants("fwd")):
            if v == 0 and lname.startswith("sliced, same"):
                continue
            ops.set_variant("fwd", v)
            ms = timeit(lambda: gc.grouped_cumprod_forward(x, key, yy))
            row = {"op": "fwd", "layout": lname, "variant": name.split(" ")[0], "launches": ops.last_launch_count(),
                   "ms": ms, "GBs": ab["fwd"] / ms / 1e6, "frac": ab["fwd"] / ms / 1e6 / peak, "status": ops.workspace_status()}
            res["rows"].append(row)
            print("fwd {layout:34s} {variant:10s} launches {launches} {ms:8.4f} ms {GBs:8.1f} GB/s {frac:6.3f} status {status}".format(**row), file=sys.stderr)
        ops.set_variant("fwd", -1)
        for v, name in enumerate(ops.variants("bwd")):
            if v == 0 and lname.startswith("sliced, same"):
                continue
            ops.set_variant("bwd", v)
            ms = timeit(lambda: gc.grouped_cumprod_backward(x, yy, g, inv, gg, e.seg_end))
            row = {"op": "bwd", "layout": lname, "variant": name.split(" ")[0], "launches": ops.last_launch_count(),
                   "ms": ms, "GBs": ab["bwd"] / ms / 1e6, "frac": ab["bwd"] / ms / 1e6 / peak, "status": ops.workspace_status()}
            res["rows"].append(row)
            print("bwd {layout:34s} {variant:10s} launches {launches} {ms:8.4f} ms {GBs:8.1f} GB/s {frac:6.3f} status {status}".format(**row), file=sys.stderr)
        ops.set_variant("bwd", -1)
    # a plain device copy of the same byte volume as one forward op (sanity ceiling)
    src = torch.empty(e.n * 3 // 2, dtype=torch.float32, device=e.x.device)
    dst = torch.empty_like(src)
    ms = timeit(lambda: dst.copy_(src))
    res["copy"] = {"ms": ms, "GBs": src.numel() * 8 / ms / 1e6}
    print(f"torch copy_ {src.numel() * 8 / 1e6:.0f} MB r+w: {ms:.4f} ms {res['copy']['GBs']:.1f} GB/s", file=sys.stderr)
    print(json.dumps(res), file=_OUT, flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the resample hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  (N > 1: launched by torch.distributed.run, one rank per GPU; images are independent, so every rank
   resamples its own full batch -- weak scaling, no data-path collective.)

A "step" is one pass of the hot path over one batch of synthetic BGRA frames that are already resident
in HBM (`value`), and -- for `e2e` -- the same call sequence through the drop-in C ABI with HOST buffers,
host<->device copies inside the timed region.  `--impl reference` times the CPU oracle (a restatement of
the reference's algorithm; the Rust reference itself cannot be built offline) on the box's host cores.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json metric line: "Mpixels/sec (4K->512 Robidoux batch)"; configs[1] shape with the metric's filter
    "c2_4k_to_512_robidoux": dict(in_wh=(3840, 2160), out_wh=(512, 512), filter=2, batch=1024, alpha=0, compose=0, sharpen=0.0, cm=None),
    # configs[1] verbatim (Lanczos3)
    "c2_4k_to_512_lanczos3": dict(in_wh=(3840, 2160), out_wh=(512, 512), filter=6, batch=1024, alpha=0, compose=0, sharpen=0.0, cm=None),
    # configs[2]: 8K -> 1080p Robidoux, linear-light round trip, sharpen_percent=50 (the reference's only sharpening)
    "c3_8k_to_1080p_robidoux_sharpen": dict(in_wh=(7680, 4320), out_wh=(1920, 1080), filter=2, batch=256, alpha=0, compose=0, sharpen=50.0, cm=None),
    # configs[3]: 1080p -> 4K Mitchell upscale, sepia colour matrix, composited over an existing canvas
    "c4_1080p_to_4k_mitchell_sepia_over": dict(in_wh=(1920, 1080), out_wh=(3840, 2160), filter=14, batch=512, alpha=1, compose=1, sharpen=0.0, cm=0),
}
DEFAULT_WORKLOAD = "c2_4k_to_512_robidoux"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override images per GPU per step")
    ap.add_argument("--alpha", type=int, default=-1, help="override alpha_meaningful (0/1)")
    ap.add_argument("--content", default="noise", choices=["noise", "gradient"])
    ap.add_argument("--e2e-images", type=int, default=48, help="host-buffer images per e2e step")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU-baseline sample time")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--strip-cols", type=int, default=0, help="ring kernel: widest strip of output columns per warp (0 = library default)")
    ap.add_argument("--min-items", type=int, default=-1, help="ring kernel: band split target (-1 = library default)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.first = [], None, index, 0

    def mark(self):
        """samples taken from now on belong to the timed region"""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows[max(0, self.first - 1):]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads() -> int:
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so ask the scheduler instead)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def algorithmic_bytes_per_image(wl):
    iw, ih = wl["in_wh"]; ow, oh = wl["out_wh"]
    b = iw * ih * 4 + ow * oh * 4                    # read input once + write output once
    if wl["compose"] == 1:
        b += ow * oh * 4                             # + canvas read for compose-onto-canvas
    return b


# --------------------------------------------------------------------------------------------------
def cpu_oracle_run(wl, n_images, threads, content, alpha, keep_outputs=False):
    """Times the CPU oracle (OpenMP over images) on n_images frames of the workload. Returns (seconds, outputs)."""
    import oracle
    from imageflow_b200 import synth
    iw, ih = wl["in_wh"]; ow, oh = wl["out_wh"]
    L = oracle.lib()
    ins, outs, descs = [], [], (oracle.Desc * n_images)()
    cm = oracle.color_filter_matrix(wl["cm"]) if wl["cm"] is not None else None
    keep = []
    distinct = min(n_images, max(2, min(16, n_images)))      # generating frames in numpy is slow: cycle through a few
    for i in range(n_images):
        if i < distinct:
            a = synth.noise_np(iw, ih, seed=i, alpha_mode="mixed" if alpha else "opaque") if content == "noise" else synth.gradient_np(iw, ih)
        else:
            a = ins[i % distinct]
        c = synth.noise_np(ow, oh, seed=100000 + i, alpha_mode="mixed") if (wl["compose"] == 1 and i < distinct) else \
            (outs[i % distinct].copy() if wl["compose"] == 1 else np.zeros((oh, ow, 4), np.uint8))
        ins.append(a); outs.append(c)
        descs[i] = oracle.make_desc(a, c, filter=wl["filter"], sharpen=wl["sharpen"], linear=True, alpha_meaningful=bool(alpha),
                                    compose=wl["compose"], color_matrix=cm, keep=keep)
    t0 = time.perf_counter()
    rc = L.ifo_scale_and_render_batch(descs, n_images, threads)
    dt = time.perf_counter() - t0
    if rc:
        raise RuntimeError(f"oracle failed rc={rc}")
    return dt, (outs if keep_outputs else None)


def run_reference(args, wl, rank, world):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    alpha = wl["alpha"] if args.alpha < 0 else args.alpha
    iw, ih = wl["in_wh"]
    # size the bounded sample: one probe image per thread, then scale to ~cpu_seconds / (steps+warmup)
    cpu_oracle_run(wl, threads, threads, args.content, alpha)                     # warm-up (tables, page faults)
    dt, _ = cpu_oracle_run(wl, 2 * threads, threads, args.content, alpha)
    budget = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    n = int(max(2 * threads, min(4096, round(budget / max(dt, 1e-3) * 2 * threads))))
    for _ in range(args.warmup):
        cpu_oracle_run(wl, n, threads, args.content, alpha)
    t = 0.0
    for _ in range(args.steps):
        d, _ = cpu_oracle_run(wl, n, threads, args.content, alpha)
        t += d
    mpx = n * args.steps * iw * ih / 1e6 / t
    sample = f"{n} frames of {iw}x{ih} per step (bounded sample of the {wl['batch']}-frame batch), OpenMP over images"
    line = {"impl": "reference", "metric": "input Mpixels/s, " + args.workload, "value": mpx, "unit": "Mpx/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, alpha, n),
            "cpu_baseline": {"value": mpx, "unit": "Mpx/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mpx, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of the reference algorithm (oracle/ifb_oracle.c); the Rust reference (zenresize) cannot be built offline"}
    print(json.dumps(line), flush=True)


def config_dict(args, wl, alpha, batch):
    iw, ih = wl["in_wh"]; ow, oh = wl["out_wh"]
    names = {2: "Robidoux", 6: "Lanczos", 14: "Mitchell"}
    return {"workload": args.workload, "in": f"{iw}x{ih}", "out": f"{ow}x{oh}", "filter": names.get(wl["filter"], str(wl["filter"])),
            "colorspace": "linear", "alpha_meaningful": bool(alpha), "compose": ["ReplaceSelf", "BlendWithSelf", "BlendWithMatte"][wl["compose"]],
            "sharpen_percent": wl["sharpen"], "color_matrix": "sepia" if wl["cm"] == 0 else None,
            "images_per_gpu_per_step": batch, "content": args.content, "parallelism": f"images sharded x{args.gpus}, no collective",
            "cache": "inputs (>= 8 GB per step) are far larger than the 126 MB L2; no explicit flush needed"}


# --------------------------------------------------------------------------------------------------
def main():
    args = parse()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import imageflow_b200 as ifb
    from imageflow_b200 import synth

    if not torch.cuda.is_available() or ifb.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    alpha = wl["alpha"] if args.alpha < 0 else args.alpha
    B = args.batch or wl["batch"]
    iw, ih = wl["in_wh"]; ow, oh = wl["out_wh"]

    # ---- synthetic inputs, resident in HBM
    free, _total = torch.cuda.mem_get_info()
    need = B * (iw * ih * 4 + 2 * ow * oh * 4)
    if need > free * 0.9:
        B = max(1, int(free * 0.9 // (iw * ih * 4 + 2 * ow * oh * 4)))
    inp = torch.empty((B, ih, iw, 4), dtype=torch.uint8, device=dev)
    for i in range(B):
        if args.content == "noise":
            synth.noise_torch(iw, ih, seed=i, alpha_mode="mixed" if alpha else "opaque", device=dev, out=inp[i])
        else:
            synth.gradient_torch(iw, ih, device=dev, out=inp[i])
    canvas0 = None
    if wl["compose"] == 1:
        canvas0 = torch.empty((B, oh, ow, 4), dtype=torch.uint8, device=dev)
        for i in range(B):
            synth.noise_torch(ow, oh, seed=100000 + i, alpha_mode="mixed", device=dev, out=canvas0[i])
    out = torch.zeros((B, oh, ow, 4), dtype=torch.uint8, device=dev)
    cm = ifb.color_filter_matrix(wl["cm"]) if wl["cm"] is not None else None

    batch = ifb.Batch(local)
    if args.strip_cols:
        batch.set_option(ifb.Batch.OPT_STRIP_COLUMNS, args.strip_cols)
    if args.min_items >= 0:
        batch.set_option(ifb.Batch.OPT_MIN_ITEMS, args.min_items)
    params = ifb.ScaleAndRenderParams(w=ow, h=oh, sharpen_percent_goal=wl["sharpen"], interpolation_filter=ifb.Filter(wl["filter"]))
    jobs = [(ifb.BitmapWindow.from_torch(inp[i], alpha_meaningful=bool(alpha)),
             ifb.BitmapWindow.from_torch(out[i], compose=ifb.BitmapCompositing(wl["compose"])), params, cm) for i in range(B)]
    descs, keep = batch.make_descs(jobs)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        if canvas0 is not None:
            out.copy_(canvas0)          # the composite reads the canvas: restore it so every step does identical work
        batch.enqueue(descs, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before the warm-up so that nvidia-smi's start-up cost is not in the timed region
    t_w = time.perf_counter()
    n_w = 0
    while n_w < max(args.warmup, 3) or time.perf_counter() - t_w < 0.5:     # >= 3 steps and >= 0.5 s: clocks ramp up
        step()
        n_w += 1
        if n_w % 8 == 0:
            torch.cuda.synchronize()
    barrier()
    launches0 = batch.kernel_launches
    if rank == 0:
        sampler.mark()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    host_ms = []
    e0.record()
    for k in range(args.steps):
        if canvas0 is not None:
            out.copy_(canvas0)
        k_ev[k][0].record()
        t_h = time.perf_counter()
        batch.enqueue(descs, stream)
        host_ms.append((time.perf_counter() - t_h) * 1e3)
        k_ev[k][1].record()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = e0.elapsed_time(e1)
    step_ms = [a.elapsed_time(b) for a, b in k_ev]
    kern_ms = float(np.mean(step_ms))
    launches = batch.kernel_launches - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * B * args.steps * iw * ih / 1e6 / (total_ms_max / 1e3)

    # ---- roofline of the dominant kernel (one fused launch per step)
    peak, peak_src = peaks()
    alg = algorithmic_bytes_per_image(wl) * B
    achieved = alg / (kern_ms / 1e3) / 1e9
    tap_flops = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms": kern_ms, "kernel_ms_min": float(np.min(step_ms)),
                "kernel_ms_all": [round(x, 4) for x in step_ms], "host_enqueue_ms": [round(x, 3) for x in host_ms], "algorithmic_bytes_per_launch": alg,
                "fused_jobs": batch.fused_jobs, "generic_jobs": batch.generic_jobs, "tile_jobs": batch.tile_jobs}
    tfile = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(tfile):
        try:
            tj = json.load(open(tfile))
            roofline["traffic"] = tj["dram_bytes_per_launch"] * (B / tj["images_per_launch"])
            roofline["traffic_source"] = tj.get("source")
        except Exception:
            pass

    # ---- parity spot check against the oracle + CPU baseline (rank 0, N == 1 only for the baseline)
    check = None
    cpu_baseline = None
    if rank == 0 and not args.no_check:
        import oracle
        oracle.build()
        n_chk = 2
        _, outs = cpu_oracle_run(wl, n_chk, min(n_chk, os.cpu_count() or 1), args.content, alpha, keep_outputs=True)
        if canvas0 is not None:       # rerun once from the pristine canvas so that out holds exactly one composite
            out.copy_(canvas0)
            batch.enqueue(descs, stream)
            torch.cuda.synchronize()
        mx = 0
        for i in range(n_chk):
            d = np.abs(out[i].cpu().numpy().astype(np.int16) - outs[i].astype(np.int16))
            mx = max(mx, int(d.max()))
        check = {"images": n_chk, "max_abs_delta_vs_oracle": mx}
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        threads = host_threads()
        cpu_oracle_run(wl, threads, threads, args.content, alpha)                 # warm-up (tables, page faults)
        dt, _ = cpu_oracle_run(wl, 2 * threads, threads, args.content, alpha)
        n = int(max(2 * threads, min(4096, round(args.cpu_seconds / max(dt, 1e-3) * 2 * threads))))
        dt, _ = cpu_oracle_run(wl, n, threads, args.content, alpha)
        cpu_baseline = {"value": n * iw * ih / 1e6 / dt, "unit": "Mpx/s", "cores": threads, "kind": "port",
                        "sample": f"{n} frames of {iw}x{ih} ({dt:.1f} s), oracle/ifb_oracle.c OpenMP over images"}

    # ---- e2e: drop-in C ABI with pinned HOST buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        ne = max(1, min(args.e2e_images, B))
        h_in = torch.empty((ne, ih, iw, 4), dtype=torch.uint8).pin_memory()
        h_in.copy_(inp[:ne])
        h_out = torch.zeros((ne, oh, ow, 4), dtype=torch.uint8).pin_memory()
        if canvas0 is not None:
            h_cv0 = canvas0[:ne].cpu()
        os.environ["IFB200_DEVICE"] = str(local)
        hjobs = [(ifb.BitmapWindow(h_in[i].data_ptr(), iw, ih, iw * 4, alpha_meaningful=bool(alpha)),
                  ifb.BitmapWindow(h_out[i].data_ptr(), ow, oh, ow * 4, compose=ifb.BitmapCompositing(wl["compose"])), params) for i in range(ne)]

        def e2e_step():
            if canvas0 is not None:
                h_out.copy_(h_cv0)
            ifb.scale_and_render_many([(wi, wc, p, cm) for (wi, wc, p) in hjobs])

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        h2d = ne * iw * ih * 4 + (ne * ow * oh * 4 if wl["compose"] == 1 else 0)
        e2e = {"value": world * ne * args.steps * iw * ih / 1e6 / dt, "unit": "Mpx/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": ne * ow * oh * 4, "images_per_step": ne,
               "api": "ifb200_scale_and_render_many (host buffers; uploads/kernels/downloads pipelined on 3 streams; returns when all results are in host memory)"}
        if rank == 0 and check is not None and canvas0 is None:
            mx = int(np.abs(h_out[0].numpy().astype(np.int16) - out[0].cpu().numpy().astype(np.int16)).max())
            check["e2e_vs_device_max_abs_delta"] = mx

    if rank == 0:
        line = {"metric": "input Mpixels/s, " + args.workload, "value": value, "unit": "Mpx/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args, wl, alpha, B),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "parity_check": check, "out_mpx_per_s": world * B * args.steps * ow * oh / 1e6 / (total_ms_max / 1e3)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

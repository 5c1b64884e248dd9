#!/usr/bin/env python
"""Benchmark of the FovVideoVDP hot path: 4K test+reference frame pairs scored per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--size WxH]

A *step* is one predict() over one synthetic clip: F (default 64) frames of 3840x2160 fp32 display-encoded
values per stream, display standard_4k, 30 fps (8-tap temporal window), non-foveated, no heat map
(BASELINE.json configs[2], the configuration the headline metric is quoted on).  With N > 1 (one process per
GPU under torchrun) every rank holds its own 64-frame block of a 64*N-frame clip (+ the 7-frame temporal halo)
and the per-band pooled energies are combined by one NCCL all-reduce before the JOD regression ("weak" scaling).

  value     frames/s with the clip resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e       the same through predict() with the clip in pinned HOST memory: host->device copies of every frame
            and the device->host read of the result are inside the timed region
  roofline  the dominant kernel (largest share of device time) against the measured HBM copy peak:
            algorithmic bytes (2*H*W*C*sizeof(in) per frame pair x frames per launch) / its mean launch time
  cpu_baseline  the numpy oracle (oracle/fvvdp_oracle.py, a port of the reference's algorithm) timed on the host
            cores for a bounded sample of the same workload

`--impl reference` times the reference's algorithm on the host CPU (the oracle port, one worker process per core, each
scoring steady-state frames of the same workload) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FPS = 30


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--size", default="3840x2160")
    ap.add_argument("--display", default="standard_4k")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = self.proc.communicate()[0]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------- CPU arms
def _cpu_worker(job):
    """Score `frames` of the clip with the oracle; the temporal window before the first frame is pre-converted
    (untimed) so that the timed part is the steady-state per-frame work: EOTF of the new frame of both streams,
    temporal FIR, pyramid, CSF, masking, pooling."""
    import numpy as np

    from fovvideovdp_b200.synthetic import synth_pair_numpy
    from oracle import fvvdp_oracle as O

    count, H, W, display = job
    fl = O.filter_len(FPS)
    # a clip of fl-1+count frames whose first fl-1 frames only feed the temporal window: frames fl-1.. are scored
    t, r = synth_pair_numpy(count + fl - 1, H, W, first_frame=9)
    md = O.metric_data()
    photo = O.photometry_from_preset(display)
    cache = {}
    for k in range(fl - 1):
        cache[(0, k)] = O.frame_luminance(t[0, :, k], photo, md["rgb2y"]["sRGB"])
        cache[(1, k)] = O.frame_luminance(r[0, :, k], photo, md["rgb2y"]["sRGB"])
    times = []
    t0 = time.perf_counter()
    O.predict(t, r, frames_per_second=FPS, display_name=display, frames=range(fl - 1, fl - 1 + count), lum_cache=cache, frame_times=times)
    return time.perf_counter() - t0, times


def cpu_oracle_rate(H, W, display, n_frames, workers):
    """frames/s of the oracle port on `workers` processes, each scoring n_frames steady-state frames."""
    import multiprocessing as mp

    jobs = [(n_frames, H, W, display) for _ in range(workers)]
    if workers == 1:
        res = [_cpu_worker(jobs[0])]
        wall = res[0][0]
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            pool.map(_noop, range(workers))  # start the workers before timing
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            wall = time.perf_counter() - t0
    scored = workers * n_frames
    steady = sum(sum(r[1]) for r in res)
    # whole-pool throughput over the steady-state part: every worker runs concurrently, so the rate is
    # frames / (mean per-worker steady time)
    rate = scored / (steady / workers) if steady > 0 else 0.0
    return rate, wall, scored


def _noop(_):
    return 0


def run_reference(args, W, H):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    try:
        free_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        per_worker_gb = 40 * H * W * 4 / 1e9 + 0.5
        workers = max(1, min(workers, int(free_gb * 0.6 / per_worker_gb)))
    except (ValueError, OSError):
        pass
    per = max(1, args.cpu_frames)
    rates, walls = [], []
    for _ in range(max(1, min(args.steps, 2))):
        rate, wall, scored = cpu_oracle_rate(H, W, args.display, per, workers)
        rates.append(rate)
        walls.append(wall)
    value = sum(rates) / len(rates)
    line = {
        "impl": "reference", "metric": "4K frames/sec (test+ref pair)", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": len(rates), "warmup": 0, "ms_per_step": 1000.0 * sum(walls) / len(walls), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic {W}x{H} fp32 test/ref pair, display={args.display}, {FPS} fps, non-foveated (BASELINE configs[2])",
                   "l2": "cpu arm"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": workers, "kind": "port",
                         "sample": f"{workers} worker processes x {per} steady-state frames of the {W}x{H} workload, numpy oracle port of "
                                   "the reference algorithm (the Python reference itself cannot travel to the GPU box)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args, W, H):
    import torch
    import torch.distributed as dist

    import fovvideovdp_b200 as m
    from fovvideovdp_b200.synthetic import synth_pair_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    fl = 8
    n_total = F * world
    first = rank * F
    halo = min(first, fl - 1)
    # this rank's frames [first - halo, first + F) of the n_total-frame clip, generated on the device
    t, r = synth_pair_torch(F + halo, H, W, dev, first_frame=first - halo)
    fv = m.fvvdp(display_name=args.display, device=dev, shard_frames=world > 1)

    def source(tt, rr):
        return m.fvvdp_video_source_array(tt, rr, FPS, display_photometry=fv.display_photometry, first_frame=first - halo,
                                          total_frames=n_total)

    vs_dev = source(t, r)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(vs, steps, read_result):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        jod = None
        for _ in range(steps):
            jod, st = fv.predict_video_source(vs)
            if read_result:
                jod = float(jod)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), float(jod)

    for _ in range(max(3, args.warmup)):
        fv.predict_video_source(vs_dev)
    fv._ctx.profile(True)
    launches0 = fv._ctx.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    ms, jod = timed(vs_dev, args.steps, False)
    clocks = sampler.stop() if sampler else None
    launches = fv._ctx.launch_count() - launches0 + args.steps  # + the pooling kernel of every step
    prof = fv._ctx.profile_read()
    fv._ctx.profile(False)
    value = n_total * args.steps / (ms / 1000.0)
    info = dict(fv.last_run)

    # dominant kernel -> roofline
    peak, peak_src = measured_peaks()
    top = max(prof.items(), key=lambda kv: kv[1][0])
    total_ms = sum(v[0] for v in prof.values())
    frames_per_launch = F * args.steps / top[1][1]
    alg_bytes = 2.0 * H * W * 4 * frames_per_launch
    dur_s = top[1][0] / top[1][1] / 1000.0
    achieved = alg_bytes / dur_s / 1e9
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms_per_launch": dur_s * 1000.0, "kernel_share_of_device_time": top[1][0] / total_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(traffic_file):
        with open(traffic_file) as f:
            tj = json.load(f)
        if top[0] in tj.get("kernels", {}):
            roofline["traffic"] = tj["kernels"][top[0]].get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")

    e2e = None
    if not args.no_e2e:
        th, rh = t.cpu().pin_memory(), r.cpu().pin_memory()
        vs_host = source(th, rh)
        fv.predict_video_source(vs_host)
        ms_e, jod_e = timed(vs_host, args.e2e_steps, True)
        e2e = {"value": n_total * args.e2e_steps / (ms_e / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": int(fv.last_run["h2d_bytes"]),
               "d2h_bytes_per_step": int(4 * ((fv._ctx.cfg.n_levels - 1) * 2 * n_total + 2)), "ms_per_step": ms_e / args.e2e_steps,
               "jod": jod_e, "host_memory": "pinned"}
        del th, rh, vs_host

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, wall, scored = cpu_oracle_rate(H, W, args.display, max(1, args.cpu_frames), 1)
        cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
               "sample": f"{scored} steady-state frames of the same {W}x{H} workload (temporal window pre-filled), numpy oracle port, {wall:.1f} s"}

    if rank == 0:
        line = {
            "metric": "4K frames/sec (test+ref pair)", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic {W}x{H} x {F}-frame fp32 test/ref pair per GPU ({n_total} frames total), display={args.display}, "
                                   f"{FPS} fps, non-foveated, replicate padding (BASELINE configs[2]" + (", sharded as configs[3])" if world > 1 else ")"),
                       "frames_per_gpu": F, "block_frames": info.get("block_frames"), "l2": "inputs exceed L2 (2 x %.1f GB per step)" % (F * H * W * 4 / 1e9),
                       "parallelism": f"frame blocks over {world} GPU(s), one all-reduce of the pooled energies" if world > 1 else "single GPU"},
            "jod": jod, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    W, H = [int(v) for v in args.size.lower().split("x")]
    # stdout carries exactly one JSON line: anything libraries print on file descriptor 1 while the benchmark runs
    # (e.g. NCCL's version banner) is sent to stderr instead
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    real_stdout, sys.stdout = sys.stdout, os.fdopen(out_fd, "w")
    try:
        if args.impl == "reference":
            run_reference(args, W, H)
        else:
            run_ours(args, W, H)
    finally:
        sys.stdout.flush()
        sys.stdout = real_stdout


if __name__ == "__main__":
    main()

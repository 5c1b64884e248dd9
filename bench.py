#!/usr/bin/env python
"""Benchmark of the FovVideoVDP hot path: 4K test+reference frame pairs scored per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--size WxH]

A *step* is one predict() over one synthetic clip: F (default 64) frames of 3840x2160 fp32 display-encoded
values per stream, display standard_4k, 30 fps (8-tap temporal window), non-foveated, no heat map
(BASELINE.json configs[2], the configuration the headline metric is quoted on).  With N > 1 (one process per
GPU under torchrun) every rank holds its own 64-frame block of a 64*N-frame clip (+ the 7-frame temporal halo)
and the per-band pooled energies are combined by one NCCL all-reduce before the JOD regression ("weak" scaling).

  value          frames/s with the clip resident in HBM (CUDA events on the launching stream, max over ranks)
  sustained      the same over a timed region of >= 2 s (the K-step region of `value` lasts ~0.1 s: burst clocks)
  e2e            the same through predict() with the clip in pinned HOST memory: host->device copies of every frame
                 and the device->host read of the result are inside the timed region
  roofline       the dominant kernel (largest share of device time) against the measured HBM copy peak:
                 algorithmic bytes (2*H*W*C*sizeof(in) per frame pair x frames per launch) / its mean launch time
  cpu_baseline   the UNMODIFIED reference (pyfvvdp from baseline/_ref, torch CPU, all host threads) on the first 8 frames
                 of the same tensors after a 2-frame warm-up; the numpy oracle port stands in
                 only when baseline/_ref is absent, and the line says so
  reference_cuda the unmodified reference on cuda:0 of the same B200 (TF32 off) on the same 64-frame tensors, with its JOD
  other_configs  BASELINE configs[1] (1080p, standard_fhd) and configs[4] (4K PQ, standard_hdr_pq, foveated, moving gaze)
  strong_256     BASELINE configs[3]: ONE 3840x2160x256-frame clip split over the N ranks (frame block + 7-frame halo each)
  independent_pairs  (N > 1) one independent 64-frame pair per GPU instead of one sharded 64N-frame clip: no halo, no all-reduce

`--impl reference` times the unmodified reference's torch-CPU path (all host threads) on bounded samples of the same
workload and prints the same JSON line with "impl": "reference".
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
sys.path.insert(0, os.path.join(ROOT, "tools"))

FPS = 30
METRIC = "4K frames/sec (test+ref pair)"


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
    ap.add_argument("--cpu-frames", type=int, default=8, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / other_configs / strong_256 / reference_cuda")
    return ap.parse_args()


def workload_string(W, H, F, n_total, display, world):
    return (f"synthetic {W}x{H} x {F}-frame fp32 test/ref pair per GPU ({n_total} frames total), display={display}, {FPS} fps, non-foveated, "
            "replicate padding (BASELINE configs[2]" + (", sharded as configs[3])" if world > 1 else ")"))


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------- CPU arms
def load_reference():
    """The unmodified pyfvvdp package (baseline/_ref, or /root/reference in the build container) or None."""
    try:
        import _refimport

        return _refimport.import_reference(), _refimport.reference_location()
    except ImportError:
        return None, None


def reference_cpu_rate(ref, display, t, r, n_frames, threads):
    """frames/s of pyfvvdp.fvvdp(device='cpu').predict on the first n_frames frames of (t, r)."""
    import torch

    torch.set_num_threads(threads)
    fv = ref.fvvdp(display_name=display, heatmap=None, quiet=True, device=torch.device("cpu"))
    with torch.no_grad():
        t0 = time.perf_counter()
        jod, _ = fv.predict(t[:, :, :n_frames], r[:, :, :n_frames], dim_order="BCFHW", frames_per_second=FPS)
        dt = time.perf_counter() - t0
    return n_frames / dt, dt, float(jod)


def _port_worker(job):
    """numpy oracle port (fallback when the reference package is absent): steady-state frames of the same workload."""
    from fovvideovdp_b200.synthetic import synth_pair_numpy
    from oracle import fvvdp_oracle as O

    count, H, W, display = job
    fl = O.filter_len(FPS)
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


def port_rate(H, W, display, n_frames, workers):
    import multiprocessing as mp

    jobs = [(n_frames, H, W, display) for _ in range(workers)]
    if workers == 1:
        res = [_port_worker(jobs[0])]
        wall = res[0][0]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_noop, range(workers))
            t0 = time.perf_counter()
            res = pool.map(_port_worker, jobs)
            wall = time.perf_counter() - t0
    steady = sum(sum(x[1]) for x in res)
    return (workers * n_frames) / (steady / workers) if steady > 0 else 0.0, wall, workers * n_frames


def _noop(_):
    return 0


def run_reference(args, W, H):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from fovvideovdp_b200.synthetic import synth_pair_numpy

    cores = os.cpu_count() or 1
    ref, where = load_reference()
    F = args.frames
    world = max(1, args.gpus)
    config = {"workload": workload_string(W, H, F, F * world, args.display, world), "l2": "cpu arm"}
    if ref is not None:
        # every step scores the first n frames of the bench clip with the unmodified reference on all host threads; n is
        # sized from the warm-up so that K steps end within a few minutes
        t, r = synth_pair_numpy(8, H, W)
        t, r = torch.from_numpy(t), torch.from_numpy(r)
        warm = []
        for _ in range(max(1, min(args.warmup, 3))):
            warm.append(reference_cpu_rate(ref, args.display, t, r, 2, cores)[1] / 2.0)
        per_frame = min(warm)
        n = int(max(2, min(8, 150.0 / (max(1, args.steps) * per_frame))))
        times, jod = [], None
        for _ in range(max(1, args.steps)):
            _, dt, jod = reference_cpu_rate(ref, args.display, t, r, n, cores)
            times.append(dt)
        value = n * len(times) / sum(times)
        kind = "reference"
        sample = (f"pyfvvdp {getattr(ref, '__version__', '1.2.2')} unmodified from {os.path.relpath(where, ROOT) if where.startswith(ROOT) else where}, "
                  f"fvvdp(device='cpu').predict on the first {n} frames of the {W}x{H} clip per step, torch.set_num_threads({cores}), "
                  f"{len(warm)} warm-up calls of 2 frames")
        steps, ms = len(times), 1000.0 * sum(times) / len(times)
        used = cores
    else:
        workers = max(1, min(cores, 32))
        rates, walls = [], []
        for _ in range(max(1, min(args.steps, 2))):
            rate, wall, _ = port_rate(H, W, args.display, 2, workers)
            rates.append(rate)
            walls.append(wall)
        value, kind, steps, ms, used, jod = sum(rates) / len(rates), "port", len(rates), 1000.0 * sum(walls) / len(walls), workers, None
        sample = (f"baseline/_ref absent (run tools/vendor_reference.py in the build container): numpy oracle port, {workers} worker processes x 2 "
                  f"steady-state frames of the {W}x{H} workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 3)),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "jod": jod,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------- GPU arm
def bind_to_local_cores(local, world):
    """Pin this rank to the cores of the NUMA node its GPU hangs off (so that the pinned host buffers it allocates afterwards
    are first-touched there); when the topology is not exposed, to an even share of the cores the process may use."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return None
    cores = None
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        node_file = f"/sys/bus/pci/devices/{bus[-12:]}/numa_node"
        node = int(open(node_file).read().strip()) if os.path.isfile(node_file) else -1
        if node >= 0:
            cl = []
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                cl.extend(range(int(a), int(b or a) + 1))
            cores = [c for c in cl if c in allowed] or None
    except Exception:
        cores = None
    if cores is None and world > 1:
        share = max(1, len(allowed) // world)
        cores = allowed[local * share:(local + 1) * share] or allowed
    if cores:
        try:
            os.sched_setaffinity(0, cores)
        except OSError:
            return None
    return cores


def run_ours(args, W, H):
    import numpy as np
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
    cores_bound = bind_to_local_cores(local, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from fovvideovdp_b200.fvvdp import frame_block

    F = args.frames
    fl = 8
    n_total = F * world

    def my_block(n_frames):  # the cut predict_video_source() makes (work-balanced: ranks > 0 also walk the 7-frame temporal halo)
        return frame_block(n_frames, rank, world, halo=fl - 1, first_halo=1)

    first, last = my_block(n_total)
    halo = min(first, fl - 1)
    # this rank's frames [first - halo, last) of the n_total-frame clip, generated on the device
    t, r = synth_pair_torch(last - first + halo, H, W, dev, first_frame=first - halo)
    fv = m.fvvdp(display_name=args.display, device=dev, shard_frames=world > 1)

    def source(tt, rr, first_frame=first - halo, total=n_total, metric=fv):
        return m.fvvdp_video_source_array(tt, rr, FPS, display_photometry=metric.display_photometry, first_frame=first_frame, total_frames=total)

    vs_dev = source(t, r)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(metric, vs, steps, read_result, fixation=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        jod = None
        for _ in range(steps):
            jod, st = metric.predict_video_source(vs, fixation_point=fixation)
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
    ms, jod = timed(fv, vs_dev, args.steps, False)
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
    frames_per_launch = (last - first) * args.steps / top[1][1]
    alg_bytes = 2.0 * H * W * 4 * frames_per_launch
    dur_s = top[1][0] / top[1][1] / 1000.0
    achieved = alg_bytes / dur_s / 1e9
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms_per_launch": dur_s * 1000.0, "kernel_share_of_device_time": top[1][0] / total_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
                "host_gap_ms_per_step": ms / args.steps - total_ms / args.steps}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(traffic_file):
        with open(traffic_file) as f:
            tj = json.load(f)
        if top[0] in tj.get("kernels", {}):
            roofline["traffic"] = tj["kernels"][top[0]].get("dram_bytes_per_launch")
            roofline["traffic_source"] = tj.get("source")

    extras = not args.no_extras
    sustained = None
    if extras:
        n_sus = max(args.steps, int(args.sustained_seconds * 1000.0 / (ms / args.steps)) + 1)
        s2 = ClockSampler(local) if rank == 0 else None
        ms_s, _ = timed(fv, vs_dev, n_sus, False)
        c2 = s2.stop() if s2 else None
        sustained = {"value": n_total * n_sus / (ms_s / 1000.0), "unit": "frames/s", "steps": n_sus, "seconds": ms_s / 1000.0, "clocks": c2}

    e2e = None
    if not args.no_e2e:
        th, rh = t.cpu().pin_memory(), r.cpu().pin_memory()
        vs_host = source(th, rh)
        fv.predict_video_source(vs_host)
        ms_e, jod_e = timed(fv, vs_host, args.e2e_steps, True)
        e2e = {"value": n_total * args.e2e_steps / (ms_e / 1000.0), "unit": "frames/s", "h2d_bytes_per_step": int(fv.last_run["h2d_bytes"]),
               "d2h_bytes_per_step": int(4 * ((fv._ctx.cfg.n_levels - 1) * 2 * n_total + 2)), "ms_per_step": ms_e / args.e2e_steps,
               "jod": jod_e, "host_memory": "pinned", "cores_bound": len(cores_bound) if cores_bound else None}
        del th, rh, vs_host

    # ---- BASELINE configs[3]: ONE 256-frame clip split over the ranks (strong scaling, halo included)
    strong = None
    if extras and (W, H) == (3840, 2160):
        N256 = 256
        b0, b1 = my_block(N256)
        h0 = min(b0, fl - 1)
        del vs_dev
        t = r = None
        torch.cuda.empty_cache()
        ts, rs = synth_pair_torch(b1 - b0 + h0, H, W, dev, first_frame=b0 - h0)
        vs256 = source(ts, rs, first_frame=b0 - h0, total=N256)
        fv.predict_video_source(vs256)
        ms_256, jod_256 = timed(fv, vs256, 3, False)
        strong = {"value": N256 * 3 / (ms_256 / 1000.0), "unit": "frames/s", "frames": N256, "frames_per_rank": b1 - b0, "halo_frames": h0,
                  "ms_per_clip": ms_256 / 3, "jod": jod_256, "scaling": "strong"}
        del ts, rs, vs256
        torch.cuda.empty_cache()
        t, r = synth_pair_torch(last - first + halo, H, W, dev, first_frame=first - halo)

    # ---- N > 1: one independent 64-frame pair per GPU (pair-level data parallelism, what the CLI's batch front end does with a list
    #      of pairs): no temporal halo, no all-reduce.  Beside `value` it shows how much of the sharded clip's loss is the halo.
    pairs = None
    if extras and world > 1:
        fvp = m.fvvdp(display_name=args.display, device=dev)
        Fp = min(F, int(t.shape[2]))
        vsp = source(t[:, :, :Fp], r[:, :, :Fp], first_frame=0, total=Fp, metric=fvp)
        for _ in range(3):
            fvp.predict_video_source(vsp)
        ms_p, jod_p = timed(fvp, vsp, args.steps, False)
        pairs = {"value": world * Fp * args.steps / (ms_p / 1000.0), "unit": "frames/s", "ms_per_step": ms_p / args.steps,
                 "what": "one independent %d-frame pair per GPU (no halo, no collective), max over ranks" % Fp, "jod_rank0": jod_p}
        del vsp, fvp

    # ---- the other single-GPU configurations of BASELINE.json
    other = None
    if extras and world == 1:
        other = {}
        # the reference CLI's second metric on the same resident clip: one PU21 kernel launch per 32-frame block
        pu = m.pu_psnr(device=dev, display_name=args.display)
        vs_pu = source(t, r)
        for _ in range(2):
            pu.predict_video_source(vs_pu)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            q_pu, _ = pu.predict_video_source(vs_pu)
        e1.record()
        barrier()
        ms_pu = e0.elapsed_time(e1) / 5
        other["pu-psnr on the configs[2] clip"] = {"value": F / (ms_pu / 1000.0), "unit": "frames/s", "db": float(q_pu), "ms_per_clip": ms_pu,
                                                   "relative_to_fvvdp": ms_pu / (ms / args.steps)}
        del vs_pu
        t2, r2 = synth_pair_torch(F, 1080, 1920, dev)
        fv2 = m.fvvdp(display_name="standard_fhd", device=dev)
        vs2 = source(t2, r2, first_frame=0, total=F, metric=fv2)
        for _ in range(3):
            fv2.predict_video_source(vs2)
        ms2, jod2 = timed(fv2, vs2, 10, False)
        other["configs[1] 1920x1080x%d fp32, standard_fhd" % F] = {"value": F * 10 / (ms2 / 1000.0), "unit": "frames/s", "jod": jod2}
        del t2, r2, vs2, fv2
        # foveated HDR: PQ code values 0.1 + 0.65 v, gaze moving corner to corner (ex_foveated_video.py:36-37)
        tq, rq = 0.1 + 0.65 * t[:, :, halo:halo + F], 0.1 + 0.65 * r[:, :, halo:halo + F]
        fv5 = m.fvvdp(display_name="standard_hdr_pq", device=dev, foveated=True)
        gaze = np.stack([np.linspace(0, W - 1, F), np.linspace(0, H - 1, F)], 1).astype(np.float32)
        vs5 = source(tq, rq, first_frame=0, total=F, metric=fv5)
        for _ in range(3):
            fv5.predict_video_source(vs5, fixation_point=gaze)
        ms5, jod5 = timed(fv5, vs5, 10, False, fixation=gaze)
        other["configs[4] %dx%dx%d fp32 PQ, standard_hdr_pq, foveated, moving gaze" % (W, H, F)] = {"value": F * 10 / (ms5 / 1000.0), "unit": "frames/s", "jod": jod5}
        del tq, rq, vs5, fv5

    # ---- the unmodified reference on the same tensors: CPU (bounded sample) and cuda:0 (TF32 off)
    cpu, ref_cuda = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref, where = load_reference()
        cores = os.cpu_count() or 1
        if ref is not None:
            try:
                os.sched_setaffinity(0, range(cores))
            except (AttributeError, OSError):
                pass
            tc, rc = t[:, :, halo:halo + args.cpu_frames].cpu(), r[:, :, halo:halo + args.cpu_frames].cpu()
            reference_cpu_rate(ref, args.display, tc, rc, 2, cores)
            rate, dt, jod_c = reference_cpu_rate(ref, args.display, tc, rc, args.cpu_frames, cores)
            cpu = {"value": rate, "unit": "frames/s", "cores": cores, "kind": "reference", "jod": jod_c,
                   "sample": f"unmodified pyfvvdp fvvdp(device='cpu').predict, first {args.cpu_frames} frames of the same {W}x{H} tensors after a 2-frame "
                             f"warm-up call, torch.set_num_threads({cores}), {dt:.1f} s"}
            if extras:
                torch.backends.cudnn.allow_tf32 = False
                torch.backends.cuda.matmul.allow_tf32 = False
                fvr = ref.fvvdp(display_name=args.display, heatmap=None, device=dev)
                with torch.no_grad():
                    fvr.predict(t[:, :, halo:halo + 2], r[:, :, halo:halo + 2], dim_order="BCFHW", frames_per_second=FPS)
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    jr, _ = fvr.predict(t[:, :, halo:], r[:, :, halo:], dim_order="BCFHW", frames_per_second=FPS)
                    jr = float(jr)
                    torch.cuda.synchronize(dev)
                    dtr = time.perf_counter() - t0
                ref_cuda = {"value": F / dtr, "unit": "frames/s", "frames": F, "tf32": False, "jod": jr, "jod_rel_err_ours": abs(jod - jr) / abs(jr),
                            "how": "unmodified pyfvvdp fvvdp(device='cuda:0').predict on the same resident 64-frame tensors, wall clock around one call"}
        else:
            rate, wall, scored = port_rate(H, W, args.display, 2, 1)
            cpu = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                   "sample": f"baseline/_ref absent: {scored} steady-state frames of the same {W}x{H} workload, numpy oracle port, {wall:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(W, H, F, n_total, args.display, world),
                       "frames_per_gpu": F, "frames_this_rank": last - first, "block_frames": info.get("block_frames"), "l2": "inputs exceed L2 (2 x %.1f GB per step)" % (F * H * W * 4 / 1e9),
                       "parallelism": f"frame blocks over {world} GPU(s), one all-reduce of the pooled energies" if world > 1 else "single GPU"},
            "jod": jod, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "sustained": sustained, "strong_256": strong, "independent_pairs": pairs, "other_configs": other, "reference_cuda": ref_cuda,
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

"""Batch front end: score many test / reference pairs, pair-level data parallel over the GPUs of one node.

The reference's command line (pyfvvdp/run_fvvdp.py:201-227) walks the pairs one after the other on one device.  Here the
pairs are dealt round-robin to one worker thread per GPU (kernel launches release the GIL; every worker owns its metric
objects, scoring contexts and CUDA stream), or -- under torchrun -- to the ranks, and the results come back in the order
of the input.  Same arguments as the reference CLI where they apply:

    python -m fovvideovdp_b200.run_fvvdp --test a.yuv b.png --ref a_ref.yuv b_ref.png --display standard_4k [--gpus 0 1 2 3]
"""
import argparse
import json
import logging
import os
import sys
import threading

import torch

from .display_model import fvvdp_display_geometry, fvvdp_display_photometry
from .fvvdp import fvvdp
from .pupsnr import pu_psnr
from .video_source_file import fvvdp_video_source_file


def deal_pairs(n_pairs, n_workers):
    """Pair indices of every worker: round robin, so that long and short clips spread evenly."""
    return [list(range(w, n_pairs, n_workers)) for w in range(n_workers)]


def expand_pairs(tests, refs):
    """One reference for every test, one test for every reference, or equally many (run_fvvdp.py:156-167)."""
    if len(tests) == 0 or len(refs) == 0:
        raise RuntimeError("No test / reference images or videos given")
    if len(tests) != len(refs) and len(tests) != 1 and len(refs) != 1:
        raise RuntimeError("Pass the same number of reference and test sources, or a single reference (to be used with all test sources), "
                           "or a single test (to be used with all reference sources).")
    n = max(len(tests), len(refs))
    return [(tests[min(k, len(tests) - 1)], refs[min(k, len(refs) - 1)]) for k in range(n)]


def score_pairs(pairs, display="standard_4k", metrics=("fvvdp",), devices=None, foveated=False, heatmap=None, temp_padding="replicate",
                nframes=-1, full_screen_resize=None, source_factory=None):
    """[(test file, reference file)] -> [{metric short name: (value, stats)}], in input order.

    devices: CUDA device indices to use (default: all visible).  source_factory(test, ref, display_photometry, display_geometry)
    may supply other video sources (e.g. the reference's ffmpeg-based fvvdp_video_source_file)."""
    if not torch.cuda.is_available():
        raise RuntimeError("fovvideovdp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    devices = list(range(torch.cuda.device_count())) if devices is None else list(devices)
    photometry = fvvdp_display_photometry.load(display)
    geometry = fvvdp_display_geometry.load(display)
    results = [None] * len(pairs)
    errors = []

    def make_source(test, ref):
        if source_factory is not None:
            return source_factory(test, ref, photometry, geometry)
        return fvvdp_video_source_file(test, ref, display_photometry=photometry, frames=nframes, full_screen_resize=full_screen_resize,
                                       resize_resolution=geometry.resolution)

    def worker(dev_index, todo):
        try:
            dev = torch.device("cuda", dev_index)
            with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.Stream(device=dev)):
                objs = []
                for mm in metrics:
                    if mm == "fvvdp":
                        objs.append(fvvdp(display_photometry=photometry, display_geometry=geometry, foveated=foveated, heatmap=heatmap,
                                          device=dev, temp_padding=temp_padding))
                    elif mm == "pu-psnr":
                        objs.append(pu_psnr(device=dev, display_photometry=photometry))
                    else:
                        raise RuntimeError(f"Unknown metric {mm}")
                for k in todo:
                    out = {}
                    for obj in objs:
                        q, stats = obj.predict_video_source(make_source(*pairs[k]))
                        out[obj.short_name()] = (float(q), stats)
                    results[k] = out
        except Exception as e:  # reported by the caller, with the pair that failed
            errors.append((dev_index, e))

    todo = deal_pairs(len(pairs), len(devices))
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        # one process per GPU: this rank takes its share, everybody gets everything
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        mine = deal_pairs(len(pairs), world)[rank]
        worker(torch.cuda.current_device(), mine)
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, ([(k, results[k]) for k in mine], [repr(e) for _, e in errors]))
        for part, errs in gathered:
            if errs:
                raise RuntimeError("; ".join(errs))
            for k, v in part:
                results[k] = v
        return results
    threads = [threading.Thread(target=worker, args=(d, t)) for d, t in zip(devices, todo) if t]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    if errors:
        raise errors[0][1]
    return results


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Evaluate FovVideoVDP (B200 core) on a set of images / raw .yuv videos")
    p.add_argument("--test", type=str, nargs="+", required=True, help="list of test images/videos")
    p.add_argument("--ref", type=str, nargs="+", required=True, help="list of reference images/videos")
    p.add_argument("--gpus", type=int, nargs="+", default=None, help="GPUs to spread the pairs over (default: all)")
    p.add_argument("--heatmap", type=str, default="none", help="type of difference map (none, raw, threshold, supra-threshold)")
    p.add_argument("--features", action="store_true", default=False, help="generate JSON files with extracted features")
    p.add_argument("--output-dir", type=str, default=None)
    p.add_argument("--foveated", action="store_true", default=False)
    p.add_argument("--display", type=str, default="standard_4k")
    p.add_argument("--nframes", type=int, default=-1)
    p.add_argument("--full-screen-resize", choices=["bilinear", "bicubic", "nearest", "area"], default=None)
    p.add_argument("--metrics", choices=["fvvdp", "pu-psnr"], nargs="+", default=["fvvdp"])
    p.add_argument("--temp-padding", choices=["replicate", "circular", "pingpong"], default="replicate")
    p.add_argument("--quiet", action="store_true", default=False)
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    logging.basicConfig(format="[%(levelname)s] %(message)s", level=logging.WARNING if args.quiet else logging.INFO)
    pairs = expand_pairs(args.test, args.ref)
    heatmap = None if args.heatmap == "none" else args.heatmap
    res = score_pairs(pairs, display=args.display, metrics=args.metrics, devices=args.gpus, foveated=args.foveated, heatmap=heatmap,
                      temp_padding=args.temp_padding, nframes=args.nframes, full_screen_resize=args.full_screen_resize)
    out_dir = "." if args.output_dir is None else args.output_dir
    os.makedirs(out_dir, exist_ok=True)
    units = {"FovVideoVDP": "JOD", "PU21-PSNR": "dB"}
    for (test, _), r in zip(pairs, res):
        for name, (q, stats) in r.items():
            print("{:0.4f}".format(q) if args.quiet else "{}={:0.4f} [{}]".format(name, q, units[name]))
            base = os.path.splitext(os.path.basename(test))[0]
            if args.features and stats is not None:
                fmap = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in stats.items() if k not in ("Q_per_ch", "heatmap")}
                Q = stats["Q_per_ch"]
                for cc in range(Q.shape[1]):
                    for bb in range(Q.shape[0]):
                        fmap[f"t{cc}_b{bb}"] = Q[bb, cc, :].tolist()
                with open(os.path.join(out_dir, base + "_fmap.json"), "w", encoding="utf-8") as f:
                    json.dump(fmap, f, ensure_ascii=False, indent=4)
            if heatmap is not None and stats is not None and "heatmap" in stats:
                torch.save(stats["heatmap"], os.path.join(out_dir, base + "_heatmap.pt"))
    return 0


if __name__ == "__main__":
    sys.exit(main())

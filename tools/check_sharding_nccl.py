#!/usr/bin/env python
"""torchrun --nproc-per-node G tools/check_sharding_nccl.py: a clip sharded over G GPUs (one NCCL all-reduce of the pooled
energies) must give, on every rank, bit-identical per-frame energies and JOD to the same clip scored on one GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ok = True
for fps, pad in ((30, "replicate"), (60, "pingpong"), (24, "circular")):
    N, H, W = 37, 540, 960
    t, r = synth_pair_torch(N, H, W, dev)
    whole, st_w = m.fvvdp(display_name="standard_fhd", device=dev, temp_padding=pad).predict(t, r, frames_per_second=fps)
    shard, st_s = m.fvvdp(display_name="standard_fhd", device=dev, temp_padding=pad, shard_frames=True).predict(t, r, frames_per_second=fps)
    same = float(whole) == float(shard) and np.array_equal(st_w["Q_per_ch"], st_s["Q_per_ch"])
    ok = ok and same
    print(f"rank {rank}/{world} fps={fps} pad={pad}: whole {float(whole):.6f} sharded {float(shard):.6f} identical={same}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag) == 1 else 1)

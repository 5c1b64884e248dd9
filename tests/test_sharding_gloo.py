"""N>1 host logic on CPU: two gloo ranks shard a clip over frame blocks exactly as fvvdp.predict_video_source does
(frame_block + temporal halo via fvvdp_video_source_array(first_frame=, total_frames=) + one all-reduce of the
zero-initialised per-band energies).  The per-frame scoring itself is done by the CPU oracle here (the CUDA kernels
need a GPU); the pooled result must equal the single-process oracle result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fovvideovdp_b200.fvvdp import frame_block, initial_window
from fovvideovdp_b200.synthetic import synth_pair_numpy
from fovvideovdp_b200.video_source import fvvdp_video_source_array
from oracle import fvvdp_oracle as O

N, H, W, FPS = 9, 64, 96, 30


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, pad, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fl = O.filter_len(FPS)
        a, b = frame_block(N, rank, world)
        first = initial_window(N, fl, pad)
        frame_at = lambda t: t if t >= 1 else first[fl - 1 + t]
        needed = sorted({frame_at(t) for f in range(a, b) for t in range(f - fl + 1, f + 1)})
        lo, hi = needed[0], needed[-1] + 1
        # this rank only holds frames [lo, hi) of the clip
        t, r = synth_pair_numpy(hi - lo, H, W, first_frame=lo)
        vs = fvvdp_video_source_array(t, r, FPS, display_photometry="standard_fhd", first_frame=lo, total_frames=N)
        assert vs.get_video_size() == (H, W, N)
        for k in needed:
            vs.local_index(k)  # every frame of the window is local
        if lo > 0:
            with pytest.raises(RuntimeError):
                vs.get_test_frame(lo - 1)
        # score the block with the oracle on the full clip restricted to this block (same window rule)
        tf, rf = synth_pair_numpy(N, H, W)
        _, st = O.predict(tf, rf, frames_per_second=FPS, display_name="standard_fhd", temp_padding=pad, frames=range(a, b))
        Q = torch.zeros(st["Q_per_ch"].shape, dtype=torch.float32)
        Q[:, :, a:b] = torch.from_numpy(st["Q_per_ch"][:, :, a:b])
        dist.all_reduce(Q)  # every column has exactly one non-zero contributor
        jod = O.pool_to_jod(Q.numpy(), O.metric_data()["parameters"])
        np.save(os.path.join(out_dir, f"q{rank}.npy"), Q.numpy())
        np.save(os.path.join(out_dir, f"jod{rank}.npy"), np.array(jod))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("pad", ["replicate", "circular"])
def test_two_rank_frame_sharding(tmp_path, pad):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), pad, str(tmp_path)), nprocs=world, join=True)
    t, r = synth_pair_numpy(N, H, W)
    want, wst = O.predict(t, r, frames_per_second=FPS, display_name="standard_fhd", temp_padding=pad)
    for rank in range(world):
        Q = np.load(tmp_path / f"q{rank}.npy")
        np.testing.assert_array_equal(Q, wst["Q_per_ch"])
        assert abs(float(np.load(tmp_path / f"jod{rank}.npy")) - want) < 1e-9

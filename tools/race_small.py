import os, sys
sys.path.insert(0, os.getcwd())
import torch
import fovvideovdp_b200 as m
from fovvideovdp_b200.synthetic import synth_pair_torch
dev = torch.device("cuda:0")
td, rd = synth_pair_torch(10, 72, 160, dev)
jod, _ = m.fvvdp(device=dev, display_name="standard_fhd").predict(td, rd, frames_per_second=30)
print(float(jod))

"""fovvideovdp_b200 -- a B200-native (sm_100a CUDA) core for the FovVideoVDP per-frame hot path."""

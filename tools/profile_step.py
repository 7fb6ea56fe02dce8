"""Kernel-time breakdown of one eager RAFT-Stereo forward at the bench shape (torch.profiler, CUDA activities)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nndepth_b200.engine import StereoEngine  # noqa: E402
from nndepth_b200.raft_stereo import BaseRAFTStereo  # noqa: E402

torch.manual_seed(0)
channels_last = "--channels-last" in sys.argv
model = BaseRAFTStereo(iters=32).eval()
model.dense_precision = next((a.split("=")[1] for a in sys.argv if a.startswith("--mode=")), "mixed2x")
engine = StereoEngine(model, use_cuda_graph=False)
if channels_last:
    engine.model = engine.model.to(memory_format=torch.channels_last)
left = torch.rand(8, 3, 375, 1242, device="cuda") * 2 - 1
right = torch.rand(8, 3, 375, 1242, device="cuda") * 2 - 1
for _ in range(2):
    engine.infer_device(left, right)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    engine.infer_device(left, right)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=100))

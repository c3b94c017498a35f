"""Per-role wait counters of the pipelined tensor-core kernels at the benchmark block shapes (MDIL_TC_TRACE=1)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mdil_ss_b200 import erfnet_RA_parallel as M
M.current_task = 0
for C, dil, H, W in ((64, 1, 128, 256), (128, 2, 64, 128), (128, 16, 64, 128)):
    mod = M.non_bottleneck_1d_RAP(C, 0.0, dil, 1).cuda().train()
    x = torch.rand(6, C, H, W, device="cuda").requires_grad_(True)
    y = mod(x)
    y.sum().backward()
    torch.cuda.synchronize()

import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mdil_ss_b200 import erfnet_RA_parallel as M
mod = M.non_bottleneck_1d(16, 0.0, 1).cuda().train()
x = torch.rand(6, 16, 256, 512, device="cuda").requires_grad_(True)
for _ in range(2):
    y = mod(x)
    y.sum().backward()
torch.cuda.synchronize()

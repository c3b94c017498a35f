"""2-rank check of CrossEntropyLoss2d(global_norm=True): the loss equals the full-batch loss and the averaged logit
gradient equals the full-batch gradient (nn.DataParallel semantics).  torchrun --nproc-per-node 2 tools/ce_global_norm_check.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mdil_ss_b200.losses import CrossEntropyLoss2d
from mdil_ss_b200.train_step import class_weights
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(3)
logits = torch.randn(4, 20, 64, 128, generator=g)
labels = torch.randint(0, 20, (4, 64, 128), generator=g)
labels[:2][labels[:2] > 9] = 19            # very different ignore (zero-weight class) fractions on the two shards
w = class_weights("cityscapes", dev)
full = logits.to(dev).requires_grad_(True)
lf = CrossEntropyLoss2d(w)(full, labels.to(dev)); lf.backward()
per = 4 // world
mine = logits[rank * per:(rank + 1) * per].to(dev).requires_grad_(True)
for gn in (False, True):
    mine.grad = None
    l = CrossEntropyLoss2d(w, global_norm=gn)(mine, labels[rank * per:(rank + 1) * per].to(dev)); l.backward()
    gfull = full.grad[rank * per:(rank + 1) * per]
    # the optimiser averages the ranks' parameter gradients: compare world-averaged-equivalent logit gradient
    err = ((mine.grad / world - gfull).abs().max() / gfull.abs().max()).item()
    lerr = abs(float(l) - float(lf)) / abs(float(lf))
    print(f"rank {rank} global_norm={gn}: loss {float(l):.6f} vs full-batch {float(lf):.6f} (rel {lerr:.2e}), grad rel err {err:.2e}", flush=True)
    if gn:
        assert lerr < 1e-6 and err < 1e-5
dist.barrier(); dist.destroy_process_group()

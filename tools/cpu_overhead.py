"""Host-side enqueue time of one training step vs its device time (is the step launch-bound?)."""
import os, sys, time, io, contextlib, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from mdil_ss_b200.erfnet_RA_parallel import Net
from mdil_ss_b200.train_step import Step1Trainer, class_weights
dev = torch.device("cuda", 0)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = Net([20], 1, 0).to(dev)
tr = Step1Trainer(model, class_weights("cityscapes", dev))
x = torch.rand(6, 3, 512, 1024, device=dev)
y = torch.randint(0, 20, (6, 1, 512, 1024), device=dev)
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
t0 = time.perf_counter(); enq = 0.0
for _ in range(10):
    a = time.perf_counter(); tr.step(x, y); enq += time.perf_counter() - a
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(f"host enqueue {enq/10*1e3:.2f} ms/step, wall {tot/10*1e3:.2f} ms/step")
# host time with the GPU idle-waiting removed: enqueue a step after a sync each time
enq2 = 0.0
for _ in range(5):
    torch.cuda.synchronize(); a = time.perf_counter(); tr.step(x, y); enq2 += time.perf_counter() - a
torch.cuda.synchronize()
print(f"host enqueue from idle {enq2/5*1e3:.2f} ms/step")

MDIL_TC_TRACE=1 timeout -s KILL 40 python tools/trace_tc.py 2>&1 | grep -E "wgrad_tc|Error|error" | cut -c1-250 | head -14
echo "--- full net one step"
timeout -s KILL 40 python -c "
import torch, io, contextlib, sys
sys.path.insert(0,'.')
from mdil_ss_b200.erfnet_RA_parallel import Net
from mdil_ss_b200.train_step import Step1Trainer, class_weights
dev=torch.device('cuda',0)
with contextlib.redirect_stdout(io.StringIO()):
    m=Net([20],1,0).to(dev)
tr=Step1Trainer(m,class_weights('cityscapes',dev))
x=torch.rand(6,3,512,1024,device=dev); y=torch.randint(0,20,(6,1,512,1024),device=dev)
for i in range(3):
    l=tr.step(x,y); torch.cuda.synchronize(); print('step',i,float(l))
" 2>&1 | tail -5

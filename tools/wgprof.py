import os, sys
os.environ["SVRS_WG_PROF"]="1"
os.environ["SVRS_WGRAD_STREAM"]="0"; os.environ["SVRS_BRANCH_STREAMS"]="0"
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")): sys.path.insert(0,p)
import torch, models
from svrs_native.trainer import FusedCondTrainer
dev=torch.device("cuda",0)
torch.manual_seed(0)
model=models.Cond_SRVAE(2,64).to(dev).train(); model.set_compute_dtype(torch.bfloat16)
tr=FusedCondTrainer(model)
x=torch.rand(128,4,64,64,device=dev); y=torch.rand(128,4,32,32,device=dev)
tr.step(x,y); torch.cuda.synchronize()
sys.stderr.write("=====STEP2\n")
tr.step(x,y); torch.cuda.synchronize()

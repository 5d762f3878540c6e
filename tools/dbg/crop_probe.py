import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200")]
import torch
from dataset import random_crop_batch, synthetic_tiles
mode = sys.argv[1]
lr, hr = synthetic_tiles(2, 256, seed=5)
o = {"aligned": [[0, 4, 8], [1, 16, 32]], "top_odd": [[0, 5, 8], [1, 17, 32]], "left_odd": [[0, 4, 3], [1, 16, 33]],
     "left_even": [[0, 4, 2], [1, 16, 34]]}[mode]
o = torch.tensor(o, dtype=torch.int32)
(y_nchw, yb), (x_nchw, xb) = random_crop_batch(lr.cuda(), hr.cuda(), 64, o, torch.bfloat16)
torch.cuda.synchronize()
t, top, left = o[1].tolist()
ref = lr[t][:, top:top + 32, left:left + 32]
mn = ref.amin(dim=(1, 2), keepdim=True); mx = ref.amax(dim=(1, 2), keepdim=True)
print(mode, "ok", torch.equal(y_nchw[1].cpu(), (ref - mn) / (mx - mn + 1e-5)))

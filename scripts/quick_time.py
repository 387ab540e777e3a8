import sys, time, torch
sys.path.insert(0, '.')
from oracle import nvp_oracle as O
from tests.helpers import make_model, sampler_like_inputs
import nvp_b200
from nvp_b200 import functional
mode = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
cfg = O.NVPConfig()
torch.manual_seed(0)
m = make_model(cfg, None, mode=mode)
n = 1245184
coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=4)
x = {"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}
gtc = gt.cuda()
for i in range(2):
    ls = m.fwd_loss_bwd(x, gtc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K=3
for i in range(K):
    ls = m.fwd_loss_bwd(x, gtc)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/K
print(f"mode {mode}: {ms:.2f} ms/step -> {n/ms/1e3:.1f} Mpx/s, launches {functional.last_launch_count()}, loss {float(ls)/(3*n*(K+2)):.5f}")

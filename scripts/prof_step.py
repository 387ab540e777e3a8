"""Short profiling driver: a few fused training steps of config S at full batch (for ncu)."""
import sys, torch
sys.path.insert(0, '.')
import bench, nvp_b200
cfg = bench.load_config(sys.argv[2] if len(sys.argv) > 2 else "s")
torch.manual_seed(0)
m = nvp_b200.NVP(out_features=3, encoding_config=cfg, mode="tc").cuda()
from nvp_b200.optim import flatten_parameters
_, flat = flatten_parameters(m)
c, t, g = [x.cuda() for x in bench.synth_batch(bench.N_SAMPLES, 0)]
ls = torch.zeros(1, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    flat.zero_()
    m.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g, loss_sum=ls)
torch.cuda.synchronize()
print("loss", float(ls))

"""Sweep the binned grid path's tile / warp / chunk knobs (env, read per call) at full batch; prints per-kind ms/step."""
import os, sys, torch
sys.path.insert(0, '.')
import bench, nvp_b200
from nvp_b200 import _lib
from nvp_b200.optim import flatten_parameters
cfg = bench.load_config(sys.argv[1] if len(sys.argv) > 1 else "s")
torch.manual_seed(0)
m = nvp_b200.NVP(out_features=3, encoding_config=cfg, mode="tc").cuda()
_, flat = flatten_parameters(m)
c, t, g = [x.cuda() for x in bench.synth_batch(bench.N_SAMPLES, 0)]
ls = torch.zeros(1, device="cuda")
combos = [dict()]
KEYS = ("NVP_GRID_BINNED", "NVP_BIN_TB", "NVP_BIN_CHUNK", "NVP_BIN_WARPS_G", "NVP_BIN_WARPS_S", "NVP_BIN_SPARSE_WARPS_G", "NVP_BIN_SPARSE_WARPS_S")
for combo in combos:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(combo)
    for i in range(3):
        flat.zero_(); m.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g, loss_sum=ls)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    K = 10
    for i in range(K):
        flat.zero_(); m.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g, loss_sum=ls)
    kern = _lib.profile_read(); _lib.profile_enable(False)
    print(combo, " ".join(f"{k}={v[0]/K:.3f}" for k, v in kern.items() if v[1]), flush=True)

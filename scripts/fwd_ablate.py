"""Time the inference forward (no stash) next to the training forward at full batch."""
import sys, torch
sys.path.insert(0, '.')
import bench, nvp_b200
from nvp_b200 import _lib
cfg = bench.load_config(sys.argv[1] if len(sys.argv) > 1 else "s")
torch.manual_seed(0)
m = nvp_b200.NVP(out_features=3, encoding_config=cfg, mode="tc").cuda()
c, t, g = [x.cuda() for x in bench.synth_batch(bench.N_SAMPLES, 0)]
x = {"all_coords": c[None], "temporal_steps": t[None]}
with torch.no_grad():
    for i in range(3): m(x)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for i in range(10): m(x)
    kern = _lib.profile_read(); _lib.profile_enable(False)
print("inference:", " ".join(f"{k}={v[0]/10:.3f}" for k, v in kern.items() if v[1]))

"""Print the in-kernel timeline of the fused forward kernel (library built with `make clean; make TIMELINE=1`).
Runs a few full-batch training steps of config S and reads the clock64 stamps of CTA 0's 4th tile."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import bench, nvp_b200
from nvp_b200 import _lib
cfg = bench.load_config(sys.argv[1] if len(sys.argv) > 1 else "s")
torch.manual_seed(0)
m = nvp_b200.NVP(out_features=3, encoding_config=cfg, mode="tc").cuda()
c, t, g = [x.cuda() for x in bench.synth_batch(bench.N_SAMPLES, 0)]
for i in range(2):
    m.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g)
torch.cuda.synchronize()
buf = (C.c_uint64 * 128)()
_lib.check(_lib.load().nvp_debug_timeline_read(buf, 128), "nvp_debug_timeline_read")
v = list(buf)
t0 = min(x for x in v if x)
names = {0: "acc ready", 1: "math done", 2: "stored", 3: "after barrier"}
for step in range(3):
    for p in range(2):
        print(f"epilogue step {step} panel {p}: " + "  ".join(f"{names[k]} {v[16 * step + 8 * p + k] - t0:7d}" for k in range(4) if v[16 * step + 8 * p + k]))
    print(f"mma      step {step}: " + "  ".join(f"{n} {v[64 + 8 * step + k] - t0:7d}" for k, n in enumerate(("start", "panel0", "panel1", "issued")) if v[64 + 8 * step + k]))

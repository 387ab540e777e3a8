"""Print the in-kernel timeline of the fused forward+backward kernel (library built with `make clean; make TIMELINE=1`):
clock64 stamps of CTA 0 in its 4th tile, relative to the start of that tile's P0."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import bench, nvp_b200
from nvp_b200 import _lib
cfg = bench.load_config("s")
torch.manual_seed(0)
m = nvp_b200.NVP(out_features=3, encoding_config=cfg, mode="tc").cuda()
c, t, g = [x.cuda() for x in bench.synth_batch(bench.N_SAMPLES, 0)]
for i in range(2):
    m.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g)
torch.cuda.synchronize()
buf = (C.c_uint64 * 128)()
_lib.check(_lib.load().nvp_debug_timeline_read(buf, 128), "nvp_debug_timeline_read")
v = list(buf)
t0 = v[0]
names = {0: "P0 acc ready", 16: "P0 start", 1: "P0 end", 2: "P1 acc ready", 17: "P1 start", 3: "P1 end", 4: "P2 start", 5: "P2 end (rgb, drgb)", 6: "P3 end",
         8: "P4 acc ready", 18: "P4 start", 9: "P4 end", 10: "P5 acc ready", 19: "P5 start", 11: "P5 end", 12: "P6 acc ready", 20: "P6 start", 13: "P6 end",
         14: "next tile P0 acc ready",
         48: "reducer0: a2 done", 49: "reducer0: dsp2 done", 50: "reducer0: dsp1 done", 51: "reducer0: dsp0 done",
         64: "mma: F1 z-part issue", 65: "mma: PD_P0[1] seen", 66: "mma: PD_P1[1] seen", 67: "mma: PD_P3[0] seen", 68: "mma: PD_P3[1] seen",
         69: "mma: PD_P4[0]+TF seen", 70: "mma: PD_P4[1] seen", 71: "mma: PD_P5[0] seen", 72: "mma: PD_P5[1] seen",
         80: "mma: AF_P1 committed", 81: "mma: AF_P2 committed", 82: "mma: AF_P4 committed", 83: "mma: AF_P5 committed"}
names[15] = "next tile loop top"
for wp in range(4, 20):
    names[96 + wp] = f"  P5 end of warp {wp}"
    names[32 + wp - 4] = f"  P1 end of warp {wp}"
print("ring wait cycles of the MMA thread in this tile:", v[90]); v[90] = 0
for k, tt in sorted(((k, x) for k, x in enumerate(v) if x), key=lambda kv: kv[1]):
    print(f"{tt - t0:8d}  {names.get(k, 'slot %d' % k)}")

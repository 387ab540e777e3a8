import sys, torch
sys.path.insert(0, '.')
from oracle import nvp_oracle as O
from tests.helpers import make_model, sampler_like_inputs
mode = sys.argv[1] if len(sys.argv) > 1 else 'tc'
cfg = O.NVPConfig()
torch.manual_seed(0)
m = make_model(cfg, None, mode=mode)
n = 1245184
coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=4)
x = {"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}
with torch.no_grad():
    for i in range(3): out = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 5
    for i in range(K): out = m(x)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"forward mode {mode}: {ms:.3f} ms -> {n/ms/1e3:.1f} Mpx/s; out mean {float(out['model_out'].mean()):.5f}")
m2 = make_model(cfg, {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.startswith('wrapper.net.')}, mode='fp32')
with torch.no_grad():
    ref = m2(x)['model_out']
print("max |tc - fp32| over full batch:", float((out['model_out'] - ref).abs().max()))

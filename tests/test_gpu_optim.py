"""GPU test: the fused AdamW step equals torch.optim.AdamW + CosineAnnealingLR (training.py:13-14,73-76)."""
import pytest
import torch

from nvp_b200.optim import FusedAdamW, flatten_parameters

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch_over_a_cosine_schedule():
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 3)).cuda()
    ours = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.Linear(53, 3)).cuda()
    ours.load_state_dict(ref.state_dict())
    T = 20
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.001)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=T, eta_min=1e-5)
    fp, fg = flatten_parameters(ours)
    fopt = FusedAdamW(fp, fg, lr=1e-2, weight_decay=0.001, t_max=T, eta_min=1e-5)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(T):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * (10.0 ** (-(step % 5))) for p in ref.parameters()]
        opt.zero_grad()
        for p, q, gr in zip(ref.parameters(), ours.parameters(), grads):
            p.grad = gr.clone()
            q.grad.add_(gr)                       # accumulate into the flat buffer like the kernels do
        assert abs(fopt.current_lr() - sched.get_last_lr()[0]) < 1e-12
        opt.step()
        sched.step()
        fopt.step(zero_grad=True)
        assert float(fg.abs().max()) == 0.0       # gradient cleared in the same pass
        for p, q in zip(ref.parameters(), ours.parameters()):
            assert float((p - q).abs().max()) <= 2e-6 * float(p.abs().max()) + 1e-9, step
    assert all(q.data_ptr() >= fp.data_ptr() for q in ours.parameters())

"""GPU: the fused step tail ``lasr_clip_adam_step`` (csrc/optim.cu) against the reference's semantics restated with stock torch:
``clip_grad_norm_(params, 5.0)`` -> skip on NaN -> ``Noam.step()`` = ``_step += 1; lr = rate(); torch.optim.Adam.step()``
(/root/reference/liteasr/trainer.py:153-171, optims/noam.py:33-46, optims/adam.py:27-34)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def noam_rate(step, factor, model_dim, warmup):
    return factor * model_dim ** (-0.5) * min(step ** (-0.5), step * warmup ** (-1.5))  # optims/noam.py:40-46


class RefTail:
    """trainer.py:153-171 on a list of parameter tensors (uneven sizes, like a real model)."""

    def __init__(self, flat, sizes, *, noam=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [torch.nn.Parameter(c.clone()) for c in flat.split(sizes)]
        self.opt = torch.optim.Adam(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.noam, self.step_no, self.skipped = noam, 0, 0

    def step(self, gflat, sizes, clip, grad_mult):
        for p, g in zip(self.params, gflat.split(sizes)):
            p.grad = g.clone() * grad_mult  # DDP hands the optimizer the rank-averaged gradient
        norm = torch.nn.utils.clip_grad_norm_(self.params, clip)
        if not math.isnan(float(norm)):
            self.step_no += 1
            if self.noam is not None:
                for pg in self.opt.param_groups:
                    pg["lr"] = noam_rate(self.step_no, *self.noam)
            self.opt.step()
        else:
            self.skipped += 1
        return float(norm)

    def flat(self):
        return torch.cat([p.detach().reshape(-1) for p in self.params])


@pytest.mark.parametrize("variant", ["noam", "adam", "adam_wd"])
def test_clip_adam_step_matches_torch(variant):
    from liteasr_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    sizes = [1000, 64, 4096 * 3, 7, 513, 2048 * 17 + 4, 256]
    n = sum(sizes)
    flat0 = torch.randn(n, generator=g, device=dev) * 0.1
    if variant == "noam":
        kw = dict(beta1=0.9, beta2=0.98, eps=1e-9, weight_decay=0.0, noam_factor=1.0, model_dim=256.0, warmup=10.0, lr=0.0)
        ref = RefTail(flat0, sizes, noam=(1.0, 256, 10), betas=(0.9, 0.98), eps=1e-9)
    else:
        wd = 0.01 if variant == "adam_wd" else 0.0
        kw = dict(beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=wd, noam_factor=0.0, lr=2e-3)
        ref = RefTail(flat0, sizes, lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    p = flat0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    state = torch.zeros(8, device=dev)
    ws = torch.zeros(1024, device=dev)
    # step 1: small gradient (clip inactive); 2: large gradient (clip active); 3: NaN injected (skipped: nothing moves, the
    # step counter does not advance); 4: grad_mult = 1/4 (a 4-rank all-reduce sum); 5-7: ordinary steps
    plan = [(0.01, 1.0, False), (3.0, 1.0, False), (0.5, 1.0, True), (2.0, 0.25, False), (0.1, 1.0, False), (1.0, 0.5, False), (0.02, 1.0, False)]
    updates = 0
    for i, (scale, mult, nan) in enumerate(plan):
        gr = torch.randn(n, generator=g, device=dev) * scale
        if nan:
            gr[12345] = float("nan")
        before = p.clone()
        ops.clip_adam_step(p, gr, m, v, state, ws, grad_mult=mult, max_norm=5.0, **kw)
        norm_ref = ref.step(gr, sizes, 5.0, mult)
        st = state.cpu().tolist()
        if nan:
            assert math.isnan(norm_ref) and st[3] == 1.0
            assert torch.equal(p, before), "a skipped step must not touch the parameters"
        else:
            updates += 1
            assert st[3] == 0.0
            assert math.isclose(st[1], norm_ref, rel_tol=2e-6), (i, st[1], norm_ref)
            clipped = norm_ref > 5.0
            assert clipped == (scale * mult * math.sqrt(n) > 5.0)
            if variant == "noam":
                assert math.isclose(st[2], noam_rate(updates, 1.0, 256, 10), rel_tol=2e-6)
        assert int(st[0]) == updates == ref.step_no
        # fp32 Adam arithmetic in a different operation order: 1e-5 of the accumulated update + 2 ulp of a parameter (|p| < 1)
        err = (p - ref.flat()).abs().max().item()
        upd = (ref.flat() - flat0).abs().max().item()
        assert upd > 1e-4 or nan or i == 0
        assert err <= 1e-5 * upd + 2.4e-7, (variant, i, err, upd)
    assert ref.skipped == 1
    assert torch.allclose(m, torch.cat([ref.opt.state[q]["exp_avg"].reshape(-1) for q in ref.params]), rtol=1e-4, atol=1e-9)
    assert torch.allclose(v, torch.cat([ref.opt.state[q]["exp_avg_sq"].reshape(-1) for q in ref.params]), rtol=1e-4, atol=1e-12)


def test_fused_noam_through_the_store_matches_reference_tail():
    """The same through ``FusedNoam`` on a real ``ParamStore`` (alignment padding between parameters must stay inert)."""
    from liteasr_b200 import functions as F
    from liteasr_b200.models.u2 import U2, U2Config
    from liteasr_b200.optims import FusedNoam, NoamConfig
    torch.manual_seed(0)
    model = U2(U2Config(input_dim=80, vocab_size=50, enc_layers=1, dec_layers=1, enc_dim=64, dec_dim=64, enc_ff_dim=128,
                        dec_ff_dim=128, enc_attn_heads=2, dec_attn_heads=2)).cuda()
    st, _, _ = F.bind(model, torch.device("cuda:0"))
    st.enable_direct_grads()
    opt = FusedNoam(st, NoamConfig(warmup=100))
    ref_params = [torch.nn.Parameter(p.detach().clone()) for _, p in st.named]
    ropt = torch.optim.Adam(ref_params, lr=0.0, betas=(0.9, 0.98), eps=1e-9)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(1, 6):
        st.zero_grads()
        for (_, p), rp in zip(st.named, ref_params):
            gr = torch.randn(p.shape, generator=g, device="cuda") * (0.5 if step != 2 else 5.0)
            p.grad.copy_(gr)
            rp.grad = gr.clone()
        opt.step(5.0)
        torch.nn.utils.clip_grad_norm_(ref_params, 5.0)
        for pg in ropt.param_groups:
            pg["lr"] = noam_rate(step, 1.0, 256, 100)
        ropt.step()
        for (name, p), rp in zip(st.named, ref_params):
            assert torch.allclose(p, rp, rtol=1e-5, atol=1e-7), (step, name)
    assert opt.num_updates() == 5

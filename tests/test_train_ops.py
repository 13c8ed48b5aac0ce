"""Training-tail building blocks: cross-entropy losses + logit gradients, fused LAMB /
Adam steps.  CPU: the LAMB oracle against the reference class.  GPU: the CUDA kernels against torch / the oracle."""
import importlib.util
import os

import pytest
import torch
import torch.nn.functional as F

import refimport
from oracle import optim_oracle

LAMB_PATH = '/root/reference/peract/helpers/optim/lamb.py'


def _tensors(seed, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 10, 1, 1, 1), (64,), (513, 129), (1,), (7, 5), (4096 * 3 + 5,), (3, 3)]
    ps = [torch.randn(s, generator=g) * (0.5 if i != 5 else 3.0) for i, s in enumerate(shapes)]
    ps[6].zero_()                                   # zero weights -> trust ratio 1 (lamb.py:112-113)
    return [p.to(device) for p in ps], g


@pytest.mark.skipif(not os.path.exists(LAMB_PATH), reason='/root/reference not mounted')
def test_lamb_oracle_matches_reference_class():
    spec = importlib.util.spec_from_file_location('ref_lamb', LAMB_PATH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ps, g = _tensors(3)
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    opt = mod.Lamb(ref, lr=5e-4, weight_decay=1e-6, betas=(0.9, 0.999), adam=False)
    mine = [p.clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    for step in range(3):
        grads = [torch.randn(p.shape, generator=g) for p in ps]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        opt.step()
        for p, gr, mm, vv in zip(mine, grads, m, v):
            optim_oracle.lamb_step(p, gr, mm, vv, 5e-4, 0.9, 0.999, 1e-6, 1e-6)
    for r, p in zip(ref, mine):
        assert torch.equal(r.data, p)


@pytest.mark.gpu
def test_cross_entropy_matches_torch(cuda_lib):
    from voxactb_b200 import train
    g = torch.Generator().manual_seed(0)
    for B, N in ((3, 72), (2, 2), (4, 1000003), (1, 4097)):
        x = (torch.randn(B, N, generator=g) * 3).cuda().requires_grad_(True)
        lab = torch.randint(0, N, (B,), generator=g).cuda()
        ref = F.cross_entropy(x, lab, reduction='none')
        (ref.sum() * 0.25).backward()
        loss, grad = train.cross_entropy(x.detach(), lab, grad_scale=0.25)
        assert torch.allclose(loss, ref.detach(), rtol=1e-5, atol=1e-5)
        assert float((grad - x.grad).abs().max()) < 1e-6
    # strided view (a slice of the rot/grip head), no gradient
    q = torch.randn(5, 218, generator=g).cuda()
    lab = torch.randint(0, 72, (5,), generator=g).cuda()
    loss, grad = train.cross_entropy(q[:, 72:144], lab)
    assert grad is None and torch.allclose(loss, F.cross_entropy(q[:, 72:144], lab, reduction='none'), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_peract_losses_match_reference_formulation(cuda_lib):
    """Index-based losses == the reference's one-hot + argmax formulation (agent:517-578)."""
    from voxactb_b200 import train
    g = torch.Generator().manual_seed(1)
    B, V, R = 3, 20, 72
    qt = torch.randn(B, 1, V, V, V, generator=g).cuda()
    qrg = torch.randn(B, 3 * R + 2, generator=g).cuda()
    qc = torch.randn(B, 2, generator=g).cuda()
    at = torch.randint(0, V, (B, 3), generator=g).cuda()
    arg = torch.cat([torch.randint(0, R, (B, 3), generator=g), torch.randint(0, 2, (B, 1), generator=g)], 1).cuda()
    aic = torch.randint(0, 2, (B, 1), generator=g).cuda()
    total, terms, grads = train.peract_losses(qt, qrg, qc, at, arg, aic, R, with_grad=True)
    ce = torch.nn.CrossEntropyLoss(reduction='none')
    qt_r, qrg_r, qc_r = (t.clone().requires_grad_(True) for t in (qt, qrg, qc))
    onehot = torch.zeros(B, 1, V, V, V, device='cuda')
    for b in range(B):
        onehot[b, :, at[b, 0], at[b, 1], at[b, 2]] = 1
    lt = ce(qt_r.view(B, -1), onehot.view(B, -1).argmax(-1))
    lr_ = sum(ce(qrg_r[:, a * R:(a + 1) * R], arg[:, a]) for a in range(3))
    ref_total = (lt + lr_ + ce(qrg_r[:, 3 * R:], arg[:, 3]) + ce(qc_r, aic[:, 0])).mean()
    ref_total.backward()
    assert abs(float(total) - float(ref_total.detach())) < 1e-5 * abs(float(ref_total.detach()))
    assert float((grads['q_trans'] - qt_r.grad).abs().max()) < 1e-7
    assert float((grads['q_rot_grip'] - qrg_r.grad).abs().max()) < 1e-6
    assert float((grads['q_collision'] - qc_r.grad).abs().max()) < 1e-6


@pytest.mark.gpu
def test_fused_lamb_and_adam(cuda_lib):
    from voxactb_b200 import train
    ps, g = _tensors(5)
    cpu = [p.clone() for p in ps]
    m = [torch.zeros_like(p) for p in ps]
    v = [torch.zeros_like(p) for p in ps]
    dev = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    opt = train.Lamb(dev, lr=5e-4, weight_decay=1e-6)
    adam_ref = [torch.nn.Parameter(p.clone()) for p in ps]
    adam_dev = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    o_ref = torch.optim.Adam(adam_ref, lr=1e-3, weight_decay=1e-4)
    o_dev = train.Adam(adam_dev, lr=1e-3, weight_decay=1e-4)
    for step in range(3):
        grads = [torch.randn(p.shape, generator=g) for p in ps]
        for p, d, gr, mm, vv in zip(cpu, dev, grads, m, v):
            optim_oracle.lamb_step(p, gr, mm, vv, 5e-4, 0.9, 0.999, 1e-6, 1e-6)
            d.grad = gr.cuda()
        opt.step()
        for r, d, gr in zip(adam_ref, adam_dev, grads):
            r.grad = gr.clone()
            d.grad = gr.cuda()
        o_ref.step()
        o_dev.step()
    for p, d in zip(cpu, dev):
        assert float((d.detach().cpu() - p).abs().max()) <= 2e-6 * max(1.0, float(p.abs().max()))
    for r, d in zip(adam_ref, adam_dev):
        assert float((d.detach().cpu() - r.detach()).abs().max()) <= 2e-6 * max(1.0, float(r.abs().max()))

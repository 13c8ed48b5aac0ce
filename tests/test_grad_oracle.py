"""Row a18 groundwork: every closed-form adjoint in oracle/grad_oracle.py (the operation-level oracle of the backward
kernels) against torch autograd over the forward restatement, CPU only."""
import pytest
import torch
import torch.nn.functional as F

from oracle import grad_oracle as G
from oracle import qnet_oracle as Q


def close(a, b, tol=2e-5):
    a, b = a.detach(), b.detach()
    scale = max(float(b.abs().max()), 1e-12)
    assert float((a - b).abs().max()) <= tol * scale, (float((a - b).abs().max()), scale)


def rnd(*shape, seed=0, grad=True):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64).requires_grad_(grad)


def test_lrelu_and_geglu():
    x, gy = rnd(5, 7, seed=1), rnd(5, 7, seed=2, grad=False)
    y = F.leaky_relu(x, 0.02)
    y.backward(gy)
    close(G.lrelu_backward(gy, y.detach()), x.grad)
    a, g = rnd(4, 6, seed=3), rnd(4, 6, seed=4)
    (a * F.gelu(g)).backward(gy[:4, :6])
    ga, gg = G.geglu_backward(gy[:4, :6], a.detach(), g.detach())
    close(ga, a.grad)
    close(gg, g.grad)


def test_linear_and_layernorm():
    x, w, b, gy = rnd(3, 5, 8, seed=5), rnd(6, 8, seed=6), rnd(6, seed=7), rnd(3, 5, 6, seed=8, grad=False)
    F.linear(x, w, b).backward(gy)
    gx, gw, gb = G.linear_backward(gy, x.detach(), w.detach())
    close(gx, x.grad); close(gw, w.grad); close(gb, b.grad)
    x, w, b, gy = rnd(2, 4, 16, seed=9), rnd(16, seed=10), rnd(16, seed=11), rnd(2, 4, 16, seed=12, grad=False)
    F.layer_norm(x, (16,), w, b).backward(gy)
    gx, gw, gb = G.layernorm_backward(gy, x.detach(), w.detach())
    close(gx, x.grad); close(gw, w.grad); close(gb, b.grad)


def test_attention_core():
    q, k, v, go = rnd(3, 9, 8, seed=13), rnd(3, 11, 8, seed=14), rnd(3, 11, 8, seed=15), rnd(3, 9, 8, seed=16, grad=False)
    (torch.softmax(q @ k.transpose(-1, -2) * 8 ** -0.5, -1) @ v).backward(go)
    gq, gk, gv = G.attention_core_backward(go, q.detach(), k.detach(), v.detach(), 8 ** -0.5)
    close(gq, q.grad); close(gk, k.grad); close(gv, v.grad)


@pytest.mark.parametrize('k,stride,n,act', [(3, 1, 6, 'lrelu'), (5, 1, 5, 'lrelu'), (5, 5, 10, 'lrelu'), (5, 4, 8, 'lrelu'),
                                            (1, 1, 4, 'lrelu'), (3, 1, 5, None)])
def test_conv3d_block(k, stride, n, act):
    x, w, b = rnd(2, 3, n, n, n, seed=17), rnd(4, 3, k, k, k, seed=18), rnd(4, seed=19)
    y = Q.conv3d_block(x, w, b, stride, act)
    gy = rnd(*y.shape, seed=20, grad=False)
    y.backward(gy)
    gx, gw, gb = G.conv3d_block_backward(gy, x.detach(), w.detach(), y.detach(), stride, act)
    close(gx, x.grad); close(gw, w.grad); close(gb, b.grad)


def test_upsample_trilinear():
    x = rnd(2, 3, 4, 3, 5, seed=21)
    y = F.interpolate(x, scale_factor=5, mode='trilinear', align_corners=False)
    close(torch.einsum('bcdhw,zd,yh,xw->bczyx', x.detach().float(), G.upsample_matrix(4, 5), G.upsample_matrix(3, 5),
                       G.upsample_matrix(5, 5)).double(), y.detach(), 1e-6)
    gy = rnd(*y.shape, seed=22, grad=False)
    y.backward(gy)
    close(G.upsample_trilinear_backward(gy.float(), 5).double(), x.grad, 1e-5)


def test_spatial_softmax_and_maxpool_and_ce():
    x = (rnd(2, 3, 4, 5, 6, seed=23, grad=False).float() * 0.02).requires_grad_(True)   # T = 0.01: keep the softmax soft
    e = Q.spatial_softmax3d(x)
    ge = rnd(*e.shape, seed=24, grad=False).float()
    e.backward(ge)
    close(G.spatial_softmax3d_backward(ge, x.detach()), x.grad, 2e-4)
    x2 = rnd(2, 3, 4, 4, 4, seed=25)
    gm = rnd(2, 3, seed=26, grad=False)
    x2.amax(dim=(2, 3, 4)).backward(gm)
    close(G.global_maxpool_backward(gm, x2.detach()), x2.grad)
    lg, idx = rnd(4, 9, seed=27), torch.tensor([0, 3, 8, 3])
    (F.cross_entropy(lg, idx, reduction='none') * 0.25).sum().backward()
    close(G.cross_entropy_backward(lg.detach(), idx, 0.25), lg.grad)


@pytest.mark.parametrize('scale,k,n', [(5, 5, 4), (4, 5, 3), (2, 3, 5), (5, 3, 3)])
def test_polyphase_fold_identity_and_transpose(scale, k, n):
    """K6b: Upsample(x s, trilinear, align_corners=False) followed by the replicate-padded k^3 convolution equals the
    s^3 folded 3x3x3 phase convolutions with clamped coarse neighbours -- borders included; the fold's transpose is the
    weight gradient."""
    from oracle import fold_oracle as FO
    x, w, b = rnd(2, 3, n, n, n, seed=31, grad=False), rnd(4, 3, k, k, k, seed=32), rnd(4, seed=33, grad=False)
    up = F.interpolate(x, scale_factor=scale, mode='trilinear', align_corners=False)
    ref = F.conv3d(F.pad(up, [k // 2] * 6, mode='replicate'), w, b)
    wf = FO.fold_upconv_weights(w, scale)
    close(FO.folded_upconv(x, wf, b, scale), ref.detach(), 1e-10)
    gy = rnd(*ref.shape, seed=34, grad=False)
    ref.backward(gy)
    # weight gradient through the folded form: phase-kernel gradients, then the fold transpose
    wf2 = wf.detach().clone().requires_grad_(True)
    FO.folded_upconv(x, wf2, b, scale).backward(gy)
    close(FO.fold_upconv_weights_backward(wf2.grad, scale, k), w.grad, 1e-9)

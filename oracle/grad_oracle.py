"""TEST INFRASTRUCTURE ONLY -- explicit adjoints of the Q-network's building blocks.

The reference obtains its gradients from torch autograd (`total_loss.backward()`, qattention_peract_bc_agent.py:581) over
the modules of perceiver_lang_io.py / helpers/network_utils.py.  This file writes every adjoint out as the closed-form
expression a CUDA backward kernel has to implement (no autograd inside), so that the kernels of the training milestone
have an operation-level oracle in the same way the forward kernels have oracle/qnet_oracle.py.  Each function cites the
forward it differentiates; tests/test_grad_oracle.py checks all of them against torch autograd on the forward
restatement (CPU), and oracle/train_oracle.py + tests/golden/train_v20*.npz pin the end-to-end gradients to the
reference.  Nothing in voxactb_b200/ imports this module.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ elementwise
def lrelu_backward(gy, y, slope=0.02):
    """LeakyReLU (network_utils.py:12-27 'lrelu', slope 0.02): y > 0 <=> pre-activation > 0, so the OUTPUT suffices."""
    return torch.where(y > 0, gy, gy * slope)


def geglu_backward(gy, a, g):
    """GEGLU (perceiver_lang_io.py:74-77): y = a * gelu_erf(g).  Returns (ga, gg)."""
    cdf = 0.5 * (1.0 + torch.erf(g * (1.0 / math.sqrt(2.0))))
    pdf = torch.exp(-0.5 * g * g) * (1.0 / math.sqrt(2.0 * math.pi))
    return gy * g * cdf, gy * a * (cdf + g * pdf)


# ------------------------------------------------------------------------------------------------ linear / norm / attention
def linear_backward(gy, x, w):
    """nn.Linear: y = x w^T + b.  Returns (gx, gw, gb); leading dimensions are flattened for the weight gradient."""
    gx = gy @ w
    g2, x2 = gy.reshape(-1, gy.shape[-1]), x.reshape(-1, x.shape[-1])
    return gx, g2.t() @ x2, g2.sum(0)


def layernorm_backward(gy, x, weight, eps=1e-5):
    """nn.LayerNorm over the last dimension (PreNorm, perceiver_lang_io.py:56-71), biased variance.
    Returns (gx, gweight, gbias)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    xhat = (x - mean) * rstd
    gxhat = gy * weight
    gx = (gxhat - gxhat.mean(-1, keepdim=True) - xhat * (gxhat * xhat).mean(-1, keepdim=True)) * rstd
    red = tuple(range(x.dim() - 1))
    return gx, (gy * xhat).sum(red), gy.sum(red)


def attention_core_backward(go, q, k, v, scale):
    """softmax(q k^T * scale) v for [B*, Nq, d] / [B*, Nk, d] (Attention.forward, perceiver_lang_io.py:111-128, no mask,
    no dropout).  Returns (gq, gk, gv).  The probabilities are recomputed (flash-style backward: only the row
    log-sum-exp needs to be kept by a kernel)."""
    s = (q @ k.transpose(-1, -2)) * scale
    p = torch.softmax(s, dim=-1)
    gv = p.transpose(-1, -2) @ go
    gp = go @ v.transpose(-1, -2)
    delta = (gp * p).sum(-1, keepdim=True)                     # = rowsum(go * o)
    gs = p * (gp - delta) * scale
    return gs @ k, gs.transpose(-1, -2) @ q, gv


# ------------------------------------------------------------------------------------------------ convolutions
def replicate_pad_backward(gxp, pad):
    """Adjoint of F.pad(x, [pad]*6, mode='replicate') (Conv3DBlock, network_utils.py:147-150): every halo element adds
    into the border element it replicates; separable per axis."""
    g = gxp
    for dim in (2, 3, 4):
        n = g.shape[dim] - 2 * pad
        body = g.narrow(dim, pad, n).clone()
        body.narrow(dim, 0, 1).add_(g.narrow(dim, 0, pad).sum(dim, keepdim=True))
        body.narrow(dim, n - 1, 1).add_(g.narrow(dim, pad + n, pad).sum(dim, keepdim=True))
        g = body
    return g


def conv3d_block_backward(gy, x, w, y, stride, activation='lrelu'):
    """Conv3DBlock (network_utils.py:128-170): y = act(conv3d(replicate_pad(x, k//2), w, b, stride)).
    Returns (gx, gw, gb).  dgrad = transposed convolution of the pre-activation gradient followed by the padding
    adjoint; wgrad = correlation of the padded input with the pre-activation gradient (an implicit GEMM with
    K = B * output positions)."""
    k = w.shape[-1]
    pad = k // 2
    gz = lrelu_backward(gy, y) if activation == 'lrelu' else gy
    xp = F.pad(x, [pad] * 6, mode='replicate')
    gw = torch.zeros_like(w)
    od, oh, ow = gz.shape[2:]
    for a in range(k):
        for b_ in range(k):
            for c in range(k):
                win = xp[:, :, a:a + stride * (od - 1) + 1:stride, b_:b_ + stride * (oh - 1) + 1:stride,
                         c:c + stride * (ow - 1) + 1:stride]
                gw[:, :, a, b_, c] = torch.einsum('bozyx,bizyx->oi', gz, win)
    gb = gz.sum((0, 2, 3, 4))
    # dgrad on the padded grid: gxp[b, ci, o*stride + tap] += sum_co gz[b, co, o] * w[co, ci, tap]
    gxp = torch.zeros_like(xp)
    for a in range(k):
        for b_ in range(k):
            for c in range(k):
                gxp[:, :, a:a + stride * (od - 1) + 1:stride, b_:b_ + stride * (oh - 1) + 1:stride,
                    c:c + stride * (ow - 1) + 1:stride] += torch.einsum('bozyx,oi->bizyx', gz, w[:, :, a, b_, c])
    return replicate_pad_backward(gxp, pad), gw, gb


def upsample_matrix(n, scale):
    """1-D matrix U [n*scale, n] of nn.Upsample(scale_factor, mode='trilinear', align_corners=False)
    (Conv3DUpsampleBlock, network_utils.py:245-247): src = (dst + 0.5) / scale - 0.5 clamped at 0, two-tap lerp."""
    U = torch.zeros(n * scale, n)
    for o in range(n * scale):
        src = max((o + 0.5) / scale - 0.5, 0.0)
        i0 = min(int(math.floor(src)), n - 1)
        i1 = min(i0 + 1, n - 1)
        f = src - i0
        U[o, i0] += 1.0 - f
        U[o, i1] += f
    return U


def upsample_trilinear_backward(gy, scale):
    """Adjoint of the trilinear x`scale` upsampling: the transposed 1-D matrix along each axis."""
    d, h, w = (s // scale for s in gy.shape[2:])
    Ud, Uh, Uw = upsample_matrix(d, scale), upsample_matrix(h, scale), upsample_matrix(w, scale)
    return torch.einsum('bczyx,zd,yh,xw->bcdhw', gy, Ud, Uh, Uw)


# ------------------------------------------------------------------------------------------------ pooling heads
def spatial_softmax_positions(d, h, w):
    """The coordinate buffers of SpatialSoftmax3D (network_utils.py:782-792), meshgrid('xy') quirk included:
    pos_x runs along tensor axis H, pos_y along D, pos_z along W.  Returns [3, d*h*w] (x, y, z)."""
    import numpy as np
    px, py, pz = np.meshgrid(np.linspace(-1., 1., d), np.linspace(-1., 1., h), np.linspace(-1., 1., w))
    return torch.stack([torch.from_numpy(a.reshape(-1)).float() for a in (px, py, pz)])


def spatial_softmax3d_backward(ge, x, temperature=0.01):
    """SpatialSoftmax3D (network_utils.py:773-809): e[b, c, :] = sum_v softmax_v(x[b,c,v] / T) * pos[v, :], returned as
    [B, 3C] with the channel-major (x, y, z) interleave of qnet_oracle.spatial_softmax3d.  ge: [B, 3C].
    gx[b,c,v] = p[v] / T * sum_k ge[b,c,k] (pos[v,k] - e[b,c,k])."""
    B, C, D, H, W = x.shape
    pos = spatial_softmax_positions(D, H, W)                  # [3, D*H*W]: pos_x, pos_y, pos_z rows
    p = torch.softmax(x.reshape(B, C, -1) / temperature, dim=-1)
    e = torch.einsum('bcv,kv->bck', p, pos)
    g = ge.reshape(B, C, 3)
    inner = torch.einsum('bck,kv->bcv', g, pos) - (g * e).sum(-1, keepdim=True)
    return (p * inner / temperature).reshape(x.shape)


def global_maxpool_backward(gm, x):
    """AdaptiveMaxPool3d(1) (perceiver_lang_io.py:242): the gradient goes to the (first) arg-max voxel of each (b, c)."""
    B, C = x.shape[:2]
    flat = x.reshape(B, C, -1)
    idx = flat.argmax(-1, keepdim=True)
    return torch.zeros_like(flat).scatter_(2, idx, gm.reshape(B, C, 1)).reshape(x.shape)


def cross_entropy_backward(logits, idx, scale):
    """CrossEntropyLoss(reduction='none') on label indices (agent:391-392) times `scale` (loss weight / batch size)."""
    g = torch.softmax(logits, dim=-1)
    g.scatter_add_(1, idx.long().reshape(-1, 1), -torch.ones(logits.shape[0], 1, dtype=logits.dtype))
    return g * scale

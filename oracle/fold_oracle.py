"""TEST INFRASTRUCTURE ONLY -- the polyphase identity behind the folded up-convolution.

Conv3DUpsampleBlock (helpers/network_utils.py:237-254) applies `nn.Upsample(scale_factor=s, mode='trilinear',
align_corners=False)` and then a k x k x k replicate-padded convolution with no non-linearity in between.  Both are
linear and separable, so for every output phase r in {0..s-1}^3 the pair is ONE 3x3x3 convolution on the coarse grid
with clamped (replicate) neighbours:

    z[s*i + r] = sum_{delta in {-1,0,1}^3} Wfold[r][delta] x[clamp(i + delta)]

`fold_matrix` gives the 1-D coefficients F[r][delta][a] (how much of fine tap a of phase r lands on coarse offset
delta); `fold_upconv_weights` contracts them with the k^3 weights -- this is what `fold_upconv_weights_kernel`
(csrc/qnet.cu) computes on the device; `fold_upconv_weights_backward` is its transpose (the weight gradient of the
training step).  tests/test_grad_oracle.py checks the identity against
F.interpolate + F.conv3d and the transpose against autograd.  Nothing in voxactb_b200/ imports this module.
"""
import math

import torch
import torch.nn.functional as F


def fold_matrix(scale, k):
    """F [scale, 3, k]: z[s*i + r] = sum_a w[a] * y[s*i + r + a - k//2] with y the (virtually unclamped) linear
    interpolation of x, expressed on the coarse neighbours i-1, i, i+1.  Requires k//2 <= scale (every fine tap of a
    phase then falls between coarse offsets -1 and +1)."""
    p = k // 2
    assert p <= scale, 'fine taps would reach beyond the three coarse neighbours'
    Fm = torch.zeros(scale, 3, k, dtype=torch.float64)
    for r in range(scale):
        for a in range(k):
            rel = (r + a - p + 0.5) / scale - 0.5          # source coordinate relative to coarse index i
            j0 = math.floor(rel)
            frac = rel - j0
            for d, wgt in ((j0, 1.0 - frac), (j0 + 1, frac)):
                if wgt != 0.0:
                    assert -1 <= d <= 1
                    Fm[r, d + 1, a] += wgt
    return Fm


def fold_upconv_weights(w, scale):
    """w [Co, Ci, k, k, k] -> Wfold [s, s, s, Co, Ci, 3, 3, 3] (phase-major)."""
    Fm = fold_matrix(scale, w.shape[-1]).to(w.dtype)
    return torch.einsum('oiabc,pda,qeb,rfc->pqroidef', w, Fm, Fm, Fm)


def fold_upconv_weights_backward(gwf, scale, k):
    """Transpose of fold_upconv_weights: gradient w.r.t. the k^3 weights from the gradient of the phase kernels."""
    Fm = fold_matrix(scale, k).to(gwf.dtype)
    return torch.einsum('pqroidef,pda,qeb,rfc->oiabc', gwf, Fm, Fm, Fm)


def folded_upconv(x, wf, bias, scale):
    """The folded form evaluated directly: per phase a 3x3x3 replicate-padded convolution on the coarse grid, the
    phases interleaved into the fine grid.  x [B, Ci, n, n, n] -> [B, Co, s*n, s*n, s*n] (no activation)."""
    B, Ci, D, H, W = x.shape
    Co = wf.shape[3]
    xp = F.pad(x, [1] * 6, mode='replicate')
    out = x.new_zeros(B, Co, D * scale, H * scale, W * scale)
    for p in range(scale):
        for q in range(scale):
            for r in range(scale):
                out[:, :, p::scale, q::scale, r::scale] = F.conv3d(xp, wf[p, q, r], bias)
    return out

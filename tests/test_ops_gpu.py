"""Per-kernel parity through the C ABI against plain PyTorch fp32 references of the same op (CPU)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

import util
from oracle import qnet_oracle
from voxactb_b200 import _lib

pytestmark = pytest.mark.gpu

MODES = [_lib.MATH_FP32_SIMT, _lib.MATH_F16X3]
TOL = {_lib.MATH_FP32_SIMT: 2e-5, _lib.MATH_F16X3: 1e-4}


def ws(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device='cuda')


class tensor_core_check:
    """In BF16X3 mode the op must have launched at least one tcgen05 GEMM (no silent FFMA fallback)."""

    def __init__(self, lib, mode, expect=True):
        self.lib, self.mode, self.expect = lib, mode, expect

    def __enter__(self):
        self.before = self.lib.vxb_umma_launch_count()

    def __exit__(self, *exc):
        if exc[0] is None and self.mode == _lib.MATH_F16X3 and self.expect:
            assert self.lib.vxb_umma_launch_count() > self.before, 'BF16X3 mode did not use tcgen05'


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('M,N,K', [(300, 128, 512), (77, 128, 512), (2048, 4096, 512), (16, 220, 64),
                                   (5, 64, 7), (1000, 64, 64), (129, 65, 33), (4096, 512, 2048),
                                   (1000, 1024, 520), (8077, 128, 128)])
def test_linear(cuda_lib, mode, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    R = torch.randn(M, N, generator=g)
    ref = F.leaky_relu(F.linear(A, W, b) * 1.0, 0.02) + R
    Ac, Wc, bc, Rc = A.cuda(), W.cuda(), b.cuda(), R.cuda()
    C = torch.empty(M, N, device='cuda')
    wk = ws(cuda_lib.vxb_linear_workspace_bytes(M, N, K))
    with tensor_core_check(cuda_lib, mode, expect=M >= 128 and N >= 32 and K >= 32):
        _lib.check(cuda_lib.vxb_linear_f32(_lib.ptr(Ac), K, _lib.ptr(Wc), K, _lib.ptr(bc), _lib.ptr(Rc), M,
                                           _lib.ptr(C), N, M, N, K, 1.0, 0.02, mode, _lib.ptr(wk), wk.numel(),
                                           _lib.stream()), 'linear')
    assert util.rel_err(C, ref) < TOL[mode]
    # broadcast residual rows (latents) + alpha + no activation
    Rb = torch.randn(3, N, generator=g)
    ref2 = 0.5 * F.linear(A, W) + Rb[torch.arange(M) % 3]
    Rbc = Rb.cuda()
    _lib.check(cuda_lib.vxb_linear_f32(_lib.ptr(Ac), K, _lib.ptr(Wc), K, None, _lib.ptr(Rbc), 3,
                                       _lib.ptr(C), N, M, N, K, 0.5, -1.0, mode, _lib.ptr(wk), wk.numel(),
                                       _lib.stream()), 'linear')
    assert util.rel_err(C, ref2) < TOL[mode]


def test_layernorm(cuda_lib):
    g = torch.Generator().manual_seed(0)
    for rows, n in ((100, 128), (2049, 512), (7, 192)):
        x = torch.randn(rows, n, generator=g) * 3 + 1
        w, b = torch.randn(n, generator=g), torch.randn(n, generator=g)
        y = torch.empty(rows, n, device='cuda')
        xc, wc, bc = x.cuda(), w.cuda(), b.cuda()   # keep the device copies alive across the call
        _lib.check(cuda_lib.vxb_layernorm_f32(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc),
                                              _lib.ptr(y), rows, n, _lib.stream()), 'layernorm')
        assert util.rel_err(y, F.layer_norm(x, (n,), w, b)) < 1e-5


@pytest.mark.parametrize('S,C', [(8, 64), (20, 128), (13, 192), (32, 64)])
def test_spatial_softmax_and_max(cuda_lib, S, C):
    g = torch.Generator().manual_seed(S)
    x = torch.randn(2, C, S, S, S, generator=g) * 0.2       # /0.01 -> logits ~ N(0, 20): peaky
    ref = qnet_oracle.spatial_softmax3d(x)
    refmax = x.amax(dim=(2, 3, 4))
    xc = x.permute(0, 2, 3, 4, 1).contiguous().cuda()
    ss = torch.empty(2, 3 * C, device='cuda')
    mx = torch.empty(2, C, device='cuda')
    w = ws(cuda_lib.vxb_spatial_softmax_workspace_bytes(2, S ** 3, C))
    _lib.check(cuda_lib.vxb_spatial_softmax_f32(_lib.ptr(xc), 2, S, S, S, C, _lib.ptr(ss), 3 * C, _lib.ptr(mx), C,
                                                _lib.ptr(w), w.numel(), _lib.stream()), 'ss')
    assert torch.equal(mx.cpu(), refmax)
    assert float((ss.cpu() - ref).abs().max()) < 2e-5


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('Di,Ci,Co,k,s', [(20, 64, 64, 5, 5), (16, 64, 64, 5, 4), (8, 128, 64, 5, 1),
                                          (12, 64, 64, 3, 1), (10, 16, 32, 3, 1), (20, 64, 64, 3, 1),
                                          (37, 64, 64, 3, 1)])
def test_conv3d(cuda_lib, mode, Di, Ci, Co, k, s):
    g = torch.Generator().manual_seed(Di + k)
    x = torch.randn(2, Ci, Di, Di, Di, generator=g)
    w = torch.randn(Co, Ci, k, k, k, generator=g) / (Ci * k ** 3) ** 0.5
    b = torch.randn(Co, generator=g)
    ref = qnet_oracle.conv3d_block(x, w, b, s, 'lrelu').permute(0, 2, 3, 4, 1)
    y = torch.empty(ref.shape, device='cuda')
    wk = ws(cuda_lib.vxb_conv3d_workspace_bytes(2, Di, Ci, Co, k))
    xc, wc, bc = x.permute(0, 2, 3, 4, 1).contiguous().cuda(), w.cuda(), b.cuda()
    with tensor_core_check(cuda_lib, mode, expect=Ci % 64 == 0 and Co == 64):
        _lib.check(cuda_lib.vxb_conv3d_f32(_lib.ptr(xc), _lib.ptr(wc),
                                           _lib.ptr(bc), _lib.ptr(y), 2, Di, Ci, Co, k, s, 0.02, mode,
                                           _lib.ptr(wk), wk.numel(), _lib.stream()), 'conv3d')
    assert util.rel_err(y, ref) < TOL[mode]


@pytest.mark.parametrize('Di,scale', [(12, 1.0), (20, 1.0), (37, 1.0), (20, 300.0), (20, 1e-3)])
def test_conv3d_f16_fp8_corrected(cuda_lib, Di, scale):
    """VXB_MATH_F16F8C: fp16 hi*hi + one E4M3 MMA for both 2^-11 correction terms (conv_f8c.cuh) holds the same tolerance as
    the three-term split, also for activations far from O(1) (the e4m3 scales are derived from the data on the device)."""
    g = torch.Generator().manual_seed(Di)
    x = torch.randn(2, 64, Di, Di, Di, generator=g) * scale
    x[0, :, 0, 0, 0] *= 7.0                               # a few outliers: the scale follows the maximum
    w = torch.randn(64, 64, 3, 3, 3, generator=g) / (64 * 27) ** 0.5
    b = torch.randn(64, generator=g) * scale
    ref = qnet_oracle.conv3d_block(x, w, b, 1, 'lrelu').permute(0, 2, 3, 4, 1)
    y = torch.empty(ref.shape, device='cuda')
    wk = ws(cuda_lib.vxb_conv3d_workspace_bytes(2, Di, 64, 64, 3))
    xc, wc, bc = x.permute(0, 2, 3, 4, 1).contiguous().cuda(), w.cuda(), b.cuda()
    before = cuda_lib.vxb_umma_launch_count()
    _lib.check(cuda_lib.vxb_conv3d_f32(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), 2, Di, 64, 64, 3, 1, 0.02,
                                       _lib.MATH_F16F8C, _lib.ptr(wk), wk.numel(), _lib.stream()), 'conv3d f8c')
    assert cuda_lib.vxb_umma_launch_count() > before
    assert util.rel_err(y, ref) < 1e-4


@pytest.mark.parametrize('V', [20, 32, 50, 37])
def test_trans_decoder_stencil(cuda_lib, V):
    """Conv3d(64 -> 1, k3, replicate pad, no activation) = trans_decoder (perceiver_lang_io.py:308-311)."""
    g = torch.Generator().manual_seed(V)
    x = torch.randn(2, 64, V, V, V, generator=g)
    w = torch.randn(1, 64, 3, 3, 3, generator=g) / (64 * 27) ** 0.5
    b = torch.randn(1, generator=g)
    ref = qnet_oracle.conv3d_block(x, w, b, 1, None).permute(0, 2, 3, 4, 1)
    y = torch.empty(ref.shape, device='cuda')
    wk = ws(cuda_lib.vxb_conv3d_workspace_bytes(2, V, 64, 1, 3))
    xc, wc, bc = x.permute(0, 2, 3, 4, 1).contiguous().cuda(), w.cuda(), b.cuda()
    _lib.check(cuda_lib.vxb_conv3d_f32(_lib.ptr(xc), _lib.ptr(wc), _lib.ptr(bc), _lib.ptr(y), 2, V, 64, 1, 3, 1,
                                       -1.0, _lib.MATH_FP32_SIMT, _lib.ptr(wk), wk.numel(), _lib.stream()), 'conv3d')
    assert util.rel_err(y, ref) < 1e-5


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('S,k,s', [(4, 5, 5), (5, 5, 4), (3, 9, 8), (6, 3, 2)])
def test_upconv3d_equals_upsample_then_conv(cuda_lib, mode, S, k, s):
    """conv_k(replicate) o trilinear-upsample folded into s^3 polyphase 3^3 convs == the reference
    composition (network_utils.py:245-251)."""
    g = torch.Generator().manual_seed(S * k)
    Ci = Co = 64
    x = torch.randn(2, Ci, S, S, S, generator=g)
    w = torch.randn(Co, Ci, k, k, k, generator=g) / (Ci * k ** 3) ** 0.5
    b = torch.randn(Co, generator=g)
    up = F.interpolate(x, scale_factor=s, mode='trilinear', align_corners=False)
    ref = qnet_oracle.conv3d_block(up, w, b, 1, 'lrelu').permute(0, 2, 3, 4, 1)
    y = torch.empty(ref.shape, device='cuda')
    wk = ws(cuda_lib.vxb_upconv3d_workspace_bytes(2, S, Ci, Co, k, s))
    xc, wc, bc = x.permute(0, 2, 3, 4, 1).contiguous().cuda(), w.cuda(), b.cuda()
    with tensor_core_check(cuda_lib, mode):
        _lib.check(cuda_lib.vxb_upconv3d_f32(_lib.ptr(xc), _lib.ptr(wc),
                                             _lib.ptr(bc), _lib.ptr(y), 2, S, Ci, Co, k, s, 0.02, mode,
                                             _lib.ptr(wk), wk.numel(), _lib.stream()), 'upconv3d')
    assert util.rel_err(y, ref) < max(TOL[mode], 3e-5)


def test_upconv3d_rejects_unfoldable_geometry(cuda_lib):
    x = torch.zeros(1, 2, 2, 2, 64, device='cuda')
    w = torch.zeros(64, 64, 7, 7, 7, device='cuda')
    y = torch.zeros(1, 4, 4, 4, 64, device='cuda')
    wk = ws(1 << 20)
    rc = cuda_lib.vxb_upconv3d_f32(_lib.ptr(x), _lib.ptr(w), None, _lib.ptr(y), 1, 2, 64, 64, 7, 2, 0.02, 0,
                                   _lib.ptr(wk), wk.numel(), _lib.stream())
    assert rc == -2 and b'folding' in cuda_lib.vxb_last_error()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('B,H,Nq,Nk,dh', [(2, 1, 200, 589, 64), (2, 8, 256, 256, 64), (1, 1, 589, 130, 64)])
def test_attention(cuda_lib, mode, B, H, Nq, Nk, dh):
    g = torch.Generator().manual_seed(Nq)
    q = torch.randn(B, Nq, H * dh, generator=g)
    kv = torch.randn(B, Nk, 2 * H * dh, generator=g)
    k, v = kv[..., :H * dh], kv[..., H * dh:]
    qh = q.view(B, Nq, H, dh).transpose(1, 2)
    kh = k.reshape(B, Nk, H, dh).transpose(1, 2)
    vh = v.reshape(B, Nk, H, dh).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) * dh ** -0.5, -1) @ vh).transpose(1, 2).reshape(B, Nq, H * dh)
    qc, kvc = q.cuda(), kv.cuda()
    out = torch.empty(B, Nq, H * dh, device='cuda')
    wk = ws(cuda_lib.vxb_attention_workspace_bytes(B, H, Nq, Nk))
    kptr = ctypes.c_void_p(kvc.data_ptr())
    vptr = ctypes.c_void_p(kvc.data_ptr() + H * dh * 4)
    _lib.check(cuda_lib.vxb_attention_f32(_lib.ptr(qc), H * dh, Nq * H * dh, kptr, vptr, 2 * H * dh,
                                          Nk * 2 * H * dh, _lib.ptr(out), H * dh, Nq * H * dh, B, H, Nq, Nk, dh,
                                          dh ** -0.5, mode, _lib.ptr(wk), wk.numel(), _lib.stream()), 'attention')
    assert util.rel_err(out, ref) < TOL[mode]

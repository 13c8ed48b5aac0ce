// Bandwidth-bound kernels of the Q-network (everything that is not a dense contraction).
#pragma once
#include "common.cuh"

namespace vxb {

// ---------------------------------------------------------------- LayerNorm (warp per row)
// nn.LayerNorm semantics (eps 1e-5, biased variance); reference PreNorm, perceiver_lang_io.py:56-71
static __global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w,
                 const float* __restrict__ b, float* __restrict__ y, int rows, int n,
                 int rows_per_batch, size_t x_batch_stride) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // input rows may be a strided slice per batch (decoder queries skip the language rows)
  const float* xr = x + (size_t)(row / rows_per_batch) * x_batch_stride + (size_t)(row % rows_per_batch) * n;
  float s = 0.f;
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    s += v.x + v.y + v.z + v.w;
  }
  s = warp_sum(s);
  const float mean = s / (float)n;
  float q = 0.f;
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    float a = v.x - mean, bb = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += a * a + bb * bb + c * c + d * d;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)n + 1e-5f);
  float* yr = y + (size_t)row * n;
  for (int i = lane * 4; i < n; i += 128) {
    float4 v = *reinterpret_cast<const float4*>(xr + i);
    float4 g = *reinterpret_cast<const float4*>(w + i);
    float4 be = *reinterpret_cast<const float4*>(b + i);
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + be.x;
    o.y = (v.y - mean) * rstd * g.y + be.y;
    o.z = (v.z - mean) * rstd * g.z + be.z;
    o.w = (v.w - mean) * rstd * g.w + be.w;
    *reinterpret_cast<float4*>(yr + i) = o;
  }
}

// ---------------------------------------------------------------- GEGLU: out = h[:, :n] * gelu_erf(h[:, n:])
// reference GEGLU, perceiver_lang_io.py:74-77 (F.gelu default = exact erf form)
static __global__ void __launch_bounds__(256)
geglu_kernel(const float* __restrict__ h, float* __restrict__ out, size_t rows, int n) {
  const size_t total4 = rows * (size_t)(n / 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (n / 4);
    const int c = (int)(i % (n / 4)) * 4;
    const float4 a = *reinterpret_cast<const float4*>(h + r * 2 * n + c);
    const float4 g = *reinterpret_cast<const float4*>(h + r * 2 * n + n + c);
    float4 o;
    o.x = a.x * (0.5f * g.x * (1.f + erff(g.x * 0.70710678118654752f)));
    o.y = a.y * (0.5f * g.y * (1.f + erff(g.y * 0.70710678118654752f)));
    o.z = a.z * (0.5f * g.z * (1.f + erff(g.z * 0.70710678118654752f)));
    o.w = a.w * (0.5f * g.w * (1.f + erff(g.w * 0.70710678118654752f)));
    *reinterpret_cast<float4*>(out + r * n + c) = o;
  }
}

// ---------------------------------------------------------------- row softmax in place (block per row)
// rows of `n` valid columns with leading dimension ld; pad columns [n, ld) are zeroed.
static __global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ x, int n, int ld) {
  float* xr = x + (size_t)blockIdx.x * ld;
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) m = fmaxf(m, xr[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float e = expf(xr[i] - m);
    xr[i] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < ld; i += 256) xr[i] = (i < n) ? xr[i] * inv : 0.f;
}

// ---------------------------------------------------------------- 1x1x1 conv (K = initial_dim) + activation
// input_preprocess, perceiver_lang_io.py:217-220,357.  x [M, Cin] -> y [M, Cout]; 4 threads per voxel.
template <int CIN>
static __global__ void __launch_bounds__(256)
pointwise_conv_kernel(const float* __restrict__ x, const float* __restrict__ w /*[Cout,CIN]*/,
                      const float* __restrict__ bias, float* __restrict__ y, size_t M, int Cout,
                      float slope) {
  extern __shared__ float sw[];  // [Cout][CIN] + [Cout]
  for (int i = threadIdx.x; i < Cout * CIN; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[Cout * CIN + i] = bias[i];
  __syncthreads();
  const int per = Cout / 4;  // channels per thread (multiple of 4)
  const size_t m = (size_t)blockIdx.x * 64 + (threadIdx.x >> 2);
  const int part = threadIdx.x & 3;
  if (m >= M) return;
  float in[CIN];
#pragma unroll
  for (int i = 0; i < CIN; ++i) in[i] = x[m * CIN + i];
  float* yo = y + m * Cout + part * per;
  for (int c = 0; c < per; c += 4) {
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = part * per + c + j;
      float a = sw[Cout * CIN + co];
#pragma unroll
      for (int i = 0; i < CIN; ++i) a = fmaf(in[i], sw[co * CIN + i], a);
      o[j] = slope >= 0.f ? lrelu(a, slope) : a;
    }
    *reinterpret_cast<float4*>(yo + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------- 3x3x3 conv to ONE channel (trans_decoder)
// perceiver_lang_io.py:308-311,465.  x [B,V,V,V,C] channels-last, w re-laid [27][C]; y [B,V^3].
// One warp per output voxel group of 4: each lane owns C/32 channels... simple version: 8 lanes per voxel.
template <int C>
static __global__ void __launch_bounds__(256)
conv3_to1_kernel(const float* __restrict__ x, const float* __restrict__ wt /*[27][C]*/,
                 const float* __restrict__ bias, float* __restrict__ y, int B, int V) {
  __shared__ __align__(16) float sw[27 * C];
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = wt[i];
  __syncthreads();
  constexpr int LPV = 8;              // lanes per voxel
  constexpr int CPL = C / LPV;        // channels per lane (8 for C=64)
  const size_t V3 = (size_t)V * V * V;
  const size_t vox = (size_t)blockIdx.x * (256 / LPV) + (threadIdx.x / LPV);
  const int part = threadIdx.x % LPV;
  const bool valid = vox < (size_t)B * V3;
  float acc = 0.f;
  if (valid) {
    const int b = (int)(vox / V3);
    const int r = (int)(vox % V3);
    const int d = r / (V * V), h = (r / V) % V, w = r % V;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
      const int id = min(max(d + dz, 0), V - 1);
#pragma unroll 1
      for (int dy = -1; dy <= 1; ++dy) {
        const int ih = min(max(h + dy, 0), V - 1);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int iw = min(max(w + dx, 0), V - 1);
          const float* xp = x + ((((size_t)b * V + id) * V + ih) * V + iw) * C + part * CPL;
          const float* wp = sw + ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)) * C + part * CPL;
#pragma unroll
          for (int c = 0; c < CPL; c += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xp + c);
            const float4 wv = *reinterpret_cast<const float4*>(wp + c);
            acc = fmaf(xv.x, wv.x, acc);
            acc = fmaf(xv.y, wv.y, acc);
            acc = fmaf(xv.z, wv.z, acc);
            acc = fmaf(xv.w, wv.w, acc);
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = LPV / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (valid && part == 0) y[vox] = acc + bias[0];
}

// ---------------------------------------------------------------- spatial soft-argmax + max
// SpatialSoftmax3D (network_utils.py:773-809, temperature 0.01) and AdaptiveMaxPool3d(1), over
// channels-last x [B,P,C].  Pass 1: each block reduces a chunk of positions for all channels with
// an online softmax (running max, rescaled sums); pass 2 merges the chunks.
// partial layout: [B][chunks][6][C] = (max of x/T, sum, sx, sy, sz, max of x)
__device__ __forceinline__ float lin_coord(int i, int n) {
  // np.linspace(-1, 1, n)[i] evaluated in double then rounded to fp32 (network_utils.py:783-792)
  if (n == 1) return -1.f;
  if (i == n - 1) return 1.f;
  return (float)(-1.0 + (double)i * (2.0 / (double)(n - 1)));
}

static __global__ void __launch_bounds__(256)
spatial_softmax_partial_kernel(const float* __restrict__ x, int P, int C, int Dd, int Hh, int Ww,
                               int chunk, float* __restrict__ partial) {
  const int b = blockIdx.y, ck = blockIdx.x, chunks = gridDim.x;
  const int lanes_p = 256 / C;  // position lanes per block (C = 64 -> 4, 128 -> 2, 192 -> 1)
  const int c = threadIdx.x % C, pl = threadIdx.x / C;
  const int p_begin = ck * chunk, p_end = min(P, p_begin + chunk);
  float m = -INFINITY, s = 0.f, sx = 0.f, sy = 0.f, sz = 0.f, rawm = -INFINITY;
  if (pl < lanes_p) {
    for (int p = p_begin + pl; p < p_end; p += lanes_p) {
      const float raw = x[((size_t)b * P + p) * C + c];
      rawm = fmaxf(rawm, raw);
      const float v = __fdiv_rn(raw, 0.01f);  // feature / temperature
      const int d = p / (Hh * Ww), h = (p / Ww) % Hh, w = p % Ww;
      // meshgrid(indexing='xy') quirk: pos_x varies along tensor axis H, pos_y along D, pos_z along W
      const float px = lin_coord(h, Hh), py = lin_coord(d, Dd), pz = lin_coord(w, Ww);
      if (v > m) {
        const float sc = expf(m - v);  // exp(-inf) = 0 on the first element
        s *= sc; sx *= sc; sy *= sc; sz *= sc;
        m = v;
      }
      const float e = expf(v - m);
      s += e;
      sx = fmaf(e, px, sx);
      sy = fmaf(e, py, sy);
      sz = fmaf(e, pz, sz);
    }
  }
  __shared__ float sm[6][256];
  sm[0][threadIdx.x] = m; sm[1][threadIdx.x] = s; sm[2][threadIdx.x] = sx;
  sm[3][threadIdx.x] = sy; sm[4][threadIdx.x] = sz; sm[5][threadIdx.x] = rawm;
  __syncthreads();
  if (threadIdx.x < C) {
    float M = -INFINITY;
    for (int l = 0; l < lanes_p; ++l) M = fmaxf(M, sm[0][l * C + c]);
    float S = 0.f, SX = 0.f, SY = 0.f, SZ = 0.f, RM = -INFINITY;
    for (int l = 0; l < lanes_p; ++l) {
      RM = fmaxf(RM, sm[5][l * C + c]);
      const float ml = sm[0][l * C + c];
      const float sc = (ml == -INFINITY) ? 0.f : expf(ml - M);
      S += sm[1][l * C + c] * sc;
      SX += sm[2][l * C + c] * sc;
      SY += sm[3][l * C + c] * sc;
      SZ += sm[4][l * C + c] * sc;
    }
    float* o = partial + (((size_t)b * chunks + ck) * 6) * C + c;
    o[0] = M; o[C] = S; o[2 * C] = SX; o[3 * C] = SY; o[4 * C] = SZ; o[5 * C] = RM;
  }
}

static __global__ void __launch_bounds__(256)
spatial_softmax_merge_kernel(const float* __restrict__ partial, int chunks, int C,
                             float* __restrict__ ss, int ss_stride, float* __restrict__ mx,
                             int mx_stride) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* p = partial + ((size_t)b * chunks * 6) * C + c;
    float M = -INFINITY, RM = -INFINITY;
    for (int k = 0; k < chunks; ++k) M = fmaxf(M, p[(size_t)k * 6 * C]);
    float S = 0.f, SX = 0.f, SY = 0.f, SZ = 0.f;
    for (int k = 0; k < chunks; ++k) {
      const float* q = p + (size_t)k * 6 * C;
      RM = fmaxf(RM, q[5 * C]);
      const float sc = (q[0] == -INFINITY) ? 0.f : expf(q[0] - M);
      S += q[C] * sc; SX += q[2 * C] * sc; SY += q[3 * C] * sc; SZ += q[4 * C] * sc;
    }
    float* o = ss + (size_t)b * ss_stride + c * 3;
    o[0] = SX / S; o[1] = SY / S; o[2] = SZ / S;
    if (mx) mx[(size_t)b * mx_stride + c] = RM;  // AdaptiveMaxPool3d(1)
  }
}

// ---------------------------------------------------------------- token assembly
// ins_seq[b, j, :]   (j < nl)  = lang_lin[b, j, :] + pos[j, :]
// ins_seq[b, nl+t, :]          = concat(patch[b, t, :im], p[b, :], (p2[b, :])) + pos[nl+t, :]
// reference perceiver_lang_io.py:370-373,389,412,417-422
static __global__ void __launch_bounds__(256)
assemble_tokens_kernel(const float* __restrict__ lang_lin, const float* __restrict__ patch,
                       const float* __restrict__ p1, const float* __restrict__ p2,
                       const float* __restrict__ pos, float* __restrict__ out, int B, int nl,
                       int T, int C, int im) {
  const size_t total = (size_t)B * (nl + T) * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t row = i / C;
    const int j = (int)(row % (nl + T));
    const int b = (int)(row / (nl + T));
    float v;
    if (j < nl) {
      v = lang_lin[((size_t)b * nl + j) * C + c];
    } else if (c < im) {
      v = patch[((size_t)b * T + (j - nl)) * im + c];
    } else if (c < 2 * im) {
      v = p1[(size_t)b * im + (c - im)];
    } else {
      v = p2[(size_t)b * im + (c - 2 * im)];
    }
    out[i] = v + pos[(size_t)j * C + c];
  }
}

// ---------------------------------------------------------------- weight preparation
// conv weight [Co,Ci,k,k,k] (PyTorch) -> tap-major [Co][k^3][Ci]
static __global__ void conv_weight_to_tapmajor_kernel(const float* __restrict__ w, float* __restrict__ o,
                                               int Co, int Ci, int k3) {
  const size_t total = (size_t)Co * Ci * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int t = (int)((i / Ci) % k3);
    const int co = (int)(i / ((size_t)Ci * k3));
    o[i] = w[((size_t)co * Ci + ci) * k3 + t];
  }
}

// Fold  conv_k(replicate pad) o upsample_s(trilinear, align_corners=False)  into s^3 polyphase
// 3x3x3 kernels on the low-resolution grid (valid when pad <= (s+1)/2):
//   out[s*q + r] = sum_{n in {-1,0,1}^3} Weff[r][co][n][ci] * low[clamp(q + n)]
//   Weff[r][co][n][ci] = sum_t W[co,ci,t] * prod_axis U(r_a + t_a - pad, n_a)
// where U(i, n) is the weight of low[q+n] in the up-sampled value at fine offset i from s*q.
// reference: Conv3DUpsampleBlock, network_utils.py:237-254
__device__ __forceinline__ float upsample_tap_weight(int i /*fine offset from s*q*/, int n, int s) {
  // src = q + (i + 0.5)/s - 0.5 ; j0 = floor(src) ; lambda = src - j0
  const float src = ((float)i + 0.5f) / (float)s - 0.5f;
  const float f = floorf(src);
  const int j0 = (int)f;
  const float lam = src - f;
  float wgt = 0.f;
  if (n == j0) wgt += 1.f - lam;
  if (n == j0 + 1) wgt += lam;
  return wgt;
}
static __global__ void fold_upconv_weights_kernel(const float* __restrict__ w /*[Co,Ci,k,k,k]*/,
                                           float* __restrict__ o /*[s^3][Co][27][Ci]*/, int Co,
                                           int Ci, int k, int s) {
  const int pad = k / 2;
  const size_t total = (size_t)s * s * s * Co * 27 * Ci;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int nb = (int)((i / Ci) % 27);
    const int co = (int)((i / ((size_t)Ci * 27)) % Co);
    const int ph = (int)(i / ((size_t)Ci * 27 * Co));
    const int rd = ph / (s * s), rh = (ph / s) % s, rw = ph % s;
    const int nd = nb / 9 - 1, nh = (nb / 3) % 3 - 1, nw = nb % 3 - 1;
    const float* wp = w + ((size_t)co * Ci + ci) * k * k * k;
    float acc = 0.f;
    for (int td = 0; td < k; ++td) {
      const float ud = upsample_tap_weight(rd + td - pad, nd, s);
      if (ud == 0.f) continue;
      for (int th = 0; th < k; ++th) {
        const float uh = upsample_tap_weight(rh + th - pad, nh, s);
        if (uh == 0.f) continue;
        for (int tw = 0; tw < k; ++tw) {
          const float uw = upsample_tap_weight(rw + tw - pad, nw, s);
          if (uw == 0.f) continue;
          acc += wp[(td * k + th) * k + tw] * (ud * uh * uw);
        }
      }
    }
    o[i] = acc;
  }
}

}  // namespace vxb

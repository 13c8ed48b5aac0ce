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

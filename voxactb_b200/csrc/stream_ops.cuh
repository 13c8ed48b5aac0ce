// HBM-bound streaming kernels of the Q-network: spatial soft-argmax + global max pooling in one pass,
// and the 3x3x3 conv to ONE channel (trans_decoder).  Every kernel reads its 256 MB/sample input once.
#pragma once
#include "planes16.cuh"
#include "common.cuh"

namespace vxb {

// ------------------------------------------------------------------------------------------------
// SpatialSoftmax3D (network_utils.py:773-809, temperature 0.01) + AdaptiveMaxPool3d(1) over
// channels-last x [B, P, C].
//   pass 1 (ss_partial_kernel): grid (chunks, B); a thread owns 4 channels and walks the positions of
//       its chunk with 128-bit loads, keeping an online softmax in the log2 domain
//       (w = 2^(x*k - m), k = log2(e)/T) with running (sum, sum*px, sum*py, sum*pz) and the raw max.
//   pass 2 (ss_merge_kernel): merges the chunk partials.
// partial layout: [B][chunks][6][C] = (m (log2 domain), s, sx, sy, sz, raw max)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ss_lin_coord(int i, int n) {
  // np.linspace(-1, 1, n)[i] evaluated in double then rounded to fp32 (network_utils.py:783-792)
  if (n == 1) return -1.f;
  if (i == n - 1) return 1.f;
  return (float)(-1.0 + (double)i * (2.0 / (double)(n - 1)));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kSSLog2eOverT = 144.26950408889634f;  // log2(e) / 0.01

struct SSState {
  float m, s, sx, sy, sz, rm;
};
__device__ __forceinline__ void ss_update(SSState& st, float raw, float px, float py, float pz) {
  st.rm = fmaxf(st.rm, raw);
  const float t = raw * kSSLog2eOverT;
  if (t > st.m) {
    const float sc = fast_exp2(st.m - t);  // 2^(-inf) = 0 on the first element
    st.s *= sc; st.sx *= sc; st.sy *= sc; st.sz *= sc;
    st.m = t;
  }
  const float e = fast_exp2(t - st.m);
  st.s += e;
  st.sx = fmaf(e, px, st.sx);
  st.sy = fmaf(e, py, st.sy);
  st.sz = fmaf(e, pz, st.sz);
}
// Four channels at one position, lazily rescaled: the reference exponent m only moves when a value exceeds it by
// more than kSSLazySlack (log2 domain), so e = 2^(t - m) <= 2^slack and the sums stay far inside the fp32 range
// (<= 2^20 positions x 2^64); the common path is branch-free and the sums run as packed FFMA2 over channel pairs.
constexpr float kSSLazySlack = 64.f;
__device__ __forceinline__ void ss_update4_lazy(SSState (&st)[4], const float (&raw)[4], float px, float py, float pz) {
  float t[4];
  bool grow = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    st[j].rm = fmaxf(st[j].rm, raw[j]);
    t[j] = raw[j] * kSSLog2eOverT;
    grow |= t[j] > st[j].m + kSSLazySlack;      // m = -inf on the first element
  }
  // lanes of a warp may leave the position loop one iteration apart: vote over the lanes that are here (a lane's own
  // flag is always part of its vote, so any subset is correct)
  if (__any_sync(__activemask(), grow)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (t[j] > st[j].m) {
        const float sc = fast_exp2(st[j].m - t[j]);
        st[j].s *= sc; st[j].sx *= sc; st[j].sy *= sc; st[j].sz *= sc;
        st[j].m = t[j];
      }
    }
  }
  const float2 pxx = make_float2(px, px), pyy = make_float2(py, py), pzz = make_float2(pz, pz);
#pragma unroll
  for (int j = 0; j < 4; j += 2) {
    const float2 e = make_float2(fast_exp2(t[j] - st[j].m), fast_exp2(t[j + 1] - st[j + 1].m));
    const float2 s = __fadd2_rn(make_float2(st[j].s, st[j + 1].s), e);
    const float2 sx = __ffma2_rn(e, pxx, make_float2(st[j].sx, st[j + 1].sx));
    const float2 sy = __ffma2_rn(e, pyy, make_float2(st[j].sy, st[j + 1].sy));
    const float2 sz = __ffma2_rn(e, pzz, make_float2(st[j].sz, st[j + 1].sz));
    st[j].s = s.x; st[j + 1].s = s.y;
    st[j].sx = sx.x; st[j + 1].sx = sx.y;
    st[j].sy = sy.x; st[j + 1].sy = sy.y;
    st[j].sz = sz.x; st[j + 1].sz = sz.y;
  }
}
__device__ __forceinline__ void ss_merge(SSState& a, const SSState& b) {
  a.rm = fmaxf(a.rm, b.rm);
  const float M = fmaxf(a.m, b.m);
  const float sa = (a.m == -INFINITY) ? 0.f : fast_exp2(a.m - M);
  const float sb = (b.m == -INFINITY) ? 0.f : fast_exp2(b.m - M);
  a.s = a.s * sa + b.s * sb;
  a.sx = a.sx * sa + b.sx * sb;
  a.sy = a.sy * sa + b.sy * sb;
  a.sz = a.sz * sa + b.sz * sb;
  a.m = M;
}

constexpr int SS_THREADS = 256;

static __global__ void __launch_bounds__(SS_THREADS)
ss_partial_kernel(const float* __restrict__ x, int P, int C, int Dd, int Hh, int Ww, int chunk,
                  float* __restrict__ partial) {
  extern __shared__ float ss_smem[];  // coordinate LUTs [Dd + Hh + Ww], then the reduction buffer
  const int b = blockIdx.y, ck = blockIdx.x, chunks = gridDim.x;
  const int G = C >> 2;                        // float4 groups per position
  const int PL = SS_THREADS / G;               // positions per block iteration
  float* lut = ss_smem;
  for (int i = threadIdx.x; i < Dd + Hh + Ww; i += SS_THREADS) {
    lut[i] = i < Dd ? ss_lin_coord(i, Dd) : (i < Dd + Hh ? ss_lin_coord(i - Dd, Hh) : ss_lin_coord(i - Dd - Hh, Ww));
  }
  __syncthreads();
  const float* lutD = lut;
  const float* lutH = lut + Dd;
  const float* lutW = lut + Dd + Hh;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  SSState st[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) st[j] = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
  const int p_begin = ck * chunk, p_end = min(P, p_begin + chunk);
  if (pl < PL) {
    int p = p_begin + pl;
    int d = p / (Hh * Ww), h = (p / Ww) % Hh, w = p % Ww;
    const float4* xp = reinterpret_cast<const float4*>(x + ((size_t)b * P) * C) + g;
    for (; p < p_end; p += PL) {
      const float4 v = __ldg(xp + (size_t)p * G);
      // meshgrid(indexing='xy') quirk: pos_x varies along tensor axis H, pos_y along D, pos_z along W
      const float px = lutH[h], py = lutD[d], pz = lutW[w];
      ss_update(st[0], v.x, px, py, pz);
      ss_update(st[1], v.y, px, py, pz);
      ss_update(st[2], v.z, px, py, pz);
      ss_update(st[3], v.w, px, py, pz);
      w += PL;
      while (w >= Ww) {
        w -= Ww;
        if (++h >= Hh) { h = 0; ++d; }
      }
    }
  }
  // combine the PL position lanes of each channel group through shared memory
  float* red = ss_smem + (Dd + Hh + Ww);       // [PL][G][4][6]
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* r = red + (((size_t)pl * G + g) * 4 + j) * 6;
      r[0] = st[j].m; r[1] = st[j].s; r[2] = st[j].sx; r[3] = st[j].sy; r[4] = st[j].sz; r[5] = st[j].rm;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += SS_THREADS) {
    SSState a = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
    for (int l = 0; l < PL; ++l) {
      const float* r = red + (((size_t)l * G + (c >> 2)) * 4 + (c & 3)) * 6;
      const SSState o = {r[0], r[1], r[2], r[3], r[4], r[5]};
      ss_merge(a, o);
    }
    float* o = partial + (((size_t)b * chunks + ck) * 6) * C + c;
    o[0] = a.m; o[C] = a.s; o[2 * C] = a.sx; o[3 * C] = a.sy; o[4 * C] = a.sz; o[5 * C] = a.rm;
  }
}

// grid (ceil(C/32), B, splits), 256 threads = 32 channels x 8 chunk lanes.  splits == 1: final result (soft-argmax
// coordinates + max).  splits > 1: block z folds its share of the chunks into partial_out [B][splits][6][C] (same
// layout, merged by a second launch) -- the fused conv tail leaves thousands of chunk partials per sample.
static __global__ void __launch_bounds__(256)
ss_merge_kernel(const float* __restrict__ partial, int chunks, int C, float* __restrict__ ss,
                int ss_stride, float* __restrict__ mx, int mx_stride, float* __restrict__ partial_out,
                float* __restrict__ stats /* optional [B][2][C]: (m, s) of the softmax, kept for the backward */) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int kl = threadIdx.x >> 5;
  const int splits = gridDim.z;
  const int per = (chunks + splits - 1) / splits;
  const int k_lo = blockIdx.z * per, k_hi = min(chunks, k_lo + per);
  SSState a = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
  if (c < C) {
    for (int k = k_lo + kl; k < k_hi; k += 8) {
      const float* q = partial + (((size_t)b * chunks + k) * 6) * C + c;
      const SSState o = {q[0], q[C], q[2 * C], q[3 * C], q[4 * C], q[5 * C]};
      ss_merge(a, o);
    }
  }
  __shared__ float red[8][32][6];
  float* r = red[kl][threadIdx.x & 31];
  r[0] = a.m; r[1] = a.s; r[2] = a.sx; r[3] = a.sy; r[4] = a.sz; r[5] = a.rm;
  __syncthreads();
  if (kl == 0 && c < C) {
    for (int l = 1; l < 8; ++l) {
      const float* q = red[l][threadIdx.x];
      const SSState o = {q[0], q[1], q[2], q[3], q[4], q[5]};
      ss_merge(a, o);
    }
    if (partial_out) {
      float* o = partial_out + (((size_t)b * splits + blockIdx.z) * 6) * C + c;
      o[0] = a.m; o[C] = a.s; o[2 * C] = a.sx; o[3 * C] = a.sy; o[4 * C] = a.sz; o[5 * C] = a.rm;
    } else {
      float* o = ss + (size_t)b * ss_stride + c * 3;
      o[0] = a.sx / a.s; o[1] = a.sy / a.s; o[2] = a.sz / a.s;
      if (mx) mx[(size_t)b * mx_stride + c] = a.rm;  // AdaptiveMaxPool3d(1)
      if (stats) { stats[((size_t)b * 2) * C + c] = a.m; stats[((size_t)b * 2 + 1) * C + c] = a.s; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// input_preprocess (Conv3d CIN -> C, k=1, + LeakyReLU; perceiver_lang_io.py:217-220,357) fused with
// ss0 + global max pool (:360): same thread layout as ss_partial_kernel (a thread owns 4 output channels,
// their CIN x 4 weights live in registers as (even, odd) pairs for packed FFMA2), the activation is stored once --
// as fp32 and / or directly as the padded 16-bit hi/lo planes the tensor-core kernels read -- and its soft-argmax /
// max partials are accumulated on the fly (lazily rescaled online softmax): d0 is never re-read.  The 40-byte input
// rows, shared by the 16 threads of a position, are staged through shared memory with cp.async two tiles ahead.
// ------------------------------------------------------------------------------------------------
constexpr int IPP_ITERS = 8;     // positions per thread and tile
constexpr int IPP_STAGES = 3;    // cp.async stages (prefetch distance 2 tiles)
__host__ __device__ inline int ipp_stage_offset_floats(int lut_floats, int lanes) {
  return (lut_floats + lanes * 24 + 3) & ~3;   // after the LUTs and the [PL][G][4][6] reduction buffer, 16-byte aligned
}
template <int CIN>
static __global__ void __launch_bounds__(SS_THREADS)
input_preprocess_ss_kernel(const float* __restrict__ x /*[B,P,CIN]*/, const float* __restrict__ w /*[C,CIN]*/,
                           const float* __restrict__ bias, float slope, float* __restrict__ y /*[B,P,C]*/,
                           int P, int C, int Dd, int Hh, int Ww, int chunk, float* __restrict__ partial,
                           __nv_bfloat16* __restrict__ phi, __nv_bfloat16* __restrict__ plo /*padded planes or null*/,
                           uint8_t* __restrict__ pc8 /*c8 plane of conv_f8c.cuh or null*/, const float* __restrict__ f8a) {
  extern __shared__ float ss_smem[];
  const int b = blockIdx.y, ck = blockIdx.x, chunks = gridDim.x;
  const int G = C >> 2;
  const int PL = SS_THREADS / G;
  float* lut = ss_smem;
  for (int i = threadIdx.x; i < Dd + Hh + Ww; i += SS_THREADS) {
    lut[i] = i < Dd ? ss_lin_coord(i, Dd) : (i < Dd + Hh ? ss_lin_coord(i - Dd, Hh) : ss_lin_coord(i - Dd - Hh, Ww));
  }
  __syncthreads();
  const float* lutD = lut;
  const float* lutH = lut + Dd;
  const float* lutW = lut + Dd + Hh;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  SSState st[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) st[j] = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
  const int p_begin = ck * chunk, p_end = min(P, p_begin + chunk);
  // The input rows of a tile of IPP_ITERS * PL consecutive positions are staged in shared memory with cp.async, two
  // tiles ahead of the one being computed (the 40-byte rows are shared by the G threads of a position; loading them
  // straight into registers one position ahead left the kernel waiting on DRAM latency).
  const int TP = IPP_ITERS * PL;
  float* stages = ss_smem + ipp_stage_offset_floats(Dd + Hh + Ww, PL * G);
  const float* xb = x + (size_t)b * P * CIN;
  const int tiles = (p_end - p_begin + TP - 1) / TP;
  auto issue = [&](int t) {
    if (t < tiles) {
      const int tp0 = p_begin + t * TP;
      const int n8 = min(TP, p_end - tp0) * (CIN / 2);            // 8-byte units (rows are CIN * 4 = 8-byte multiples)
      const float* src = xb + (size_t)tp0 * CIN;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stages + (size_t)(t % IPP_STAGES) * TP * CIN);
      for (int k = threadIdx.x; k < n8; k += SS_THREADS)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * k), "l"(src + 2 * k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  issue(1);
  // weights as (even, odd) input-channel pairs: the dot product runs as CIN/2 packed FFMA2 per output channel
  static_assert(CIN % 2 == 0, "input rows are read as float2");
  float2 wr[4][CIN / 2];
  float br[4];
  int p = p_begin + pl;
  int d = p / (Hh * Ww), h = (p / Ww) % Hh, wv = p % Ww;
  float4* yb = reinterpret_cast<float4*>(y + (size_t)b * P * C) + g;
  // hi/lo planes of the replicate-padded grid [B, D+2, H+2, W+2, C] (interior; the halo is filled afterwards):
  // the padded row offset is carried along with (d, h, wv) instead of being recomputed
  const size_t pbase = (size_t)b * (Dd + 2) * (Hh + 2) * (Ww + 2) * C + g * 4;   // this sample, this channel group
  __nv_bfloat16* phb = phi ? phi + pbase : nullptr;
  __nv_bfloat16* plb = plo ? plo + pbase : nullptr;
  // c8 plane: the row's 128 bytes are two 32-channel blocks of [32 x e4m3(2^11 a lo) | 32 x e4m3(a x)]; this thread's 4 channels
  uint8_t* pcb = pc8 ? pc8 + ((size_t)b * (Dd + 2) * (Hh + 2) * (Ww + 2) * C) * 2 + (g >> 3) * 64 + (g & 7) * 4 : nullptr;
  const float fa = pc8 ? __ldg(f8a) : 0.f;
  // element offset of this position's padded row inside the sample (< 2^31 for any grid that fits the planes)
  uint32_t poff = ((uint32_t)((d + 1) * (Hh + 2) + h + 1) * (uint32_t)(Ww + 2) + (uint32_t)wv + 1u) * (uint32_t)C;
  const uint32_t step_w = (uint32_t)PL * C, wrap_w = 2u * C, wrap_h = 2u * (uint32_t)(Ww + 2) * C;
  const float sl = slope >= 0.f ? slope : 1.f;          // max(v, v * sl) == LeakyReLU for 0 <= sl <= 1, identity for sl = 1
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      br[j] = bias[g * 4 + j];
#pragma unroll
      for (int i = 0; i < CIN / 2; ++i)
        wr[j][i] = make_float2(w[(g * 4 + j) * CIN + 2 * i], w[(g * 4 + j) * CIN + 2 * i + 1]);
    }
  }
  for (int t = 0; t < tiles; ++t) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");     // this thread's copies of tile t have landed
    __syncthreads();                                          // everyone's have, and tile t-1 is fully consumed
    issue(t + 2);                                             // into the stage tile t-1 used
    if (pl < PL) {
      const float2* xs = reinterpret_cast<const float2*>(stages + (size_t)(t % IPP_STAGES) * TP * CIN) + pl * (CIN / 2);
#pragma unroll 2
      for (int it = 0; it < IPP_ITERS; ++it) {
        if (p >= p_end) break;
        float2 in[CIN / 2];
#pragma unroll
        for (int i = 0; i < CIN / 2; ++i) in[i] = xs[(size_t)it * PL * (CIN / 2) + i];
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 a = make_float2(br[j], 0.f);
#pragma unroll
          for (int i = 0; i < CIN / 2; ++i) a = __ffma2_rn(in[i], wr[j][i], a);
          const float v = a.x + a.y;
          o[j] = fmaxf(v, v * sl);
        }
        if (y) yb[(size_t)p * G] = make_float4(o[0], o[1], o[2], o[3]);
        if (phb) {
          const __nv_bfloat162 h01 = pl2_from_floats(o[0], o[1]), h23 = pl2_from_floats(o[2], o[3]);
          const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
          const __nv_bfloat162 l01 = pl2_from_floats(o[0] - f01.x, o[1] - f01.y);
          const __nv_bfloat162 l23 = pl2_from_floats(o[2] - f23.x, o[3] - f23.y);
          uint2 hv, lv;
          hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
          lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(phb + poff) = hv;
          *reinterpret_cast<uint2*>(plb + poff) = lv;
          if (pcb) {
            uint8_t* rowb = pcb + (size_t)poff * 2;        // poff counts 2-byte elements of a C-channel row
            *reinterpret_cast<uint32_t*>(rowb) = pl_e4m3x4(o[0] - f01.x, o[1] - f01.y, o[2] - f23.x, o[3] - f23.y, fa * 2048.f);
            *reinterpret_cast<uint32_t*>(rowb + 32) = pl_e4m3x4(o[0], o[1], o[2], o[3], fa);
          }
        }
        ss_update4_lazy(st, o, lutH[h], lutD[d], lutW[wv]);
        p += PL; wv += PL; poff += step_w;
        while (wv >= Ww) {
          wv -= Ww; poff += wrap_w;
          if (++h >= Hh) { h = 0; ++d; poff += wrap_h; }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  float* red = ss_smem + (Dd + Hh + Ww);
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* r = red + (((size_t)pl * G + g) * 4 + j) * 6;
      r[0] = st[j].m; r[1] = st[j].s; r[2] = st[j].sx; r[3] = st[j].sy; r[4] = st[j].sz; r[5] = st[j].rm;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += SS_THREADS) {
    SSState a = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
    for (int l = 0; l < PL; ++l) {
      const float* r = red + (((size_t)l * G + (c >> 2)) * 4 + (c & 3)) * 6;
      const SSState o = {r[0], r[1], r[2], r[3], r[4], r[5]};
      ss_merge(a, o);
    }
    float* o = partial + (((size_t)b * chunks + ck) * 6) * C + c;
    o[0] = a.m; o[C] = a.s; o[2 * C] = a.sx; o[3 * C] = a.sy; o[4 * C] = a.sz; o[5 * C] = a.rm;
  }
}

inline int ss_num_chunks(size_t P, int B) {
  // ~8 blocks per SM over the whole launch, at least 512 positions per block
  const size_t want = (size_t)(148 * 8 + B - 1) / (size_t)B;
  const size_t cap = std::max<size_t>(1, P / 512);
  return (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(want, cap), 1024));
}
inline size_t ss_partial_floats(size_t P, int B, int C) { return (size_t)B * ss_num_chunks(P, B) * 6 * C; }

// launches 2 kernels
inline int spatial_softmax_run(const float* x, int B, int Dd, int Hh, int Ww, int C, float* ss, int ss_stride,
                               float* mx, int mx_stride, float* partial, cudaStream_t st, float* stats = nullptr) {
  VXB_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 1024, "spatial_softmax: C=%d must be a multiple of 4, <= 1024", C);
  VXB_CHECK_ARG(((uintptr_t)x & 15) == 0, "spatial_softmax: input must be 16-byte aligned");
  const size_t P = (size_t)Dd * Hh * Ww;
  const int chunks = ss_num_chunks(P, B);
  const int chunk = (int)((P + chunks - 1) / chunks);
  const int G = C / 4, PL = SS_THREADS / G;
  const size_t smem = ((size_t)(Dd + Hh + Ww) + (size_t)PL * G * 24) * sizeof(float);
  ss_partial_kernel<<<dim3(chunks, B), SS_THREADS, smem, st>>>(x, (int)P, C, Dd, Hh, Ww, chunk, partial);
  VXB_LAUNCH_CHECK();
  ss_merge_kernel<<<dim3(cdiv(C, 32), B), 256, 0, st>>>(partial, chunks, C, ss, ss_stride, mx, mx_stride, nullptr, stats);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// d0 = act(conv1x1(x)); ss = soft-argmax(d0), mx = max(d0); launches 2 kernels
template <int CIN>
inline int input_preprocess_ss_run(const float* x, const float* w, const float* bias, float slope, float* y, int B,
                                   int Dd, int Hh, int Ww, int C, float* ss, int ss_stride, float* mx, int mx_stride,
                                   float* partial, cudaStream_t st, __nv_bfloat16* phi = nullptr,
                                   __nv_bfloat16* plo = nullptr, float* stats = nullptr, uint8_t* pc8 = nullptr,
                                   const float* f8a = nullptr) {
  VXB_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 1024, "input_preprocess: C=%d must be a multiple of 4, <= 1024", C);
  VXB_CHECK_ARG(slope <= 1.f, "input_preprocess: LeakyReLU slope %g must be <= 1 (negative = no activation)", (double)slope);
  VXB_CHECK_ARG(!pc8 || (C == 64 && phi && plo && f8a), "input_preprocess: the c8 plane needs C = 64 and the hi/lo planes");
  const size_t P = (size_t)Dd * Hh * Ww;
  const int chunks = ss_num_chunks(P, B);
  const int chunk = (int)((P + chunks - 1) / chunks);
  const int G = C / 4, PL = SS_THREADS / G;
  const size_t smem = ((size_t)ipp_stage_offset_floats(Dd + Hh + Ww, PL * G) + (size_t)IPP_STAGES * IPP_ITERS * PL * CIN) * sizeof(float);
  VXB_CHECK_ARG(smem <= 48 * 1024, "input_preprocess: %zu bytes of shared memory needed (C=%d, grid %dx%dx%d)", smem, C, Dd, Hh, Ww);
  input_preprocess_ss_kernel<CIN><<<dim3(chunks, B), SS_THREADS, smem, st>>>(x, w, bias, slope, y, (int)P, C, Dd, Hh, Ww,
                                                                            chunk, partial, phi, plo, pc8, f8a);
  VXB_LAUNCH_CHECK();
  ss_merge_kernel<<<dim3(cdiv(C, 32), B), 256, 0, st>>>(partial, chunks, C, ss, ss_stride, mx, mx_stride, nullptr, stats);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------
// trans_decoder: Conv3d(64 -> 1, k=3, replicate pad 1), no activation (perceiver_lang_io.py:308-311,465)
// over channels-last u [B, V, V, V, C].
//
// A CTA owns a (TY x TX) output tile and marches along z.  For every input plane each thread takes ONE
// voxel of the haloed (TY+2) x (TX+2) tile (coordinates clamped = replicate padding), holds its C
// channels in registers and forms the 27 tap dot products p[t] = <w_t, u[v]> (weights broadcast from
// shared memory); the plane's p values go to shared memory, and every output voxel gathers its 9
// (dy, dx) neighbours for each dz into three running sums (outputs z-1, z, z+1).  u is read once
// (+ the tile halo), nothing is re-gathered per tap from L1/L2.
// ------------------------------------------------------------------------------------------------
constexpr int TS_MAX_THREADS = 640;

template <int C>
static __global__ void __launch_bounds__(TS_MAX_THREADS, 1)
trans_stencil_kernel(const float* __restrict__ u, const float* __restrict__ wt /*[27][C]*/,
                     const float* __restrict__ bias, float* __restrict__ y, int V, int TX, int TY,
                     int tiles_x, int tiles_y, int z_split) {
  extern __shared__ __align__(16) float ts_smem[];
  float* sw = ts_smem;                 // [27][C]
  float* sp = ts_smem + 27 * C;        // [27][HP] tap-major p values of the current plane
  const int HX = TX + 2, HY = TY + 2, HP = HX * HY;
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sw[i] = wt[i];
  int bid = blockIdx.x;
  const int zs = bid % z_split; bid /= z_split;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y; bid /= tiles_y;
  const int b = bid;
  const int x0 = tx * TX, y0 = ty * TY;
  const int zlen = (V + z_split - 1) / z_split;
  const int z_begin = zs * zlen, z_end = min(V, z_begin + zlen);   // output planes [z_begin, z_end)
  const int t = threadIdx.x;
  // role 1: haloed voxel
  const bool has_vox = t < HP;
  const int hy = t / HX, hx = t % HX;
  const int gy = min(max(y0 + hy - 1, 0), V - 1), gx = min(max(x0 + hx - 1, 0), V - 1);
  // role 2: output voxel
  const bool has_out = t < TX * TY;
  const int oy = t / TX, ox = t % TX;
  const bool out_ok = has_out && (y0 + oy) < V && (x0 + ox) < V;
  float acc_m1 = 0.f, acc_0 = 0.f, acc_p1 = 0.f;   // outputs zi-1, zi, zi+1 while input plane zi is processed
  const float bv = bias[0];
  __syncthreads();
  for (int zi = z_begin - 1; zi <= z_end; ++zi) {
    if (has_vox) {
      const int gz = min(max(zi, 0), V - 1);
      const float4* row = reinterpret_cast<const float4*>(u + ((((size_t)b * V + gz) * V + gy) * V + gx) * C);
      float4 r[C / 4];
#pragma unroll
      for (int j = 0; j < C / 4; ++j) r[j] = __ldg(row + j);
#pragma unroll 1
      for (int tp = 0; tp < 27; ++tp) {
        const float4* w4 = reinterpret_cast<const float4*>(sw + tp * C);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < C / 4; j += 2) {
          const float4 wa = w4[j], wb = w4[j + 1];
          a0 = fmaf(r[j].x, wa.x, a0); a0 = fmaf(r[j].y, wa.y, a0);
          a0 = fmaf(r[j].z, wa.z, a0); a0 = fmaf(r[j].w, wa.w, a0);
          a1 = fmaf(r[j + 1].x, wb.x, a1); a1 = fmaf(r[j + 1].y, wb.y, a1);
          a1 = fmaf(r[j + 1].z, wb.z, a1); a1 = fmaf(r[j + 1].w, wb.w, a1);
        }
        sp[tp * HP + t] = a0 + a1;
      }
    }
    __syncthreads();
    if (has_out) {
      // tap index = (dz+1)*9 + (dy+1)*3 + (dx+1); input plane zi feeds output zo = zi - dz
      float s[3];
#pragma unroll
      for (int dzc = 0; dzc < 3; ++dzc) {
        float a = 0.f;
#pragma unroll
        for (int dyc = 0; dyc < 3; ++dyc)
#pragma unroll
          for (int dxc = 0; dxc < 3; ++dxc)
            a += sp[(dzc * 9 + dyc * 3 + dxc) * HP + (oy + dyc) * HX + (ox + dxc)];
        s[dzc] = a;
      }
      acc_m1 += s[2];   // dz = +1 -> output zi - 1 (now complete)
      acc_0 += s[1];    // dz =  0 -> output zi
      acc_p1 += s[0];   // dz = -1 -> output zi + 1
      const int zo = zi - 1;
      if (out_ok && zo >= z_begin && zo < z_end)
        y[(((size_t)b * V + zo) * V + (y0 + oy)) * V + (x0 + ox)] = acc_m1 + bv;
      acc_m1 = acc_0; acc_0 = acc_p1; acc_p1 = 0.f;
    }
    __syncthreads();
  }
}

inline void trans_stencil_tiles(int V, int& TX, int& TY) {
  // pick the tile extents with the least padding waste whose haloed tile fits TS_MAX_THREADS threads
  const int cand_x[] = {32, 25, 20, 16, 10, 8};
  const int cand_y[] = {20, 16, 10, 8, 5, 4};
  double best = 1e30;
  TX = 8; TY = 4;
  for (int cx : cand_x)
    for (int cy : cand_y) {
      if ((cx + 2) * (cy + 2) > TS_MAX_THREADS) continue;
      const double covered = (double)cdiv(V, cx) * cx * (double)cdiv(V, cy) * cy;
      const double cost = covered / ((double)V * V) * ((double)(cx + 2) * (cy + 2) / ((double)cx * cy));
      if (cost < best - 1e-9) { best = cost; TX = cx; TY = cy; }
    }
}

template <int C>
static int trans_stencil_run(const float* u, const float* wt, const float* bias, float* y, int B, int V,
                             cudaStream_t st) {
  int TX, TY;
  trans_stencil_tiles(V, TX, TY);
  const int tiles_x = cdiv(V, TX), tiles_y = cdiv(V, TY);
  const int HP = (TX + 2) * (TY + 2);
  const int threads = (HP + 31) / 32 * 32;
  // split z so that the launch has at least ~3 CTAs per SM (each split re-reads 2 halo planes)
  int z_split = 1;
  while ((long long)B * tiles_x * tiles_y * z_split < 148 * 3 && z_split * 8 < V) z_split *= 2;
  const size_t smem = ((size_t)27 * C + (size_t)27 * HP) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    VXB_CUDA(cudaFuncSetAttribute(trans_stencil_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  trans_stencil_kernel<C><<<B * tiles_x * tiles_y * z_split, threads, smem, st>>>(u, wt, bias, y, V, TX, TY, tiles_x,
                                                                                 tiles_y, z_split);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

}  // namespace vxb

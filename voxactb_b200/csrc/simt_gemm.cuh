// fp32 FFMA tiled GEMM / implicit-GEMM convolution (math_mode VXB_MATH_FP32_SIMT).
// This is the exact-arithmetic path used for parity and for the small shapes; the big
// contractions run on tcgen05 (umma_gemm.cuh) when math_mode == VXB_MATH_F16X3.
#pragma once
#include "common.cuh"

namespace vxb {

// A_TRANS: A stored transposed, element (m, k) at A[k * lda + m] (weight gradients: dW = dY^T X without a transpose pass)
// A_CONV_T: transposed im2col, element (m = tap*Cin + c, k = output position) = x[clamp(o*stride - pad + tap)][c]
//           (convolution weight gradients: dW[tap][ci][co] = sum_rows im2col^T gz)
enum { A_PLAIN = 0, A_CONV = 1, A_TRANS = 2, A_CONV_T = 3 };
enum { B_NT = 0 /* W[N,K] row-major */, B_NN = 1 /* W[K,N] row-major */ };
// O_ATOMIC: split-K, blockIdx.z = K slice of p.kchunk elements, C += alpha * partial with atomicAdd (C zeroed by the caller)
enum { O_PLAIN = 0, O_PHASE = 1, O_ATOMIC = 2 };

struct GemmParams {
  int M, N, K;
  // ---- A operand
  const float* A;   // plain: row-major [M, lda]
  int lda;
  long long a_stride_zb, a_stride_zh;  // batch strides (blockIdx.z = zb*Hz + zh)
  // conv gather (A_CONV): rows are output positions (b, od, oh, ow) of a Do^3 grid per sample,
  // k = tap*Cin + c, tap = (dz*kk + dy)*kk + dx, input voxel = clamp(o*stride - pad + d, 0, Di-1)
  const float* src0; const float* src1;
  int C0, C1;       // channels in src0 / src1 (channels-last), Cin = C0 + C1
  int Di, Do, kk, cstride, pad;
  int zero_oob;     // A_CONV: taps outside the input grid contribute 0 instead of the clamped (replicate) voxel
  int kchunk;       // O_ATOMIC: K elements per blockIdx.z slice (multiple of GBK)
  // ---- B operand
  const float* W;
  int ldw;
  long long w_stride_zb, w_stride_zh;
  // ---- epilogue
  const float* bias;       // [N] or null
  const float* residual;   // [(m % res_rows), ldr] or null
  int res_rows, ldr;
  long long r_stride_zb;
  float alpha;
  float act_slope;         // < 0: none
  float* C;
  int ldc;
  long long c_stride_zb, c_stride_zh;
  int Hz;                  // inner batch count (heads); 1 if unused
  // O_PHASE: row m=(b,qd,qh,qw) on a Do^3 grid is written at voxel (q*ps + r) of a (Do*ps)^3 grid,
  // r = phase decoded from blockIdx.z (phase = (rd*ps + rh)*ps + rw)
  int ps;
};

constexpr int GBM = 128, GBN = 64, GBK = 16, GTHREADS = 256;

template <int AMODE, int BMODE, int OMODE>
__global__ void __launch_bounds__(GTHREADS)
simt_gemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][GBK][GBM + 4];
  __shared__ __align__(16) float Bs[2][GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
  const int z = (OMODE == O_ATOMIC) ? 0 : blockIdx.z;
  const int zb = z / p.Hz, zh = z % p.Hz;
  const float* __restrict__ Ab = (AMODE == A_PLAIN || AMODE == A_TRANS) ? p.A + zb * p.a_stride_zb + zh * p.a_stride_zh : nullptr;
  const float* __restrict__ Wb = p.W + zb * p.w_stride_zb + zh * p.w_stride_zh;
  const int k_begin = (OMODE == O_ATOMIC) ? blockIdx.z * p.kchunk : 0;
  const int k_end = (OMODE == O_ATOMIC) ? min(p.K, k_begin + p.kchunk) : p.K;

  // ---- A load assignment: 128 rows x 4 float4 per k-chunk; thread handles rows r0 and r0+64
  const int a_kq = tid & 3;
  const int a_r0 = tid >> 2;
  int a_b[2], a_d[2], a_h[2], a_w[2];
  bool a_valid[2];
  const int Cin = p.C0 + p.C1;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int m = m0 + a_r0 + i * 64;
    a_valid[i] = m < p.M;
    if (AMODE == A_CONV) {
      int mm = a_valid[i] ? m : 0;
      int Do3 = p.Do * p.Do * p.Do;
      a_b[i] = mm / Do3;
      int r = mm - a_b[i] * Do3;
      a_d[i] = (r / (p.Do * p.Do)) * p.cstride - p.pad;
      a_h[i] = ((r / p.Do) % p.Do) * p.cstride - p.pad;
      a_w[i] = (r % p.Do) * p.cstride - p.pad;
    } else {
      a_b[i] = a_d[i] = a_h[i] = a_w[i] = 0;
    }
  }
  const bool a_vec = (AMODE == A_CONV || AMODE == A_CONV_T) ? true : ((p.lda & 3) == 0 && ((size_t)Ab & 15) == 0);
  // transposed A modes: the tile is 16 k-rows x 128 m-columns, a thread loads the float4 (k = t_k + 8 i, m = t_m4 .. +3)
  const int t_k = tid >> 5, t_m4 = (tid & 31) * 4;
  int t_dz = 0, t_dy = 0, t_dx = 0, t_c = 0;
  if (AMODE == A_CONV_T) {
    const int m = min(m0 + t_m4, p.M - 1);
    const int tap = m / Cin;
    t_c = m - tap * Cin;
    t_dx = tap % p.kk; t_dy = (tap / p.kk) % p.kk; t_dz = tap / (p.kk * p.kk);
  }

  auto load_a = [&](int k0, float4 (&ra)[2]) {
    if constexpr (AMODE == A_TRANS || AMODE == A_CONV_T) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + t_k + i * 8;
        const int m = m0 + t_m4;
        if (k < k_end && m < p.M) {
          if (AMODE == A_CONV_T) {
            const int Do3 = p.Do * p.Do * p.Do;
            const int b = k / Do3;
            const int r = k - b * Do3;
            const int id = min(max((r / (p.Do * p.Do)) * p.cstride - p.pad + t_dz, 0), p.Di - 1);
            const int ih = min(max(((r / p.Do) % p.Do) * p.cstride - p.pad + t_dy, 0), p.Di - 1);
            const int iw = min(max((r % p.Do) * p.cstride - p.pad + t_dx, 0), p.Di - 1);
            const size_t vox = (((size_t)b * p.Di + id) * p.Di + ih) * p.Di + iw;
            const float* ptr = (t_c < p.C0) ? p.src0 + vox * p.C0 + t_c : p.src1 + vox * p.C1 + (t_c - p.C0);
            v = *reinterpret_cast<const float4*>(ptr);      // Cin, C0 multiples of 4: the four m share the tap
          } else {
            const float* ptr = Ab + (size_t)k * p.lda + m;
            if (a_vec && m + 3 < p.M) {
              v = *reinterpret_cast<const float4*>(ptr);
            } else {
              v.x = ptr[0];
              if (m + 1 < p.M) v.y = ptr[1];
              if (m + 2 < p.M) v.z = ptr[2];
              if (m + 3 < p.M) v.w = ptr[3];
            }
          }
        }
        ra[i] = v;
      }
    } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = k0 + a_kq * 4;
      if (a_valid[i] && k < k_end) {
        if (AMODE == A_CONV) {
          const int tap = k / Cin;
          const int c = k - tap * Cin;
          const int dx = tap % p.kk, dy = (tap / p.kk) % p.kk, dz = tap / (p.kk * p.kk);
          int id = a_d[i] + dz, ih = a_h[i] + dy, iw = a_w[i] + dx;
          const bool oob = (unsigned)id >= (unsigned)p.Di || (unsigned)ih >= (unsigned)p.Di || (unsigned)iw >= (unsigned)p.Di;
          id = min(max(id, 0), p.Di - 1); ih = min(max(ih, 0), p.Di - 1); iw = min(max(iw, 0), p.Di - 1);
          const size_t vox = (((size_t)a_b[i] * p.Di + id) * p.Di + ih) * p.Di + iw;
          const float* ptr = (c < p.C0) ? p.src0 + vox * p.C0 + c : p.src1 + vox * p.C1 + (c - p.C0);
          if (!(p.zero_oob && oob)) v = *reinterpret_cast<const float4*>(ptr);
        } else {
          const float* ptr = Ab + (size_t)(m0 + a_r0 + i * 64) * p.lda + k;
          if (a_vec && k + 3 < k_end) {
            v = *reinterpret_cast<const float4*>(ptr);
          } else {
            v.x = ptr[0];
            if (k + 1 < k_end) v.y = ptr[1];
            if (k + 2 < k_end) v.z = ptr[2];
            if (k + 3 < k_end) v.w = ptr[3];
          }
        }
      }
      ra[i] = v;
    }
    }
  };
  auto store_a = [&](int buf, const float4 (&ra)[2]) {
    if constexpr (AMODE == A_TRANS || AMODE == A_CONV_T) {
#pragma unroll
      for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&As[buf][t_k + i * 8][t_m4]) = ra[i];
    } else {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = a_r0 + i * 64;
      As[buf][a_kq * 4 + 0][r] = ra[i].x;
      As[buf][a_kq * 4 + 1][r] = ra[i].y;
      As[buf][a_kq * 4 + 2][r] = ra[i].z;
      As[buf][a_kq * 4 + 3][r] = ra[i].w;
    }
    }
  };

  // ---- B load assignment
  const bool w_vec = (p.ldw & 3) == 0 && ((size_t)Wb & 15) == 0;
  auto load_b = [&](int k0, float4& rb) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (BMODE == B_NT) {
      const int n = n0 + (tid >> 2);
      const int k = k0 + (tid & 3) * 4;
      if (n < p.N && k < k_end) {
        const float* ptr = Wb + (size_t)n * p.ldw + k;
        if (w_vec && k + 3 < k_end) {
          v = *reinterpret_cast<const float4*>(ptr);
        } else {
          v.x = ptr[0];
          if (k + 1 < k_end) v.y = ptr[1];
          if (k + 2 < k_end) v.z = ptr[2];
          if (k + 3 < k_end) v.w = ptr[3];
        }
      }
    } else {
      const int k = k0 + (tid >> 4);
      const int n = n0 + (tid & 15) * 4;
      if (k < k_end && n < p.N) {
        const float* ptr = Wb + (size_t)k * p.ldw + n;
        if (w_vec && n + 3 < p.N) {
          v = *reinterpret_cast<const float4*>(ptr);
        } else {
          v.x = ptr[0];
          if (n + 1 < p.N) v.y = ptr[1];
          if (n + 2 < p.N) v.z = ptr[2];
          if (n + 3 < p.N) v.w = ptr[3];
        }
      }
    }
    rb = v;
  };
  auto store_b = [&](int buf, const float4& rb) {
    if (BMODE == B_NT) {
      const int n = tid >> 2, kq = tid & 3;
      Bs[buf][kq * 4 + 0][n] = rb.x;
      Bs[buf][kq * 4 + 1][n] = rb.y;
      Bs[buf][kq * 4 + 2][n] = rb.z;
      Bs[buf][kq * 4 + 3][n] = rb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rb;
    }
  };

  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread tile 8 (m) x 4 (n)
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;
  load_a(k_begin, ra);
  load_b(k_begin, rb);
  store_a(0, ra);
  store_b(0, rb);
  __syncthreads();
  const int nk = (max(k_end - k_begin, 0) + GBK - 1) / GBK;
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) {
      load_a(k_begin + (kc + 1) * GBK, ra);
      load_b(k_begin + (kc + 1) * GBK, rb);
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < nk) {
      store_a(buf ^ 1, ra);
      store_b(buf ^ 1, rb);
    }
    __syncthreads();
  }

  // ---- epilogue
  float* __restrict__ Cb = p.C + zb * p.c_stride_zb + zh * p.c_stride_zh;
  const float* __restrict__ Rb = p.residual ? p.residual + zb * p.r_stride_zb : nullptr;
  int rd = 0, rh = 0, rw = 0;
  if (OMODE == O_PHASE) {
    rw = z % p.ps;
    rh = (z / p.ps) % p.ps;
    rd = z / (p.ps * p.ps);
  }
  const int n = n0 + tx * 4;
  const bool c_vec = (p.ldc & 3) == 0 && ((size_t)Cb & 15) == 0 && n + 3 < p.N;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    size_t orow = m;
    if (OMODE == O_PHASE) {
      const int Do3 = p.Do * p.Do * p.Do;
      const int b = m / Do3;
      const int r = m - b * Do3;
      const int qd = r / (p.Do * p.Do), qh = (r / p.Do) % p.Do, qw = r % p.Do;
      const int Vo = p.Do * p.ps;
      orow = (((size_t)b * Vo + qd * p.ps + rd) * Vo + qh * p.ps + rh) * Vo + qw * p.ps + rw;
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = acc[i][j] * p.alpha;
      if (n + j < p.N) {
        if (p.bias) t += p.bias[n + j];
        if (p.act_slope >= 0.f) t = lrelu(t, p.act_slope);
        if (Rb) t += Rb[(size_t)(m % p.res_rows) * p.ldr + n + j];
      }
      v[j] = t;
    }
    float* dst = Cb + orow * p.ldc + n;
    if (OMODE == O_ATOMIC) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.N) atomicAdd(dst + j, acc[i][j] * p.alpha);
      continue;
    }
    if (c_vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < p.N) dst[j] = v[j];
    }
  }
}

inline void gemm_params_init(GemmParams& p) {
  memset(&p, 0, sizeof(p));
  p.alpha = 1.f;
  p.act_slope = -1.f;
  p.Hz = 1;
  p.res_rows = 1;
  p.ps = 1;
}

template <int AMODE, int BMODE, int OMODE>
inline int launch_simt_gemm(const GemmParams& p, int batches, cudaStream_t st) {
  dim3 grid(cdiv(p.M, GBM), cdiv(p.N, GBN), batches);
  simt_gemm_kernel<AMODE, BMODE, OMODE><<<grid, GTHREADS, 0, st>>>(p);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

}  // namespace vxb

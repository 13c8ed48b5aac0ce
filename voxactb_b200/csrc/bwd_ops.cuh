// Backward (adjoint) kernels of the Q-network's building blocks -- SURVEY.md section 8 row a18.
// The reference obtains these from torch autograd (`total_loss.backward()`, qattention_peract_bc_agent.py:581) over
// perceiver_lang_io.py / helpers/network_utils.py; every kernel here implements the closed-form adjoint that
// oracle/grad_oracle.py restates and checks against autograd.  Dense contractions (dgrad / wgrad of linears and
// convolutions, attention) go through the GEMM engine with transposed / transposed-im2col operand modes; everything
// else is an HBM-bound streaming kernel.
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include "simt_gemm.cuh"
#include "stream_ops.cuh"
#include "umma_host.cuh"

namespace vxb {
namespace bwd {

// Arithmetic of the backward contractions for the current vxb_qnet_backward_f32 call: VXB_MATH_F16X3 routes the large
// GEMMs / convolution dgrads to the tcgen05 split-fp16 engine (operands scaled per tensor, umma::gemm_any_f32), anything
// small -- and everything in VXB_MATH_FP32_SIMT -- runs as fp32 FFMA.
struct TensorCtx { int mm; void* scratch; size_t scratch_bytes; };
static thread_local TensorCtx g_tc = {VXB_MATH_FP32_SIMT, nullptr, 0};
inline bool use_tensor(int M, int N, int K) {
  return g_tc.mm == VXB_MATH_F16X3 && g_tc.scratch && M >= 128 && N >= 32 && K >= 64;
}
// returns VXB_OK when the tensor path ran, 1 when the caller should run the FFMA path, < 0 on error
inline int try_tensor_gemm(const float* A, long long lda, bool at, const float* W, long long ldw, bool wt, float* C, int ldc,
                           int M, int N, int K, bool accumulate, cudaStream_t st) {
  if (!use_tensor(M, N, K)) return 1;
  Arena local(g_tc.scratch, g_tc.scratch_bytes);
  const int rc = umma::gemm_any_f32(A, lda, at, W, ldw, wt, C, ldc, M, N, K, accumulate, local, st);
  return rc == VXB_E_WORKSPACE_TOO_SMALL ? 1 : rc;
}

// ------------------------------------------------------------------------------------------------ GEMM forms
// C[M,N] (+)= A[M,K] W[K,N]            (dgrad of a linear: dX = dY W)
inline int gemm_nn(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                   bool accumulate, cudaStream_t st) {
  const int rc = try_tensor_gemm(A, lda, false, W, ldw, true, C, ldc, M, N, K, accumulate, st);
  if (rc <= 0) return rc;
  GemmParams p;
  gemm_params_init(p);
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc;
  if (accumulate) { p.residual = C; p.res_rows = M; p.ldr = ldc; }
  return launch_simt_gemm<A_PLAIN, B_NN, O_PLAIN>(p, 1, st);
}
// C[M,N] (+)= A[M,K] W[N,K]^T
inline int gemm_nt(const float* A, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                   bool accumulate, cudaStream_t st) {
  const int rc = try_tensor_gemm(A, lda, false, W, ldw, false, C, ldc, M, N, K, accumulate, st);
  if (rc <= 0) return rc;
  GemmParams p;
  gemm_params_init(p);
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc;
  if (accumulate) { p.residual = C; p.res_rows = M; p.ldr = ldc; }
  return launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, 1, st);
}
inline int pick_ksplit(int M, int N, int K) {
  const long long tiles = (long long)cdiv(M, GBM) * cdiv(N, GBN);
  if (tiles >= 148 * 2 || K < 4096) return 1;
  long long want = (148 * 4 + tiles - 1) / tiles;
  want = std::min<long long>(want, K / 1024);
  return (int)std::max<long long>(1, want);
}
// C[m][n] (+)= sum_k At[k][m] W[k][n] for M = 64 and N <= NX <= 16 over a very tall K: the weight gradient of the 1 x 1
// input_preprocess convolution (64 x 10 over B V^3 = 16 M voxels).  A 128 x 64 GEMM tile computes 12 x more products than the
// 64 x 10 result needs (6 ms as a split-K FFMA GEMM); here a warp streams rows (256-byte gradient row: two channels per lane,
// the N inputs broadcast), 2 NX accumulators per lane, one shared-memory + one global atomic reduction per block: HBM-bound.
template <int NX>
static __global__ void __launch_bounds__(256)
wgrad_tall64_kernel(const float* __restrict__ At, int lda, const float* __restrict__ W, int ldw, int n_valid, long long K,
                    float* __restrict__ C, int ldc, const float* __restrict__ y, float slope, float* __restrict__ db) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc0[NX], acc1[NX], b0 = 0.f, b1 = 0.f;
#pragma unroll
  for (int n = 0; n < NX; ++n) acc0[n] = acc1[n] = 0.f;
  // optional fused LeakyReLU adjoint: the gradient row is taken through y > 0 ? g : slope * g on the way in (y = the layer's
  // output, same [K, lda] layout), so the in-place lrelu pass and the bias column sum over the 4 GB tensor are not needed
  auto grow = [&](long long k) -> float2 {
    float2 g = *reinterpret_cast<const float2*>(At + k * lda + lane * 2);
    if (y) {
      const float2 yv = *reinterpret_cast<const float2*>(y + k * lda + lane * 2);
      g.x = yv.x > 0.f ? g.x : g.x * slope;
      g.y = yv.y > 0.f ? g.y : g.y * slope;
    }
    return g;
  };
  const long long nw = (long long)gridDim.x * 8;
  long long k = (long long)blockIdx.x * 8 + warp;
  for (; k + nw < K; k += 2 * nw) {                 // two rows in flight
    const float2 ga = grow(k), gb = grow(k + nw);
    const float* xa = W + k * ldw;
    const float* xb = W + (k + nw) * ldw;
    b0 += ga.x + gb.x; b1 += ga.y + gb.y;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      const float va = n < n_valid ? __ldg(xa + n) : 0.f, vb = n < n_valid ? __ldg(xb + n) : 0.f;
      acc0[n] = fmaf(ga.x, va, acc0[n]); acc1[n] = fmaf(ga.y, va, acc1[n]);
      acc0[n] = fmaf(gb.x, vb, acc0[n]); acc1[n] = fmaf(gb.y, vb, acc1[n]);
    }
  }
  for (; k < K; k += nw) {
    const float2 ga = grow(k);
    const float* xa = W + k * ldw;
    b0 += ga.x; b1 += ga.y;
#pragma unroll
    for (int n = 0; n < NX; ++n) {
      const float va = n < n_valid ? __ldg(xa + n) : 0.f;
      acc0[n] = fmaf(ga.x, va, acc0[n]); acc1[n] = fmaf(ga.y, va, acc1[n]);
    }
  }
  __shared__ float red[64 * NX + 64];
  for (int i = threadIdx.x; i < 64 * NX + 64; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int n = 0; n < NX; ++n) {
    atomicAdd(&red[(lane * 2) * NX + n], acc0[n]);
    atomicAdd(&red[(lane * 2 + 1) * NX + n], acc1[n]);
  }
  atomicAdd(&red[64 * NX + lane * 2], b0);
  atomicAdd(&red[64 * NX + lane * 2 + 1], b1);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * NX; i += 256) {
    const int m = i / NX, n = i - m * NX;
    if (n < n_valid) atomicAdd(C + (size_t)m * ldc + n, red[i]);
  }
  if (db && threadIdx.x < 64) atomicAdd(db + threadIdx.x, red[64 * NX + threadIdx.x]);
}
inline bool wgrad_tall64_ok(const float* At, int lda, int M, int N, int K) {
  return M == 64 && N >= 1 && N <= 16 && K >= 8192 && (lda & 1) == 0 && (reinterpret_cast<uintptr_t>(At) & 7) == 0;
}
// y / slope: optional fused LeakyReLU adjoint of the gradient rows (y == nullptr: At is used as it is); db: optional bias gradient
inline int wgrad_tall64(const float* At, int lda, const float* W, int ldw, float* C, int ldc, int N, long long K, bool accumulate,
                        cudaStream_t st, const float* y = nullptr, float slope = 1.f, float* db = nullptr) {
  if (!accumulate) VXB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), 64, st));
  if (db) VXB_CUDA(cudaMemsetAsync(db, 0, 64 * sizeof(float), st));
  const int grid = 148 * 8;
  if (N <= 4) wgrad_tall64_kernel<4><<<grid, 256, 0, st>>>(At, lda, W, ldw, N, K, C, ldc, y, slope, db);
  else if (N <= 8) wgrad_tall64_kernel<8><<<grid, 256, 0, st>>>(At, lda, W, ldw, N, K, C, ldc, y, slope, db);
  else if (N <= 12) wgrad_tall64_kernel<12><<<grid, 256, 0, st>>>(At, lda, W, ldw, N, K, C, ldc, y, slope, db);
  else wgrad_tall64_kernel<16><<<grid, 256, 0, st>>>(At, lda, W, ldw, N, K, C, ldc, y, slope, db);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// C[M,N] (+)= At[K,M]^T W[K,N]          (wgrad of a linear: dW = dY^T X); C contiguous (ldc == N) when split-K is used
inline int gemm_tn(const float* At, int lda, const float* W, int ldw, float* C, int ldc, int M, int N, int K,
                   bool accumulate, cudaStream_t st) {
  const int rc = try_tensor_gemm(At, lda, true, W, ldw, true, C, ldc, M, N, K, accumulate, st);
  if (rc <= 0) return rc;
  if (wgrad_tall64_ok(At, lda, M, N, K)) return wgrad_tall64(At, lda, W, ldw, C, ldc, N, K, accumulate, st);
  GemmParams p;
  gemm_params_init(p);
  p.M = M; p.N = N; p.K = K;
  p.A = At; p.lda = lda; p.W = W; p.ldw = ldw; p.C = C; p.ldc = ldc;
  const int ks = (ldc == N) ? pick_ksplit(M, N, K) : 1;
  if (ks > 1) {
    if (!accumulate) VXB_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
    p.kchunk = cdiv(cdiv(K, ks), GBK) * GBK;
    return launch_simt_gemm<A_TRANS, B_NN, O_ATOMIC>(p, cdiv(K, p.kchunk), st);
  }
  if (accumulate) { p.residual = C; p.res_rows = M; p.ldr = ldc; }
  return launch_simt_gemm<A_TRANS, B_NN, O_PLAIN>(p, 1, st);
}

// ------------------------------------------------------------------------------------------------ small reductions
// out[n] (+)= sum_m g[m * ld + n]   (bias gradients); out zeroed here unless accumulate
static __global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ g, long long ld, long long M, int N, long long rows_per_block,
              float* __restrict__ out) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;
  const long long m0 = (long long)blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (n < N)
    for (long long m = m0 + rl; m < m1; m += 8) acc += g[m * ld + n];
  __shared__ float red[8][32];
  red[rl][threadIdx.x & 31] = acc;
  __syncthreads();
  if (rl == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + n, s);
  }
}
inline int colsum(const float* g, long long ld, long long M, int N, float* out, bool accumulate, cudaStream_t st) {
  if (!accumulate) VXB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), st));
  const int xb = cdiv(N, 32);
  long long yb = std::max<long long>(1, std::min<long long>((148 * 8 + xb - 1) / xb, (M + 63) / 64));
  const long long rpb = (M + yb - 1) / yb;
  yb = (M + rpb - 1) / rpb;
  colsum_kernel<<<dim3(xb, (unsigned)yb), 256, 0, st>>>(g, ld, M, N, rpb, out);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// out[i] (+)= sum_b x[b * stride + i], i < per   (gradients of batch-broadcast parameters: latents, pos_encoding)
static __global__ void __launch_bounds__(256)
batch_sum_kernel(const float* __restrict__ x, long long stride, int B, long long per, float* __restrict__ out, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    float s = accumulate ? out[i] : 0.f;
    for (int b = 0; b < B; ++b) s += x[b * stride + i];
    out[i] = s;
  }
}
inline int batch_sum(const float* x, long long stride, int B, long long per, float* out, bool accumulate, cudaStream_t st) {
  batch_sum_kernel<<<(int)std::min<long long>((per + 255) / 256, 148 * 8), 256, 0, st>>>(x, stride, B, per, out, accumulate ? 1 : 0);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------ elementwise adjoints
// LeakyReLU: g = y > 0 ? g : slope * g, in place (y > 0 <=> pre-activation > 0)
static __global__ void __launch_bounds__(256)
lrelu_bwd_kernel(float* __restrict__ g, const float* __restrict__ y, long long n4, float slope) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 gv = reinterpret_cast<float4*>(g)[i];
    const float4 yv = reinterpret_cast<const float4*>(y)[i];
    gv.x = yv.x > 0.f ? gv.x : gv.x * slope;
    gv.y = yv.y > 0.f ? gv.y : gv.y * slope;
    gv.z = yv.z > 0.f ? gv.z : gv.z * slope;
    gv.w = yv.w > 0.f ? gv.w : gv.w * slope;
    reinterpret_cast<float4*>(g)[i] = gv;
  }
}
inline int lrelu_bwd(float* g, const float* y, long long n, float slope, cudaStream_t st) {
  if (slope < 0.f) return VXB_OK;   // no activation
  if (n % 4) {
    set_error("lrelu_bwd: element count must be a multiple of 4");
    return VXB_E_BADARG;
  }
  lrelu_bwd_kernel<<<(int)std::min<long long>((n / 4 + 255) / 256, 148 * 16), 256, 0, st>>>(g, y, n / 4, slope);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// GEGLU (perceiver_lang_io.py:74-77): y = a * gelu_erf(g), h = [a | g] rows of 2n; gh = [gy * gelu(g) | gy * a * gelu'(g)]
static __global__ void __launch_bounds__(256)
geglu_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ h, float* __restrict__ gh, long long rows, int n) {
  const long long total = rows * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int c = (int)(i % n);
    const float a = h[r * 2 * n + c], g = h[r * 2 * n + n + c], go = gy[i];
    const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
    const float pdf = expf(-0.5f * g * g) * 0.3989422804014327f;
    gh[r * 2 * n + c] = go * g * cdf;
    gh[r * 2 * n + n + c] = go * a * (cdf + g * pdf);
  }
}
inline int geglu_bwd(const float* gy, const float* h, float* gh, long long rows, int n, cudaStream_t st) {
  geglu_bwd_kernel<<<148 * 16, 256, 0, st>>>(gy, h, gh, rows, n);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// LayerNorm (eps 1e-5, biased variance).  Warp per row.  gx (=|+=) per the closed form of grad_oracle.layernorm_backward;
// dw / db are accumulated with atomics (zeroed by the caller).  x and gx rows may be a strided slice per batch.
static __global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x, size_t x_batch_stride, int rows_per_batch,
                     const float* __restrict__ w, float* __restrict__ gx, int accumulate, float* __restrict__ dw,
                     float* __restrict__ db, long long rows, int n) {
  extern __shared__ float ln_smem[];   // [2][n] block partials of dw, db
  for (int i = threadIdx.x; i < 2 * n; i += 256) ln_smem[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * 8) {
    const size_t xo = (size_t)(row / rows_per_batch) * x_batch_stride + (size_t)(row % rows_per_batch) * n;
    const float* xr = x + xo;
    const float* gr = gy + (size_t)row * n;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += xr[i];
    s = warp_sum(s);
    const float mean = s / (float)n;
    float q = 0.f;
    for (int i = lane; i < n; i += 32) { const float d = xr[i] - mean; q += d * d; }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)n + 1e-5f);
    float c1 = 0.f, c2 = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float xh = (xr[i] - mean) * rstd, gh = gr[i] * w[i];
      c1 += gh; c2 += gh * xh;
    }
    c1 = warp_sum(c1) / (float)n;
    c2 = warp_sum(c2) / (float)n;
    float* gxr = gx + xo;
    for (int i = lane; i < n; i += 32) {
      const float xh = (xr[i] - mean) * rstd, g = gr[i];
      const float v = (g * w[i] - c1 - xh * c2) * rstd;
      gxr[i] = accumulate ? gxr[i] + v : v;
      atomicAdd(&ln_smem[i], g * xh);
      atomicAdd(&ln_smem[n + i], g);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += 256) {
    atomicAdd(dw + i, ln_smem[i]);
    atomicAdd(db + i, ln_smem[n + i]);
  }
}
inline int layernorm_bwd(const float* gy, const float* x, size_t x_batch_stride, int rows_per_batch, const float* w, float* gx,
                         bool accumulate, float* dw, float* db, long long rows, int n, cudaStream_t st) {
  VXB_CUDA(cudaMemsetAsync(dw, 0, (size_t)n * sizeof(float), st));
  VXB_CUDA(cudaMemsetAsync(db, 0, (size_t)n * sizeof(float), st));
  const int blocks = (int)std::min<long long>((rows + 7) / 8, 148 * 4);
  layernorm_bwd_kernel<<<blocks, 256, 2 * n * sizeof(float), st>>>(gy, x, x_batch_stride, rows_per_batch, w, gx,
                                                                    accumulate ? 1 : 0, dw, db, rows, n);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------ attention
// counter-based dropout mask (train-mode nn.Dropout on the attention probabilities, perceiver_lang_io.py:127-128): element
// `idx` of stream `seed` is kept when its hashed 32-bit value >= p * 2^32.  Recomputed (never stored) in the backward.
// (dropout_keep itself lives in common.cuh: the fused attention kernel of the training forward applies the same mask)
inline unsigned int dropout_threshold(float p) {
  const double t = (double)p * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : (unsigned int)t;
}
// out = dropout(P): rows of n valid columns, leading dimension ld (in place allowed)
static __global__ void __launch_bounds__(256)
dropout_rows_kernel(const float* __restrict__ P, float* __restrict__ out, long long rows, int n, int ld,
                    unsigned long long seed, unsigned int thresh, float inv_keep) {
  const long long total = rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    float v = 0.f;
    if (c < n) v = dropout_keep(seed, (unsigned long long)i, thresh) ? P[i] * inv_keep : 0.f;
    out[i] = v;
  }
}
// softmax backward in place on gA (gradient w.r.t. the dropped attention): gS = P * (gP - sum(gP * P)) * scale,
// gP = gA * mask / keep.  Block per row.
static __global__ void __launch_bounds__(256)
softmax_bwd_rows_kernel(const float* __restrict__ P, float* __restrict__ gA, int n, int ld, float scale,
                        unsigned long long seed, unsigned int thresh, float inv_keep) {
  const long long row = blockIdx.x;
  const float* pr = P + row * ld;
  float* gr = gA + row * ld;
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float d = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    float g = gr[i];
    if (thresh) g = dropout_keep(seed, (unsigned long long)(row * ld + i), thresh) ? g * inv_keep : 0.f;
    d += g * pr[i];
  }
  d = warp_sum(d);
  if (lane == 0) red[warp] = d;
  __syncthreads();
  d = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) d += red[i];
  for (int i = threadIdx.x; i < ld; i += 256) {
    float v = 0.f;
    if (i < n) {
      float g = gr[i];
      if (thresh) g = dropout_keep(seed, (unsigned long long)(row * ld + i), thresh) ? g * inv_keep : 0.f;
      v = pr[i] * (g - d) * scale;
    }
    gr[i] = v;
  }
}

struct AttnDropout { float p; unsigned long long seed; };

// Register-resident forms of the two row kernels above (a row of ld <= 1024 * NV floats lives in NV float4 per thread: one
// read and one write per element instead of three / two passes over global memory with scalar accesses):
//   softmax_rows_reg_kernel: P = softmax(x) in place and, with `dropped`, A = dropout(P) in the same pass (was a separate
//   pass over [B*H*Nq, Nk] with a 64-bit modulo per element: 10.5 ms of a training step at B=16);
//   softmax_bwd_rows_reg_kernel: gS = P * (gP - sum(gP * P)) * scale in place of gA, gP = gA * mask / keep.
__device__ __forceinline__ float block_reduce_256(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();                                  // `red` may still be read from the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}
template <int NV>
static __global__ void __launch_bounds__(256)
softmax_rows_reg_kernel(float* __restrict__ x, float* __restrict__ dropped, int n, int ld, unsigned long long seed,
                        unsigned int thresh, float inv_keep) {
  const size_t row = blockIdx.x;
  float4* xr = reinterpret_cast<float4*>(x + row * ld);
  const int nv = ld >> 2;
  __shared__ float red[8];
  float4 v[NV];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int q = threadIdx.x + k * 256;
    v[k] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (q < nv) {
      v[k] = xr[q];
      const int c = q * 4;
      if (c + 3 >= n) {                             // pad columns [n, ld) do not take part (exp -> 0)
        if (c >= n) v[k].x = -INFINITY;
        if (c + 1 >= n) v[k].y = -INFINITY;
        if (c + 2 >= n) v[k].z = -INFINITY;
        v[k].w = -INFINITY;
      }
      m = fmaxf(m, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
    }
  }
  m = block_reduce_256(m, red, true);
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k].x = expf(v[k].x - m); v[k].y = expf(v[k].y - m); v[k].z = expf(v[k].z - m); v[k].w = expf(v[k].w - m);
    s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  }
  s = block_reduce_256(s, red, false);
  const float inv = 1.f / s;
  float4* dr = dropped ? reinterpret_cast<float4*>(dropped + row * ld) : nullptr;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int q = threadIdx.x + k * 256;
    if (q < nv) {
      float4 p = make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
      xr[q] = p;
      if (dr) {
        const unsigned long long i0 = (unsigned long long)row * ld + (unsigned long long)(q * 4);
        p.x = dropout_keep(seed, i0, thresh) ? p.x * inv_keep : 0.f;
        p.y = dropout_keep(seed, i0 + 1, thresh) ? p.y * inv_keep : 0.f;
        p.z = dropout_keep(seed, i0 + 2, thresh) ? p.z * inv_keep : 0.f;
        p.w = dropout_keep(seed, i0 + 3, thresh) ? p.w * inv_keep : 0.f;
        dr[q] = p;
      }
    }
  }
}
template <int NV>
static __global__ void __launch_bounds__(256)
softmax_bwd_rows_reg_kernel(const float* __restrict__ P, float* __restrict__ gA, int n, int ld, float scale,
                            unsigned long long seed, unsigned int thresh, float inv_keep) {
  const size_t row = blockIdx.x;
  const float4* pr = reinterpret_cast<const float4*>(P + row * ld);
  float4* gr = reinterpret_cast<float4*>(gA + row * ld);
  const int nv = ld >> 2;
  __shared__ float red[8];
  float4 p[NV], g[NV];
  float d = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int q = threadIdx.x + k * 256;
    p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    g[k] = p[k];
    if (q < nv) {
      p[k] = pr[q];                                 // pad columns of P are zero (softmax_rows_*): they add nothing
      g[k] = gr[q];
      if (thresh) {
        const unsigned long long i0 = (unsigned long long)row * ld + (unsigned long long)(q * 4);
        g[k].x = dropout_keep(seed, i0, thresh) ? g[k].x * inv_keep : 0.f;
        g[k].y = dropout_keep(seed, i0 + 1, thresh) ? g[k].y * inv_keep : 0.f;
        g[k].z = dropout_keep(seed, i0 + 2, thresh) ? g[k].z * inv_keep : 0.f;
        g[k].w = dropout_keep(seed, i0 + 3, thresh) ? g[k].w * inv_keep : 0.f;
      }
      const int c = q * 4;
      if (c + 3 >= n) {                             // gA's pad columns hold whatever the GEMM left there
        if (c >= n) g[k].x = 0.f;
        if (c + 1 >= n) g[k].y = 0.f;
        if (c + 2 >= n) g[k].z = 0.f;
        g[k].w = 0.f;
      }
      d += (g[k].x * p[k].x + g[k].y * p[k].y) + (g[k].z * p[k].z + g[k].w * p[k].w);
    }
  }
  d = block_reduce_256(d, red, false);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int q = threadIdx.x + k * 256;
    if (q < nv)
      gr[q] = make_float4(p[k].x * (g[k].x - d) * scale, p[k].y * (g[k].y - d) * scale, p[k].z * (g[k].z - d) * scale,
                          p[k].w * (g[k].w - d) * scale);
  }
}

// forward with materialised probabilities (dispatch.cuh attention_materialized) + train-mode dropout on them
inline int dropout_rows(const float* P, float* out, long long rows, int n, int ld, const AttnDropout& dr, cudaStream_t st) {
  dropout_rows_kernel<<<148 * 16, 256, 0, st>>>(P, out, rows, n, ld, dr.seed, dropout_threshold(dr.p), 1.f / (1.f - dr.p));
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// P = softmax(P) in place (+ A = dropout(P) when dr.p > 0); rows of n valid columns, leading dimension ld (multiple of 4)
inline int softmax_rows_dropout(float* P, float* A, long long rows, int n, int ld, const AttnDropout& dr, cudaStream_t st) {
  const bool drop = dr.p > 0.f;
  const unsigned int th = drop ? dropout_threshold(dr.p) : 0u;
  const float ik = drop ? 1.f / (1.f - dr.p) : 1.f;
  const int nv = ld / 4;
  const bool vec = ld % 4 == 0 && (((uintptr_t)P | (uintptr_t)A) & 15) == 0;
  if (vec && nv <= 256 * 2) softmax_rows_reg_kernel<2><<<(unsigned)rows, 256, 0, st>>>(P, drop ? A : nullptr, n, ld, dr.seed, th, ik);
  else if (vec && nv <= 256 * 4) softmax_rows_reg_kernel<4><<<(unsigned)rows, 256, 0, st>>>(P, drop ? A : nullptr, n, ld, dr.seed, th, ik);
  else if (vec && nv <= 256 * 8) softmax_rows_reg_kernel<8><<<(unsigned)rows, 256, 0, st>>>(P, drop ? A : nullptr, n, ld, dr.seed, th, ik);
  else {
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(P, n, ld);
    VXB_LAUNCH_CHECK();
    if (drop) return dropout_rows(P, A, rows, n, ld, dr, st);
  }
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
inline int softmax_bwd_rows(const float* P, float* gA, long long rows, int n, int ld, float scale, const AttnDropout& dr,
                            cudaStream_t st) {
  const bool drop = dr.p > 0.f;
  const unsigned int th = drop ? dropout_threshold(dr.p) : 0u;
  const float ik = drop ? 1.f / (1.f - dr.p) : 1.f;
  const int nv = ld / 4;
  const bool vec = ld % 4 == 0 && (((uintptr_t)P | (uintptr_t)gA) & 15) == 0;
  if (vec && nv <= 256 * 2) softmax_bwd_rows_reg_kernel<2><<<(unsigned)rows, 256, 0, st>>>(P, gA, n, ld, scale, dr.seed, th, ik);
  else if (vec && nv <= 256 * 4) softmax_bwd_rows_reg_kernel<4><<<(unsigned)rows, 256, 0, st>>>(P, gA, n, ld, scale, dr.seed, th, ik);
  else if (vec && nv <= 256 * 8) softmax_bwd_rows_reg_kernel<8><<<(unsigned)rows, 256, 0, st>>>(P, gA, n, ld, scale, dr.seed, th, ik);
  else softmax_bwd_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(P, gA, n, ld, scale, dr.seed, th, ik);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// Adjoint of softmax(scale q k^T) v per (batch, head) (grad_oracle.attention_core_backward); probabilities recomputed.
// q [B or 1 (qbs = 0)][Nq][H*dh] (ldq), k / v rows [B][Nk][...] (ldkv, kvbs), go [B][Nq][H*dh] (ldo, obs).
// Outputs: gq [B][Nq][H*dh] (ldgq, gqbs: always per batch), gk / gv [B][Nk][...] (ldgkv, gkvbs).
// bufP / bufG / bufA: [B*H*Nq][pad4(Nk)] floats each (bufA only used with dropout).
inline int attention_bwd(const float* q, int ldq, long long qbs, const float* k, const float* v, int ldkv, long long kvbs,
                         const float* go, int ldo, long long obs, float* gq, int ldgq, long long gqbs, float* gk, float* gv,
                         int ldgkv, long long gkvbs, int B, int H, int Nq, int Nk, int dh, float scale, float* bufP,
                         float* bufG, float* bufA, const AttnDropout& dr, cudaStream_t st) {
  const int Nkp = (Nk + 3) / 4 * 4;
  const long long szb = (long long)H * Nq * Nkp, szh = (long long)Nq * Nkp;
  GemmParams p;
  // the two K = dh products (P logits, gA) on the tensor cores when the step runs in VXB_MATH_F16X3; 1 = run the FFMA form
  auto scores_tc = [&](const float* a, int lda, long long abs_, const float* w, float* out, float alpha, bool dyn) -> int {
    if (g_tc.mm != VXB_MATH_F16X3 || !g_tc.scratch) return 1;
    Arena local(g_tc.scratch, g_tc.scratch_bytes);
    const int rc = umma::attn_scores_f32(a, lda, abs_, w, ldkv, kvbs, out, Nkp, B, H, Nq, Nk, dh, alpha, dyn, local, st);
    return (rc == VXB_E_WORKSPACE_TOO_SMALL || rc == VXB_E_UNSUPPORTED_SHAPE) ? 1 : rc;
  };
  // (1) P = softmax(scale q k^T)
  int rc = scores_tc(q, ldq, qbs, k, bufP, scale, false);
  if (rc < 0) return rc;
  if (rc == 1) {
  gemm_params_init(p);
  p.M = Nq; p.N = Nk; p.K = dh;
  p.A = q; p.lda = ldq; p.a_stride_zb = qbs; p.a_stride_zh = dh;
  p.W = k; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = bufP; p.ldc = Nkp; p.c_stride_zb = szb; p.c_stride_zh = szh;
  p.Hz = H; p.alpha = scale;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, B * H, st)));
  }
  const bool drop = dr.p > 0.f;
  VXB_TRY(softmax_rows_dropout(bufP, bufA, (long long)B * H * Nq, Nk, Nkp, dr, st));   // also bufA = dropout(P)
  // (2) gA = go v^T
  rc = scores_tc(go, ldo, obs, v, bufG, 1.f, true);
  if (rc < 0) return rc;
  if (rc == 1) {
  gemm_params_init(p);
  p.M = Nq; p.N = Nk; p.K = dh;
  p.A = go; p.lda = ldo; p.a_stride_zb = obs; p.a_stride_zh = dh;
  p.W = v; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = bufG; p.ldc = Nkp; p.c_stride_zb = szb; p.c_stride_zh = szh;
  p.Hz = H;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, B * H, st)));
  }
  // (3) gv = A^T go, A = dropout(P)
  const float* A = drop ? bufA : bufP;
  gemm_params_init(p);
  p.M = Nk; p.N = dh; p.K = Nq;
  p.A = A; p.lda = Nkp; p.a_stride_zb = szb; p.a_stride_zh = szh;
  p.W = go; p.ldw = ldo; p.w_stride_zb = obs; p.w_stride_zh = dh;
  p.C = gv; p.ldc = ldgkv; p.c_stride_zb = gkvbs; p.c_stride_zh = dh;
  p.Hz = H;
  VXB_TRY((launch_simt_gemm<A_TRANS, B_NN, O_PLAIN>(p, B * H, st)));
  // (4) gS in place of gA
  VXB_TRY(softmax_bwd_rows(bufP, bufG, (long long)B * H * Nq, Nk, Nkp, scale, dr, st));
  // (5) gq = gS k
  gemm_params_init(p);
  p.M = Nq; p.N = dh; p.K = Nk;
  p.A = bufG; p.lda = Nkp; p.a_stride_zb = szb; p.a_stride_zh = szh;
  p.W = k; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = gq; p.ldc = ldgq; p.c_stride_zb = gqbs; p.c_stride_zh = dh;
  p.Hz = H;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NN, O_PLAIN>(p, B * H, st)));
  // (6) gk = gS^T q
  gemm_params_init(p);
  p.M = Nk; p.N = dh; p.K = Nq;
  p.A = bufG; p.lda = Nkp; p.a_stride_zb = szb; p.a_stride_zh = szh;
  p.W = q; p.ldw = ldq; p.w_stride_zb = qbs; p.w_stride_zh = dh;
  p.C = gk; p.ldc = ldgkv; p.c_stride_zb = gkvbs; p.c_stride_zh = dh;
  p.Hz = H;
  return launch_simt_gemm<A_TRANS, B_NN, O_PLAIN>(p, B * H, st);
}

// ------------------------------------------------------------------------------------------------ pooling heads
// first arg-max position of every (b, c) of channels-last x [B, P, C], given the maxima mx[b * mx_stride + c]
static __global__ void __launch_bounds__(256)
channel_argmax_kernel(const float* __restrict__ x, const float* __restrict__ mx, int mx_stride, int B, long long P, int C,
                      int* __restrict__ idx) {
  const long long total = (long long)B * P * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pos = (i / C) % P;
    const int b = (int)(i / (P * C));
    if (x[i] == mx[(size_t)b * mx_stride + c]) atomicMin(idx + b * C + c, (int)pos);
  }
}
// the same for C / 4 a divisor of 256 (every shipped call: C = 64): grid (chunks, B), a thread keeps its 4 channels for the
// whole walk, 128-bit loads, no division (the generic kernel pays two 64-bit divisions per element: 2.7 ms per 4 GB tensor)
static __global__ void __launch_bounds__(256)
channel_argmax_vec_kernel(const float* __restrict__ x, const float* __restrict__ mx, int mx_stride, long long P, int C,
                          int* __restrict__ idx) {
  const int G = C >> 2, b = blockIdx.y;
  const int cg = threadIdx.x % G;
  const float4 m = *reinterpret_cast<const float4*>(mx + (size_t)b * mx_stride + cg * 4);
  const float4* xb = reinterpret_cast<const float4*>(x + (size_t)b * P * C);
  const int ppb = 256 / G;                                       // positions per block and iteration
  int best[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
  for (long long pos = (long long)blockIdx.x * ppb + threadIdx.x / G; pos < P; pos += (long long)gridDim.x * ppb) {
    const float4 v = xb[pos * G + cg];
    if (v.x == m.x) best[0] = min(best[0], (int)pos);
    if (v.y == m.y) best[1] = min(best[1], (int)pos);
    if (v.z == m.z) best[2] = min(best[2], (int)pos);
    if (v.w == m.w) best[3] = min(best[3], (int)pos);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (best[j] != 0x7fffffff) atomicMin(idx + b * C + cg * 4 + j, best[j]);
}
inline int channel_argmax(const float* x, const float* mx, int mx_stride, int B, long long P, int C, int* idx, cudaStream_t st) {
  VXB_CUDA(cudaMemsetAsync(idx, 0x7f, (size_t)B * C * sizeof(int), st));
  if (C % 4 == 0 && 256 % (C / 4) == 0 && mx_stride % 4 == 0 && P < (1ll << 31) &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(mx)) & 15) == 0) {
    const int ppb = 256 / (C / 4);
    const int bx = (int)std::max<long long>(1, std::min<long long>((P + ppb - 1) / ppb, (148 * 16 + B - 1) / B));
    channel_argmax_vec_kernel<<<dim3(bx, B), 256, 0, st>>>(x, mx, mx_stride, P, C, idx);
    VXB_LAUNCH_CHECK();
    return VXB_OK;
  }
  channel_argmax_kernel<<<148 * 16, 256, 0, st>>>(x, mx, mx_stride, B, P, C, idx);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// SpatialSoftmax3D + AdaptiveMaxPool3d(1) adjoint (grad_oracle.spatial_softmax3d_backward / global_maxpool_backward):
// g[b,v,c] (=|+=) p[v] / T * sum_k ge[b,c,k] (pos[v,k] - e[b,c,k]) + [v == argmax] gm[b,c],
// p[v] = 2^(x k - m) / s with the forward's own (m, s) (stats [B][2][C]), e = the forward's soft-argmax output.
// grid (chunks, B): a thread owns 4 channels and walks the positions of its chunk (same layout as ss_partial_kernel).
static __global__ void __launch_bounds__(256)
ss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ e, int e_stride,
              const float* __restrict__ ge, int ge_stride, const float* __restrict__ gm, int gm_stride,
              const int* __restrict__ argidx, float* __restrict__ g, int accumulate, int Dd, int Hh, int Ww, int C, int chunk) {
  extern __shared__ float ssb_lut[];   // coordinate LUTs [Dd + Hh + Ww]
  const int b = blockIdx.y;
  const int P = Dd * Hh * Ww;
  for (int i = threadIdx.x; i < Dd + Hh + Ww; i += 256)
    ssb_lut[i] = i < Dd ? ss_lin_coord(i, Dd) : (i < Dd + Hh ? ss_lin_coord(i - Dd, Hh) : ss_lin_coord(i - Dd - Hh, Ww));
  __syncthreads();
  const float* lutD = ssb_lut;
  const float* lutH = ssb_lut + Dd;
  const float* lutW = ssb_lut + Dd + Hh;
  const int G = C >> 2, PL = 256 / G;
  const int gq = threadIdx.x % G, pl = threadIdx.x / G;
  if (pl >= PL) return;
  const int c = gq * 4;
  float m[4], is[4], ex[4], ey[4], ez[4], gx[4], gy[4], gz[4], gmx[4];
  int am[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[j] = stats[((size_t)b * 2) * C + c + j];
    is[j] = 1.f / stats[((size_t)b * 2 + 1) * C + c + j];
    const float* eb = e + (size_t)b * e_stride + (c + j) * 3;
    const float* gb = ge + (size_t)b * ge_stride + (c + j) * 3;
    ex[j] = eb[0]; ey[j] = eb[1]; ez[j] = eb[2];
    gx[j] = gb[0] * 100.f; gy[j] = gb[1] * 100.f; gz[j] = gb[2] * 100.f;     // 1 / T, T = 0.01
    gmx[j] = gm[(size_t)b * gm_stride + c + j];
    am[j] = argidx[b * C + c + j];
  }
  const int p_begin = blockIdx.x * chunk, p_end = min(P, p_begin + chunk);
  int p = p_begin + pl;
  int d = p / (Hh * Ww), h = (p / Ww) % Hh, w = p % Ww;
  const float4* xp = reinterpret_cast<const float4*>(x + ((size_t)b * P) * C) + gq;
  float4* gp = reinterpret_cast<float4*>(g + ((size_t)b * P) * C) + gq;
  for (; p < p_end; p += PL) {
    const float4 v = __ldg(xp + (size_t)p * G);
    // meshgrid('xy') quirk of network_utils.py:782-792: pos_x along H, pos_y along D, pos_z along W
    const float px = lutH[h], py = lutD[d], pz = lutW[w];
    const float xv[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float pr = exp2f(xv[j] * kSSLog2eOverT - m[j]) * is[j];
      o[j] = pr * (gx[j] * (px - ex[j]) + gy[j] * (py - ey[j]) + gz[j] * (pz - ez[j]));
      if (p == am[j]) o[j] += gmx[j];
    }
    float4 r = make_float4(o[0], o[1], o[2], o[3]);
    if (accumulate) { const float4 t = gp[(size_t)p * G]; r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
    gp[(size_t)p * G] = r;
    w += PL;
    while (w >= Ww) {
      w -= Ww;
      if (++h >= Hh) { h = 0; ++d; }
    }
  }
}
inline int ss_bwd(const float* x, const float* stats, const float* e, int e_stride, const float* ge, int ge_stride,
                  const float* gm, int gm_stride, const int* argidx, float* g, bool accumulate, int B, int Dd, int Hh, int Ww,
                  int C, cudaStream_t st) {
  if (C % 4 || C > 1024) {
    set_error("ss_bwd: C=%d must be a multiple of 4, <= 1024", C);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  const size_t P = (size_t)Dd * Hh * Ww;
  const int chunks = ss_num_chunks(P, B);
  const int chunk = (int)((P + chunks - 1) / chunks);
  ss_bwd_kernel<<<dim3(chunks, B), 256, (size_t)(Dd + Hh + Ww) * sizeof(float), st>>>(x, stats, e, e_stride, ge, ge_stride, gm, gm_stride,
                                                                                      argidx, g, accumulate ? 1 : 0, Dd, Hh, Ww, C, chunk);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------ convolutions
// Adjoint of replicate padding (grad_oracle.replicate_pad_backward) generalised to a padded-gradient grid of extent
// Pn per axis (Pn = V + 2 pad for stride-1 convolutions, Pn = S * s for the stride-s patchify windows):
//   g[b, v, gc0 + c] (=|+=) inv * sum_{p in [0, Pn)^3 : clamp(p - pad, 0, V-1) == v} gxp[b, p, c0 + c],  c < nch
// (inv = 1 / *scale when the padded gradient was produced from scaled tensor-core operands)
static __global__ void __launch_bounds__(256)
fold_pad_kernel(const float* __restrict__ gxp, int Cx, int c0, int Pn, int pad, float* __restrict__ g, int Cg, int gc0, int nch,
                int V, int B, int accumulate, const float* __restrict__ scale) {
  const float inv = scale ? 1.f / *scale : 1.f;
  const int cg = nch / 4;
  const long long total = (long long)B * V * V * V * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 4;
    long long r = i / cg;
    const int x = (int)(r % V); r /= V;
    const int y = (int)(r % V); r /= V;
    const int z = (int)(r % V);
    const int b = (int)(r / V);
    const int zlo = z == 0 ? 0 : z + pad, zhi = min(z == V - 1 ? Pn - 1 : z + pad, Pn - 1);
    const int ylo = y == 0 ? 0 : y + pad, yhi = min(y == V - 1 ? Pn - 1 : y + pad, Pn - 1);
    const int xlo = x == 0 ? 0 : x + pad, xhi = min(x == V - 1 ? Pn - 1 : x + pad, Pn - 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pz = zlo; pz <= zhi; ++pz)
      for (int py = ylo; py <= yhi; ++py)
        for (int px = xlo; px <= xhi; ++px) {
          const float4 v = *reinterpret_cast<const float4*>(gxp + ((((size_t)b * Pn + pz) * Pn + py) * Pn + px) * Cx + c0 + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    float4* o = reinterpret_cast<float4*>(g + (i / cg) * Cg + gc0 + c);
    if (accumulate) { const float4 t = *o; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
    *o = acc;
  }
}
inline int fold_pad(const float* gxp, int Cx, int c0, int Pn, int pad, float* g, int Cg, int V, int B, bool accumulate,
                    cudaStream_t st, int gc0 = 0, int nch = -1, const float* scale = nullptr) {
  fold_pad_kernel<<<148 * 16, 256, 0, st>>>(gxp, Cx, c0, Pn, pad, g, Cg, gc0, nch < 0 ? Cg : nch, V, B, accumulate ? 1 : 0, scale);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// dgrad of a stride-1 replicate-padded convolution followed by the padding adjoint: destination d (of n) receives the
// input channels [64 d, 64 d + 64) ... (tensor path) or channel slices of the interleaved padded gradient (FFMA path).
struct FoldDst { float* g; int ld; int c0; };   // gradient tensor [B, V^3, ld], first channel of the 64-channel slice
inline int conv_dgrad_padded(const float* gz, int Co, const float* wd, int Ci, float* gxp, int B, int V, int k, cudaStream_t st);
inline int conv_dgrad_fold(const float* gz, int Cz, const float* wd, int Cx, float* gxp, int B, int V, int k, const FoldDst* dst,
                           bool accumulate, cudaStream_t st) {
  const int pad = k / 2, Pn = V + 2 * pad;
  if (g_tc.mm == VXB_MATH_F16X3 && g_tc.scratch && Cz % 64 == 0 && Cx % 64 == 0) {
    Arena local(g_tc.scratch, g_tc.scratch_bytes);
    float* scale = local.get<float>(64);
    const int rc = umma::conv_dgrad_f32(gz, Cz, wd, Cx, gxp, B, V, k, scale, local, st);
    if (rc == VXB_OK) {
      const size_t block = (size_t)B * Pn * Pn * Pn * 64;
      for (int j = 0; j < Cx / 64; ++j)
        VXB_TRY(fold_pad(gxp + j * block, 64, 0, Pn, pad, dst[j].g, dst[j].ld, V, B, accumulate, st, dst[j].c0, 64, scale));
      return VXB_OK;
    }
    if (rc != VXB_E_WORKSPACE_TOO_SMALL) return rc;
  }
  VXB_TRY(conv_dgrad_padded(gz, Cz, wd, Cx, gxp, B, V, k, st));
  for (int j = 0; j < Cx / 64; ++j)
    VXB_TRY(fold_pad(gxp, Cx, 64 * j, Pn, pad, dst[j].g, dst[j].ld, V, B, accumulate, st, dst[j].c0, 64, nullptr));
  return VXB_OK;
}

// dgrad weights of a stride-1 convolution from the PyTorch layout w[Co][Ci][k^3]:
//   wd[ci][t'][co] = w[co][ci][k^3 - 1 - t']   (all three axes flipped: gxp[p] = sum_t w_t^T gz[p - t])
static __global__ void conv_dgrad_weight_kernel(const float* __restrict__ w, float* __restrict__ wd, int Co, int Ci, int k3) {
  const size_t total = (size_t)Co * Ci * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    const int t = (int)((i / Co) % k3);
    const int ci = (int)(i / ((size_t)Co * k3));
    wd[i] = w[((size_t)co * Ci + ci) * k3 + (k3 - 1 - t)];
  }
}
// the same for the folded up-convolution: wfold[r][co][nb][ci] -> wd[ci][nb'][r * Co + co] = wfold[r][co][26 - nb'][ci]
static __global__ void fold_dgrad_weight_kernel(const float* __restrict__ wf, float* __restrict__ wd, int R, int Co, int Ci) {
  const size_t total = (size_t)R * Co * 27 * Ci;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int rc = (int)(i % ((size_t)R * Co));
    const int nb = (int)((i / ((size_t)R * Co)) % 27);
    const int ci = (int)(i / ((size_t)R * Co * 27));
    const int r = rc / Co, co = rc % Co;
    wd[i] = wf[(((size_t)r * Co + co) * 27 + (26 - nb)) * Ci + ci];
  }
}
// folded up-convolution dgrad as GEMM + col2im: wt2[(nb, ci)][(r, co)] = wfold[r][co][nb][ci], so that
// T[q][(nb, ci)] = sum_(r,co) g_ph[q][(r,co)] wt2[(nb,ci)][(r,co)] is output q's contribution to low[clamp(q + nb - 1)][ci]
static __global__ void fold_gemm_weight_kernel(const float* __restrict__ wf, float* __restrict__ wt2, int R, int Co, int Ci) {
  const size_t total = (size_t)R * Co * 27 * Ci;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int rc = (int)(i % ((size_t)R * Co));
    const int ci = (int)((i / ((size_t)R * Co)) % Ci);
    const int nb = (int)(i / ((size_t)R * Co * Ci));
    wt2[i] = wf[((size_t)rc * 27 + nb) * Ci + ci];
  }
}
// g[b, v, c] = sum_{nb} sum_{q : clamp(q + nb - 1, 0, S-1) == v} T[(b, q)][nb][c]   (col2im with the replicate-padding
// adjoint folded in; same source enumeration as trans_bwd_kernel)
static __global__ void __launch_bounds__(256)
col2im3_fold_kernel(const float* __restrict__ T, float* __restrict__ g, int B, int S, int C) {
  const int cg = C / 4;
  const long long total = (long long)B * S * S * S * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 4;
    long long r = i / cg;
    const int x = (int)(r % S); r /= S;
    const int y = (int)(r % S); r /= S;
    const int z = (int)(r % S);
    const int b = (int)(r / S);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dz = 0; dz < 3; ++dz) {
      int zs[2], nz = 0;
      { const int a = z - dz + 1; if (a >= 0 && a < S) zs[nz++] = a; if (z == 0 && dz == 0) zs[nz++] = 0; if (z == S - 1 && dz == 2) zs[nz++] = S - 1; }
      for (int dy = 0; dy < 3; ++dy) {
        int ys[2], ny = 0;
        { const int a = y - dy + 1; if (a >= 0 && a < S) ys[ny++] = a; if (y == 0 && dy == 0) ys[ny++] = 0; if (y == S - 1 && dy == 2) ys[ny++] = S - 1; }
        for (int dx = 0; dx < 3; ++dx) {
          int xs[2], nx = 0;
          { const int a = x - dx + 1; if (a >= 0 && a < S) xs[nx++] = a; if (x == 0 && dx == 0) xs[nx++] = 0; if (x == S - 1 && dx == 2) xs[nx++] = S - 1; }
          const int nb = (dz * 3 + dy) * 3 + dx;
          for (int a = 0; a < nz; ++a)
            for (int bb = 0; bb < ny; ++bb)
              for (int cc = 0; cc < nx; ++cc) {
                const size_t q = (((size_t)b * S + zs[a]) * S + ys[bb]) * S + xs[cc];
                const float4 v = *reinterpret_cast<const float4*>(T + (q * 27 + nb) * C + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
              }
        }
      }
    }
    *reinterpret_cast<float4*>(g + (i / cg) * C + c) = acc;
  }
}

// patchify (stride == k) dgrad weights: wd[t][ci][co] = w[co][ci][t]
static __global__ void patch_dgrad_weight_kernel(const float* __restrict__ w, float* __restrict__ wd, int Co, int Ci, int k3) {
  const size_t total = (size_t)Co * Ci * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Co);
    const int ci = (int)((i / Co) % Ci);
    const int t = (int)(i / ((size_t)Co * Ci));
    wd[i] = w[((size_t)co * Ci + ci) * k3 + t];
  }
}
// wgrad GEMM output [k^3][Ci][Co] -> PyTorch layout dw[Co][Ci][k^3]
static __global__ void wgrad_to_torch_kernel(const float* __restrict__ t, float* __restrict__ dw, int Co, int Ci, int k3) {
  const size_t total = (size_t)Co * Ci * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % k3);
    const int ci = (int)((i / k3) % Ci);
    const int co = (int)(i / ((size_t)k3 * Ci));
    dw[i] = t[((size_t)tap * Ci + ci) * Co + co];
  }
}

// gxp[b, p, 0:Ci] = sum_t w_t^T gz[b, p - t] on the padded-gradient grid (V + 2 pad)^3, gz zero outside [0, V)^3:
// implicit GEMM with the zero-out-of-bounds gather.  wd from conv_dgrad_weight_kernel ([Ci][k^3][Co]).
inline int conv_dgrad_padded(const float* gz, int Co, const float* wd, int Ci, float* gxp, int B, int V, int k, cudaStream_t st) {
  if (Co % GBK || Co % 4) {
    set_error("conv_dgrad: Co=%d must be a multiple of %d", Co, GBK);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  const int pad = k / 2, Pn = V + 2 * pad;
  GemmParams p;
  gemm_params_init(p);
  p.M = B * Pn * Pn * Pn; p.N = Ci; p.K = k * k * k * Co;
  p.src0 = gz; p.src1 = nullptr; p.C0 = Co; p.C1 = 0;
  p.Di = V; p.Do = Pn; p.kk = k; p.cstride = 1; p.pad = 2 * pad; p.zero_oob = 1;
  p.W = wd; p.ldw = p.K;
  p.C = gxp; p.ldc = Ci;
  return launch_simt_gemm<A_CONV, B_NT, O_PLAIN>(p, 1, st);
}

// dwt[(tap, ci)][co] = sum_rows x[clamp(o * stride - pad + tap)][ci] gz[row][co]   (transposed-im2col GEMM, split-K)
// x = concat(src0[C0], src1[C1]) channels-last on a Di^3 grid, gz [B, Do^3, Co]; dwt zeroed here.
inline int conv_wgrad(const float* src0, const float* src1, int C0, int C1, const float* gz, int Co, float* dwt, int B, int Di,
                      int Do, int k, int stride, cudaStream_t st) {
  const int Cin = C0 + C1;
  if (Cin % 4 || C0 % 4) {
    set_error("conv_wgrad: channel counts must be multiples of 4");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  GemmParams p;
  gemm_params_init(p);
  p.M = k * k * k * Cin; p.N = Co; p.K = B * Do * Do * Do;
  p.src0 = src0; p.src1 = src1; p.C0 = C0; p.C1 = C1;
  p.Di = Di; p.Do = Do; p.kk = k; p.cstride = stride; p.pad = k / 2;
  p.W = gz; p.ldw = Co;
  p.C = dwt; p.ldc = Co;
  VXB_CUDA(cudaMemsetAsync(dwt, 0, (size_t)p.M * Co * sizeof(float), st));
  const int ks = pick_ksplit(p.M, p.N, p.K);
  p.kchunk = cdiv(cdiv(p.K, ks), GBK) * GBK;
  return launch_simt_gemm<A_CONV_T, B_NN, O_ATOMIC>(p, cdiv(p.K, p.kchunk), st);
}

// fine grid [B, (S s)^3, C] -> coarse-major phases [B, S^3, s^3 * C]: out[b, q, r * C + c] = in[b, s q + r, c],
// r = (rd s + rh) s + rw (the phase order of fold_upconv_weights_kernel)
static __global__ void __launch_bounds__(256)
phase_gather_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int S, int s, int C) {
  const int cg = C / 4, R = s * s * s, V = S * s;
  const long long total = (long long)B * S * S * S * R * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 4;
    long long t = i / cg;
    const int r = (int)(t % R); t /= R;
    const int qw = (int)(t % S); t /= S;
    const int qh = (int)(t % S); t /= S;
    const int qd = (int)(t % S);
    const int b = (int)(t / S);
    const int rd = r / (s * s), rh = (r / s) % s, rw = r % s;
    const size_t src = ((((size_t)b * V + qd * s + rd) * V + qh * s + rh) * V + qw * s + rw) * C + c;
    *reinterpret_cast<float4*>(out + (i / cg) * C + c) = *reinterpret_cast<const float4*>(in + src);
  }
}

// transpose of fold_upconv_weights_kernel: dwt[(nb, ci)][(r, co)] (the wgrad GEMM output of the folded form)
// -> dw[co][ci][k^3] of the original k x k x k convolution applied after the trilinear x s upsampling
static __global__ void fold_upconv_weights_bwd_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int Co, int Ci,
                                                      int k, int s) {
  const int pad = k / 2, k3 = k * k * k, R = s * s * s;
  const size_t total = (size_t)Co * Ci * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % k3);
    const int ci = (int)((i / k3) % Ci);
    const int co = (int)(i / ((size_t)k3 * Ci));
    const int td = t / (k * k), th = (t / k) % k, tw = t % k;
    float acc = 0.f;
    for (int rd = 0; rd < s; ++rd)
      for (int nd = -1; nd <= 1; ++nd) {
        const float ud = upsample_tap_weight(rd + td - pad, nd, s);
        if (ud == 0.f) continue;
        for (int rh = 0; rh < s; ++rh)
          for (int nh = -1; nh <= 1; ++nh) {
            const float uh = upsample_tap_weight(rh + th - pad, nh, s);
            if (uh == 0.f) continue;
            for (int rw = 0; rw < s; ++rw)
              for (int nw = -1; nw <= 1; ++nw) {
                const float uw = upsample_tap_weight(rw + tw - pad, nw, s);
                if (uw == 0.f) continue;
                const int nb = ((nd + 1) * 3 + (nh + 1)) * 3 + (nw + 1);
                const int r = (rd * s + rh) * s + rw;
                acc += dwt[((size_t)nb * Ci + ci) * ((size_t)R * Co) + (size_t)r * Co + co] * (ud * uh * uw);
              }
          }
      }
    dw[i] = acc;
  }
}

// trans_decoder (Conv3d 64 -> 1, k3, replicate pad, no activation) adjoint, gather form:
//   G[v][t] = sum_{o : clamp(o + t - 1) == v} g[o];   gu[v][c] (+)= sum_t w[t][c] G[v][t];   dw[t][c] += u[v][c] G[v][t]
// 16 threads per voxel (4 channels each); dw accumulated per block in shared memory, then atomically (dw zeroed by caller).
template <int C>
static __global__ void __launch_bounds__(128, 3)       // <= 168 registers (27 x 4 weight-gradient accumulators per thread): 12 warps / SM
trans_bwd_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ wt /*[27][C]*/,
                 float* __restrict__ gu, int accumulate, float* __restrict__ dwt /*[27][C]*/, int B, int V) {
  constexpr int G4 = C / 4;
  static_assert(G4 == 16, "one half-warp per voxel");
  __shared__ float sdw[27 * C];
  __shared__ __align__(16) float swt[27 * C];     // the stencil weights: 27 LDS.128 per voxel instead of 27 L1 requests
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) { sdw[i] = 0.f; swt[i] = wt[i]; }
  __syncthreads();
  const long long total = (long long)B * V * V * V * G4;
  float acc[27][4];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
  const int hl = threadIdx.x & 15;              // lane within the half-warp = channel group
  const int c = hl * 4;
  // warp-uniform trip count: every lane of a warp runs the same iterations (shuffles below), tails are predicated
  const long long wbase = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long i0 = wbase; i0 < total; i0 += (long long)gridDim.x * blockDim.x) {
    const long long i = i0 + (threadIdx.x & 31);
    const bool valid = i < total;
    const long long vox = (valid ? i : 0) / G4;
    unsigned int r = (unsigned int)vox;            // B * V^3 < 2^32 (launcher): 32-bit index arithmetic
    const int x = (int)(r % (unsigned int)V); r /= (unsigned int)V;
    const int y = (int)(r % (unsigned int)V); r /= (unsigned int)V;
    const int z = (int)(r % (unsigned int)V);
    const int b = (int)(r / (unsigned int)V);
    const float* gb = g + (size_t)b * V * V * V;
    // interior voxels (94 % at V = 100) have exactly one source per tap: o = v - t + 1
    const bool interior = x >= 1 && x <= V - 2 && y >= 1 && y <= V - 2 && z >= 1 && z <= V - 2;
    // the 27 gathered sums G[t] of this voxel: lane hl forms taps hl and hl + 16, the half-warp shares them by shuffle
    float gpart[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int t = hl + 16 * k;
      if (t < 27 && valid && interior) {
        const int dz = t / 9, dy = (t / 3) % 3, dx = t % 3;
        gpart[k] = __ldg(gb + ((size_t)(z - dz + 1) * V + (y - dy + 1)) * V + (x - dx + 1));
      } else if (t < 27 && valid) {
        const int dz = t / 9, dy = (t / 3) % 3, dx = t % 3;
        int zs[2], ys[2], xs[2], nz = 0, ny = 0, nx = 0;
        { const int a = z - dz + 1; if (a >= 0 && a < V) zs[nz++] = a; if (z == 0 && dz == 0) zs[nz++] = 0; if (z == V - 1 && dz == 2) zs[nz++] = V - 1; }
        { const int a = y - dy + 1; if (a >= 0 && a < V) ys[ny++] = a; if (y == 0 && dy == 0) ys[ny++] = 0; if (y == V - 1 && dy == 2) ys[ny++] = V - 1; }
        { const int a = x - dx + 1; if (a >= 0 && a < V) xs[nx++] = a; if (x == 0 && dx == 0) xs[nx++] = 0; if (x == V - 1 && dx == 2) xs[nx++] = V - 1; }
        float G = 0.f;
        for (int a = 0; a < nz; ++a)
          for (int bb = 0; bb < ny; ++bb)
            for (int cc = 0; cc < nx; ++cc) G += __ldg(gb + ((size_t)zs[a] * V + ys[bb]) * V + xs[cc]);
        gpart[k] = G;
      }
    }
    float4 uv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) uv = *reinterpret_cast<const float4*>(u + vox * C + c);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float G = __shfl_sync(0xffffffffu, gpart[t >> 4], t & 15, 16);
      const float4 wv = *reinterpret_cast<const float4*>(swt + t * C + c);
      o.x = fmaf(wv.x, G, o.x); o.y = fmaf(wv.y, G, o.y); o.z = fmaf(wv.z, G, o.z); o.w = fmaf(wv.w, G, o.w);
      acc[t][0] = fmaf(uv.x, G, acc[t][0]); acc[t][1] = fmaf(uv.y, G, acc[t][1]);
      acc[t][2] = fmaf(uv.z, G, acc[t][2]); acc[t][3] = fmaf(uv.w, G, acc[t][3]);
    }
    if (valid) {
      float4* dst = reinterpret_cast<float4*>(gu + vox * C + c);
      if (accumulate) { const float4 t4 = *dst; o.x += t4.x; o.y += t4.y; o.z += t4.z; o.w += t4.w; }
      *dst = o;
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&sdw[t * C + c + j], acc[t][j]);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) atomicAdd(dwt + i, sdw[i]);
}

// replicate-padded 3x3x3 im2col of a channels-last grid: out[(b, q)][(nb, c)] = x[b, clamp(q + nb - 1), c]
static __global__ void __launch_bounds__(256)
im2col3_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int S, int C) {
  const int cg = C / 4;
  const long long total = (long long)B * S * S * S * 27 * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 4;
    long long r = i / cg;
    const int nb = (int)(r % 27); r /= 27;
    const int qw = (int)(r % S); r /= S;
    const int qh = (int)(r % S); r /= S;
    const int qd = (int)(r % S);
    const int b = (int)(r / S);
    const int d = min(max(qd + nb / 9 - 1, 0), S - 1), h = min(max(qh + (nb / 3) % 3 - 1, 0), S - 1),
              w = min(max(qw + nb % 3 - 1, 0), S - 1);
    *reinterpret_cast<float4*>(out + (i / cg) * C + c) =
        *reinterpret_cast<const float4*>(x + ((((size_t)b * S + d) * S + h) * S + w) * C + c);
  }
}

// token disassembly (adjoint of assemble_tokens_kernel): g_ins [B, nl+T, C] ->
//   g_lang [B, nl, C], g_patch [B, T, im], g_p1 [B, im] (+ g_p2) = sum over the T voxel tokens of the proprio channels
static __global__ void __launch_bounds__(256)
disassemble_tokens_kernel(const float* __restrict__ gins, float* __restrict__ glang, float* __restrict__ gpatch, int B, int nl,
                          int T, int C, int im) {
  const size_t total = (size_t)B * (nl + T) * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t row = i / C;
    const int j = (int)(row % (nl + T));
    const int b = (int)(row / (nl + T));
    if (j < nl) glang[((size_t)b * nl + j) * C + c] = gins[i];
    else if (c < im) gpatch[((size_t)b * T + (j - nl)) * im + c] = gins[i];
  }
}
// g_p[b, c] = sum_t gins[b, nl + t, c_off + c]
static __global__ void __launch_bounds__(256)
proprio_token_sum_kernel(const float* __restrict__ gins, float* __restrict__ gp, int nl, int T, int C, int c_off, int im) {
  const int b = blockIdx.x;
  const int c = threadIdx.x & 63, tl = threadIdx.x >> 6;   // im == 64: 4 token lanes
  float acc = 0.f;
  if (c < im)
    for (int t = tl; t < T; t += 4) acc += gins[((size_t)b * (nl + T) + nl + t) * C + c_off + c];
  __shared__ float red[4][64];
  red[tl][c] = acc;
  __syncthreads();
  if (tl == 0 && c < im) gp[(size_t)b * im + c] = red[0][c] + red[1][c] + red[2][c] + red[3][c];
}

}  // namespace bwd
}  // namespace vxb

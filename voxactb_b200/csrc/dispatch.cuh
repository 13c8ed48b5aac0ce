// Contraction dispatch: fp32 FFMA (simt_gemm.cuh) or tcgen05 split-bf16 (umma_gemm.cuh).
#pragma once
#include "common.cuh"
#include "simt_gemm.cuh"
#include "ops.cuh"
#include "umma_host.cuh"

namespace vxb {

// C[M,N] = act(alpha * A[M,K] W[N,K]^T + bias) + residual[(m % res_rows)]
inline int linear(const float* A, int lda, const float* W, int ldw, const float* bias,
                  const float* residual, int res_rows, int ldr, float* C, int ldc, int M, int N,
                  int K, float alpha, float act_slope, int math_mode, cudaStream_t st,
                  Arena* scratch = nullptr, const umma::Planes* Wpre = nullptr) {
  if (math_mode == VXB_MATH_F16X3 && scratch && M >= 128 && N >= 32 && K >= 32) {
    Arena local(scratch->base, scratch->cap);   // every op bumps from the start of the scratch arena
    return umma::linear_f32(A, lda, W, ldw, Wpre, bias, residual, res_rows, ldr, C, ldc, M, N, K, alpha,
                            act_slope, local, st);
  }
  GemmParams p;
  gemm_params_init(p);
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda;
  p.W = W; p.ldw = ldw;
  p.bias = bias;
  p.residual = residual; p.res_rows = res_rows > 0 ? res_rows : 1; p.ldr = ldr;
  p.alpha = alpha; p.act_slope = act_slope;
  p.C = C; p.ldc = ldc;
  return launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, 1, st);
}

// channels-last conv3d with replicate padding k/2; input = concat(src0[C0], src1[C1]) on channels;
// wt tap-major [Co][k^3][C0+C1]; out [B,Do^3,Co]
inline int conv3d(const float* src0, const float* src1, int C0, int C1, const float* wt,
                  const float* bias, float* out, int B, int Di, int Do, int Co, int k, int stride,
                  float act_slope, int math_mode, cudaStream_t st, Arena* scratch = nullptr,
                  const umma::Planes* Wpre = nullptr) {
  if (math_mode == VXB_MATH_F16X3 && scratch && stride == 1 && Di == Do && C0 % 64 == 0 && C1 % 64 == 0 &&
      Co == 64) {
    Arena local(scratch->base, scratch->cap);
    umma::Planes wp;
    const long long Kt = (long long)k * k * k * (C0 + C1);
    if (Wpre) {
      wp = *Wpre;
    } else {
      wp.hi = local.get<__nv_bfloat16>((size_t)Co * Kt);
      wp.lo = local.get<__nv_bfloat16>((size_t)Co * Kt);
      wp.ld = Kt;
      if (!local.ok) {
        set_error("conv3d: scratch too small");
        return VXB_E_WORKSPACE_TOO_SMALL;
      }
      VXB_TRY(umma::split_rows(wt, Kt, Co, (int)Kt, wp, st));
    }
    return umma::conv3d_f32(src0, src1, C0, C1, wp, bias, out, B, Di, Co, k, act_slope, local, st);
  }
  const int Cin = C0 + C1;
  if (Cin % GBK != 0 || C0 % GBK != 0) {
    set_error("conv3d: channel counts must be multiples of %d (C0=%d, C1=%d)", GBK, C0, C1);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  GemmParams p;
  gemm_params_init(p);
  p.M = B * Do * Do * Do; p.N = Co; p.K = k * k * k * Cin;
  p.src0 = src0; p.src1 = src1; p.C0 = C0; p.C1 = C1;
  p.Di = Di; p.Do = Do; p.kk = k; p.cstride = stride; p.pad = k / 2;
  p.W = wt; p.ldw = p.K;
  p.bias = bias; p.act_slope = act_slope;
  p.C = out; p.ldc = Co;
  return launch_simt_gemm<A_CONV, B_NT, O_PLAIN>(p, 1, st);
}

// polyphase form of conv_k o upsample_s: s^3 independent 3x3x3 convolutions on the S^3 grid,
// phase r written at fine voxel s*q + r.  wfold [s^3][Co][27][Ci]
inline int upconv3d_folded(const float* low, const float* wfold, const float* bias, float* out,
                           int B, int S, int Ci, int Co, int s, float act_slope, int math_mode,
                           cudaStream_t st, Arena* scratch = nullptr, const umma::Planes* Wpre = nullptr,
                           const umma::Planes* out_planes = nullptr, const float* f8a = nullptr,
                           const umma::F8cGemm* f8g = nullptr, const umma::UpconvSparsity* sp = nullptr) {
  if (math_mode == VXB_MATH_F16X3 && scratch && Ci % 64 == 0 && Co == 64) {
    Arena local(scratch->base, scratch->cap);
    umma::Planes wp;
    const long long Kt = 27ll * Ci, Nt = (long long)s * s * s * Co;
    if (Wpre) {
      wp = *Wpre;
    } else {
      wp.hi = local.get<__nv_bfloat16>((size_t)Nt * Kt);
      wp.lo = local.get<__nv_bfloat16>((size_t)Nt * Kt);
      wp.ld = Kt;
      if (!local.ok) {
        set_error("upconv3d: scratch too small");
        return VXB_E_WORKSPACE_TOO_SMALL;
      }
      VXB_TRY(umma::split_rows(wfold, Kt, Nt, (int)Kt, wp, st));
    }
    return umma::upconv_f32(low, wp, bias, out, B, S, Ci, Co, s, act_slope, local, st, out_planes, f8a, f8g, sp);
  }
  GemmParams p;
  gemm_params_init(p);
  p.M = B * S * S * S; p.N = Co; p.K = 27 * Ci;
  p.src0 = low; p.src1 = nullptr; p.C0 = Ci; p.C1 = 0;
  p.Di = S; p.Do = S; p.kk = 3; p.cstride = 1; p.pad = 1;
  p.W = wfold; p.ldw = p.K; p.w_stride_zb = (long long)Co * p.K;
  p.bias = bias; p.act_slope = act_slope;
  p.C = out; p.ldc = Co;
  p.ps = s;
  return launch_simt_gemm<A_CONV, B_NT, O_PHASE>(p, s * s * s, st);
}

// softmax(scale * q k^T) v per (batch, head) with materialised scores
inline int attention_materialized(const float* q, int ldq, long long qbs, const float* k, const float* v,
                     int ldkv, long long kvbs, float* out, int ldo, long long obs, int B, int H,
                     int Nq, int Nk, int dh, float scale, float* sim, int math_mode,
                     cudaStream_t st) {
  const int Nkp = (Nk + 3) / 4 * 4;
  GemmParams p;
  gemm_params_init(p);
  p.M = Nq; p.N = Nk; p.K = dh;
  p.A = q; p.lda = ldq; p.a_stride_zb = qbs; p.a_stride_zh = dh;
  p.W = k; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = sim; p.ldc = Nkp; p.c_stride_zb = (long long)H * Nq * Nkp; p.c_stride_zh = (long long)Nq * Nkp;
  p.Hz = H; p.alpha = scale;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, B * H, st)));
  softmax_rows_kernel<<<(unsigned)((size_t)B * H * Nq), 256, 0, st>>>(sim, Nk, Nkp);
  VXB_LAUNCH_CHECK();
  gemm_params_init(p);
  p.M = Nq; p.N = dh; p.K = Nk;
  p.A = sim; p.lda = Nkp; p.a_stride_zb = (long long)H * Nq * Nkp; p.a_stride_zh = (long long)Nq * Nkp;
  p.W = v; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = out; p.ldc = ldo; p.c_stride_zb = obs; p.c_stride_zh = dh;
  p.Hz = H;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NN, O_PLAIN>(p, B * H, st)));
  (void)math_mode;
  return VXB_OK;
}


}  // namespace vxb

// C-ABI wrappers of the individual building blocks (per-op parity tests) and action selection.
#include "common.cuh"
#include "ops.cuh"
#include "simt_gemm.cuh"
#include "stream_ops.cuh"
#include "dispatch.cuh"

using namespace vxb;

namespace vxb {

// ---------------------------------------------------------------- action selection
// argmax over V^3 per sample: (value, index) pairs, ties -> lowest index (torch.argmax on CPU
// returns the first maximal element).  Pass 1: per-chunk, pass 2: merge + heads.
// torch.argmax order: NaN counts as the maximum, ties (and NaN ties) go to the lowest index
__device__ __forceinline__ bool argmax_better(float v, int i, float best, int bi) {
  if (v != v) return best == best || i < bi;
  if (best != best) return false;
  return v > best || (v == best && i < bi);
}

__global__ void __launch_bounds__(256)
argmax_partial_kernel(const float* __restrict__ q, size_t n, int chunks, float* __restrict__ pv,
                      int* __restrict__ pi) {
  const int b = blockIdx.y, ck = blockIdx.x;
  const size_t per = (n + chunks - 1) / chunks;
  const size_t beg = ck * per, end = min(n, beg + per);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (size_t i = beg + threadIdx.x; i < end; i += 256) {
    const float v = q[(size_t)b * n + i];
    if (argmax_better(v, (int)i, best, bi)) { best = v; bi = (int)i; }
  }
  __shared__ float sv[256];
  __shared__ int si[256];
  sv[threadIdx.x] = best; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float v2 = sv[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      if (argmax_better(v2, i2, sv[threadIdx.x], si[threadIdx.x])) {
        sv[threadIdx.x] = v2; si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { pv[b * chunks + ck] = sv[0]; pi[b * chunks + ck] = si[0]; }
}

__device__ __forceinline__ int small_argmax(const float* v, int n) {
  int bi = 0;
  float best = v[0];
  for (int i = 1; i < n; ++i)
    if (argmax_better(v[i], i, best, bi)) { best = v[i]; bi = i; }
  return bi;
}

__global__ void select_action_final_kernel(const float* __restrict__ pv, const int* __restrict__ pi,
                                           int chunks, const float* __restrict__ rot_grip,
                                           const float* __restrict__ collision,
                                           const float* __restrict__ bounds, int Bb, int B, int V,
                                           int R, int32_t* __restrict__ coords,
                                           int32_t* __restrict__ rg_idx, int32_t* __restrict__ coll_idx,
                                           float* __restrict__ att_xyz) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = 0; k < chunks; ++k) {
    const float v = pv[b * chunks + k];
    const int i = pi[b * chunks + k];
    if (argmax_better(v, i, best, bi)) { best = v; bi = i; }
  }
  // _argmax_3d (qattention_peract_bc_agent.py:57-63): ((idx // h) // d, (idx // h) % w, idx % w)
  const int c0 = (bi / V) / V, c1 = (bi / V) % V, c2 = bi % V;
  coords[b * 3 + 0] = c0; coords[b * 3 + 1] = c1; coords[b * 3 + 2] = c2;
  if (rot_grip && rg_idx) {
    const float* rg = rot_grip + (size_t)b * (3 * R + 2);
    rg_idx[b * 4 + 0] = small_argmax(rg, R);
    rg_idx[b * 4 + 1] = small_argmax(rg + R, R);
    rg_idx[b * 4 + 2] = small_argmax(rg + 2 * R, R);
    rg_idx[b * 4 + 3] = small_argmax(rg + 3 * R, 2);
  }
  if (collision && coll_idx) coll_idx[b] = small_argmax(collision + (size_t)b * 2, 2);
  if (att_xyz && bounds) {
    const float* bd = bounds + (Bb == 1 ? 0 : b) * 6;
    const int cc[3] = {c0, c1, c2};
    for (int a = 0; a < 3; ++a) {
      // res = (bounds[:,3:] - bounds[:,:3]) / voxel_size ; coord = bounds[:, :3] + res*idx + res/2  (agent:701,724)
      const float res = __fdiv_rn(__fsub_rn(bd[3 + a], bd[a]), (float)V);
      att_xyz[b * 3 + a] = __fadd_rn(__fadd_rn(bd[a], __fmul_rn(res, (float)cc[a])), __fdiv_rn(res, 2.f));
    }
  }
}

}  // namespace vxb

static const int kArgmaxChunks = 64;

extern "C" size_t vxb_select_action_workspace_bytes(int B, int V) {
  (void)V;
  return align_up((size_t)B * kArgmaxChunks * 4, 256) * 2;
}

extern "C" int vxb_select_action_f32(const float* q_trans, const float* rot_grip,
                                     const float* collision, const float* bounds, int Bb, int B,
                                     int V, int R, int32_t* coords, int32_t* rot_grip_idx,
                                     int32_t* coll_idx, float* attention_xyz, void* ws,
                                     size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(q_trans && coords && ws, "select_action: null pointer");
  VXB_CHECK_ARG(B > 0 && V > 0, "select_action: bad sizes");
  VXB_CHECK_ARG(!bounds || Bb == 1 || Bb == B, "select_action: bounds batch must be 1 or B");
  if (ws_bytes < vxb_select_action_workspace_bytes(B, V)) {
    set_error("select_action: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* pv = (float*)ws;
  int* pi = (int*)((char*)ws + align_up((size_t)B * kArgmaxChunks * 4, 256));
  const size_t n = (size_t)V * V * V;
  argmax_partial_kernel<<<dim3(kArgmaxChunks, B), 256, 0, st>>>(q_trans, n, kArgmaxChunks, pv, pi);
  VXB_LAUNCH_CHECK();
  select_action_final_kernel<<<cdiv(B, 64), 64, 0, st>>>(pv, pi, kArgmaxChunks, rot_grip, collision, bounds,
                                                         Bb, B, V, R, coords, rot_grip_idx, coll_idx,
                                                         attention_xyz);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ---------------------------------------------------------------- act() tail
// continuous_action = [attention_coordinate(3), quaternion(4), gripper(1), ignore_collision(1)] as QAttentionStackAgent.act
// assembles it on the host through numpy / scipy (reference qattention_stack_agent.py:78-89 with
// helpers/utils.py:103-105 discrete_euler_to_quaternion = Rotation.from_euler('xyz', idx*res - 180, degrees).as_quat()):
// extrinsic x-y-z Euler angles -> q = qz * (qy * qx), scalar-last, evaluated in float64 like scipy.
namespace vxb {
__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double o[4]) {   // (x, y, z, w), o = a * b
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}
__global__ void act_tail_kernel(const int32_t* __restrict__ rg, const int32_t* __restrict__ coll, const float* __restrict__ xyz,
                                double resolution, float* __restrict__ action, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double q[3][4];
  for (int a = 0; a < 3; ++a) {
    const double half = ((double)rg[b * 4 + a] * resolution - 180.0) * (3.14159265358979323846 / 180.0) * 0.5;
    q[a][0] = q[a][1] = q[a][2] = 0.0;
    q[a][a] = sin(half);
    q[a][3] = cos(half);
  }
  double t[4], r[4];
  quat_mul(q[1], q[0], t);      // qy * qx
  quat_mul(q[2], t, r);         // qz * (qy * qx)
  float* o = action + (size_t)b * 9;
  o[0] = xyz[b * 3 + 0]; o[1] = xyz[b * 3 + 1]; o[2] = xyz[b * 3 + 2];
  o[3] = (float)r[0]; o[4] = (float)r[1]; o[5] = (float)r[2]; o[6] = (float)r[3];
  o[7] = (float)rg[b * 4 + 3];
  o[8] = (float)coll[b];
}
}  // namespace vxb

extern "C" int vxb_act_tail_f32(const int32_t* rot_grip_idx, const int32_t* coll_idx, const float* attention_xyz,
                                float rotation_resolution, float* action, int B, void* stream) {
  VXB_CHECK_ARG(rot_grip_idx && coll_idx && attention_xyz && action && B > 0, "act_tail: bad arguments");
  act_tail_kernel<<<cdiv(B, 64), 64, 0, (cudaStream_t)stream>>>(rot_grip_idx, coll_idx, attention_xyz,
                                                                (double)rotation_resolution, action, B);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ---------------------------------------------------------------- SE(3) augmentation of the point clouds
// perturb_se3 (reference peract/voxel/augmentation.py:7-65) on one camera's planar point cloud [B, 3, N]:
//   p' = (p - a) . R + c,   a = keyframe gripper position, R = rot_shift[0:3, 0:3] (row vector times matrix, :40-41),
//   c = clamp(a + trans_shift, scene bounds) (:44-58) -- one fused streaming pass instead of ~10 full-size torch temporaries
//   (ones, repeat, transpose, bmm, transpose, stack, add) per camera.  xform [B,15] = a(3), R(9, row-major), c(3).
namespace vxb {
__global__ void __launch_bounds__(256)
se3_perturb_kernel(const float* __restrict__ pcd, const float* __restrict__ xform, float* __restrict__ out, long long N) {
  const int b = blockIdx.y;
  __shared__ float xf[15];
  if (threadIdx.x < 15) xf[threadIdx.x] = xform[b * 15 + threadIdx.x];
  __syncthreads();
  const float* px = pcd + (size_t)b * 3 * N;
  float* ox = out + (size_t)b * 3 * N;
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
    const float d0 = __fsub_rn(px[n], xf[0]), d1 = __fsub_rn(px[N + n], xf[1]), d2 = __fsub_rn(px[2 * N + n], xf[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float r = fmaf(d2, xf[3 + 6 + k], fmaf(d1, xf[3 + 3 + k], __fmul_rn(d0, xf[3 + k])));
      ox[k * N + n] = __fadd_rn(r, xf[12 + k]);
    }
  }
}
}  // namespace vxb

extern "C" int vxb_se3_perturb_f32(const float* pcd, const float* xform, float* out, int B, long long N, void* stream) {
  VXB_CHECK_ARG(pcd && xform && out && B > 0 && N > 0, "se3_perturb: bad arguments");
  const int bx = (int)std::max<long long>(1, std::min<long long>((N + 255) / 256, (148 * 8 + B - 1) / B));
  se3_perturb_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(pcd, xform, out, N);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ---------------------------------------------------------------- building blocks
extern "C" long long vxb_umma_launch_count(void) { return umma::launches(); }

extern "C" size_t vxb_linear_workspace_bytes(int M, int N, int K) {
  return umma::linear_scratch_bytes(M, N, K, true) + 256;
}

extern "C" int vxb_linear_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                              const float* residual, int res_rows, float* C, int ldc, int M, int N,
                              int K, float alpha, float act_slope, int math_mode, void* ws,
                              size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(A && W && C && M > 0 && N > 0 && K >= 0, "linear: bad arguments");
  Arena scratch(ws, ws_bytes);
  return linear(A, lda, W, ldw, bias, residual, res_rows, ldc, C, ldc, M, N, K, alpha, act_slope,
                math_mode, (cudaStream_t)stream, ws ? &scratch : nullptr, nullptr);
}

extern "C" int vxb_layernorm_f32(const float* x, const float* w, const float* b, float* y, int rows,
                                 int n, void* stream) {
  VXB_CHECK_ARG(x && w && b && y && rows > 0 && n > 0 && n % 4 == 0, "layernorm: bad arguments");
  layernorm_kernel<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, w, b, y, rows, n, rows, 0);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

extern "C" size_t vxb_spatial_softmax_workspace_bytes(int B, int P, int C) {
  return ss_partial_floats((size_t)P, B, C) * sizeof(float);
}

extern "C" int vxb_spatial_softmax_f32(const float* x, int B, int Dd, int Hh, int Ww, int C, float* ss,
                                       int ss_stride, float* mx, int mx_stride, void* ws,
                                       size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(x && ss && ws && B > 0 && C > 0, "spatial_softmax: bad arguments");
  const size_t P = (size_t)Dd * Hh * Ww;
  if (ws_bytes < vxb_spatial_softmax_workspace_bytes(B, (int)P, C)) {
    set_error("spatial_softmax: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  return spatial_softmax_run(x, B, Dd, Hh, Ww, C, ss, ss_stride, mx, mx_stride, (float*)ws, (cudaStream_t)stream);
}

static size_t conv_w_bytes(int Ci, int Co, int k) { return align_up((size_t)Ci * Co * k * k * k * sizeof(float), 256); }

extern "C" size_t vxb_conv3d_workspace_bytes(int B, int Di, int Ci, int Co, int k) {
  // tap-major fp32 weight + (bf16x3 path) its planes and the padded input planes
  return 2 * conv_w_bytes(Ci, Co, k) + std::max(umma::conv3d_scratch_bytes(B, Di, Ci, 0, k), umma::conv3_f8c_scratch_bytes(B, Di)) + 1024;
}

extern "C" int vxb_conv3d_f32(const float* x, const float* w, const float* bias, float* y, int B,
                              int Di, int Ci, int Co, int k, int s, float act_slope, int math_mode,
                              void* ws, size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(x && w && y && ws && B > 0 && Di > 0 && (k & 1) && s > 0, "conv3d: bad arguments");
  if (ws_bytes < vxb_conv3d_workspace_bytes(B, Di, Ci, Co, k)) {
    set_error("conv3d: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int k3 = k * k * k;
  conv_weight_to_tapmajor_kernel<<<cdiv((size_t)Co * Ci * k3, 256), 256, 0, st>>>(w, (float*)ws, Co, Ci, k3);
  VXB_LAUNCH_CHECK();
  const int Do = (Di + 2 * (k / 2) - k) / s + 1;
  if (Co == 1 && Ci == 64 && k == 3 && s == 1 && act_slope < 0.f)   // trans_decoder shape: streaming stencil
    return trans_stencil_run<64>(x, (const float*)ws, bias, y, B, Di, st);
  Arena scratch((char*)ws + conv_w_bytes(Ci, Co, k), ws_bytes - conv_w_bytes(Ci, Co, k));
  if (math_mode == VXB_MATH_F16X3 && s > 1 && Co == 64 && Ci == 64 && bias) {
    // strided (patchify) convolution: gather-loader tcgen05 kernel (patchify_umma.cuh)
    __nv_bfloat16* wc = scratch.get<__nv_bfloat16>(umma::patchify_weight_elems(k));
    if (!scratch.ok) {
      set_error("conv3d: workspace too small");
      return VXB_E_WORKSPACE_TOO_SMALL;
    }
    VXB_TRY(umma::patchify_prepare_weights((const float*)ws, k, wc, st));
    return umma::patchify_f32(x, wc, bias, act_slope, y, B, Di, k, s, st);
  }
  if (math_mode == VXB_MATH_F16F8C) {
    // fp16 + E4M3-corrected input-stationary convolution (conv_f8c.cuh); only the final-convolution geometry exists
    VXB_CHECK_ARG(k == 3 && s == 1 && Co == 64 && Ci == 64 && bias, "conv3d: VXB_MATH_F16F8C needs k=3, s=1, Ci=Co=64 and a bias");
    return umma::conv3_f8c_f32(x, (const float*)ws, bias, act_slope, y, B, Di, scratch, st);
  }
  if (math_mode == VXB_MATH_F16X3 && k == 3 && s == 1 && Co == 64 && Ci == 64 && bias) {
    // input-stationary tcgen05 convolution (conv_umma.cuh): padded hi/lo planes + re-laid weights
    __nv_bfloat16* wc = scratch.get<__nv_bfloat16>(umma::conv3_weight_elems(Ci));
    umma::Planes xp;
    const long long prow = (long long)B * (Di + 2) * (Di + 2) * (Di + 2);
    xp.hi = scratch.get<__nv_bfloat16>((size_t)prow * 64);
    xp.lo = scratch.get<__nv_bfloat16>((size_t)prow * 64);
    xp.ld = 64;
    if (!scratch.ok) {
      set_error("conv3d: workspace too small");
      return VXB_E_WORKSPACE_TOO_SMALL;
    }
    VXB_TRY(umma::conv3_prepare_weights((const float*)ws, Ci, wc, st));
    VXB_TRY(umma::pad_split(x, B, Di, 1, Ci, xp, st));
    return umma::conv3_planes(xp, nullptr, Ci, 0, wc, bias, act_slope, y, B, Di, st);
  }
  return conv3d(x, nullptr, Ci, 0, (const float*)ws, bias, y, B, Di, Do, Co, k, s, act_slope, math_mode, st,
                &scratch, nullptr);
}

static size_t fold_w_bytes(int Ci, int Co, int s) { return align_up((size_t)s * s * s * Co * 27 * Ci * sizeof(float), 256); }

extern "C" size_t vxb_upconv3d_workspace_bytes(int B, int S, int Ci, int Co, int k, int s) {
  (void)k;
  return 2 * fold_w_bytes(Ci, Co, s) + umma::upconv_scratch_bytes(B, S, Ci) + 1024;
}

extern "C" int vxb_upconv3d_f32(const float* x, const float* w, const float* bias, float* y, int B,
                                int S, int Ci, int Co, int k, int s, float act_slope, int math_mode,
                                void* ws, size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(x && w && y && ws && B > 0 && S > 0 && (k & 1) && s > 0, "upconv3d: bad arguments");
  if (2 * (k / 2) > s + 1) {
    set_error("upconv3d: folding needs k/2 <= (s+1)/2 (k=%d, s=%d)", k, s);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (ws_bytes < vxb_upconv3d_workspace_bytes(B, S, Ci, Co, k, s)) {
    set_error("upconv3d: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)s * s * s * Co * 27 * Ci;
  fold_upconv_weights_kernel<<<cdiv(total, 256), 256, 0, st>>>(w, (float*)ws, Co, Ci, k, s);
  VXB_LAUNCH_CHECK();
  Arena scratch((char*)ws + fold_w_bytes(Ci, Co, s), ws_bytes - fold_w_bytes(Ci, Co, s));
  return upconv3d_folded(x, (const float*)ws, bias, y, B, S, Ci, Co, s, act_slope, math_mode, st, &scratch, nullptr);
}

extern "C" size_t vxb_attention_workspace_bytes(int B, int H, int Nq, int Nk) {
  const size_t Nkp = ((size_t)Nk + 3) / 4 * 4;
  return std::max(align_up((size_t)B * H * Nq * Nkp * sizeof(float), 256),
                  umma::attention_f32_scratch_bytes(B, H, Nq, Nk, 64) + 256);
}

extern "C" int vxb_attention_f32(const float* q, int ldq, long long q_batch_stride, const float* k,
                                 const float* v, int ldkv, long long kv_batch_stride, float* out,
                                 int ldo, long long o_batch_stride, int B, int H, int Nq, int Nk,
                                 int dh, float scale, int math_mode, void* ws, size_t ws_bytes,
                                 void* stream) {
  VXB_CHECK_ARG(q && k && v && out && ws && B > 0 && H > 0 && Nq > 0 && Nk > 0 && dh > 0,
                "attention: bad arguments");
  if (ws_bytes < vxb_attention_workspace_bytes(B, H, Nq, Nk)) {
    set_error("attention: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  if (math_mode == VXB_MATH_F16X3 && dh == 64) {
    Arena scratch(ws, ws_bytes);
    return umma::attention_f32(q, ldq, q_batch_stride, k, v, ldkv, kv_batch_stride, out, ldo, o_batch_stride, B, H,
                               Nq, Nk, dh, scale, scratch, (cudaStream_t)stream);
  }
  return attention_materialized(q, ldq, q_batch_stride, k, v, ldkv, kv_batch_stride, out, ldo,
                                o_batch_stride, B, H, Nq, Nk, dh, scale, (float*)ws, math_mode,
                                (cudaStream_t)stream);
}

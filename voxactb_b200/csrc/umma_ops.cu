// Host side of the tcgen05 split-bf16 engine: TMA tensor maps, operand splitting / padding kernels
// and the launch wrappers used by dispatch.cuh.
#include "umma_gemm.cuh"
#include "conv_umma.cuh"
#include "conv_f8c.cuh"
#include "patchify_umma.cuh"
#include "flash_umma.cuh"
#include <cstdlib>
#include "umma_host.cuh"
#include <cudaTypedefs.h>
#include <mutex>

namespace vxb {
namespace umma {

// ------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2D bf16 row-major [rows, cols] with `ld` elements between rows; box = 64 columns x box_rows rows,
// SWIZZLE_128B (64 bf16 = 128 B inner extent), out-of-bounds elements read as zero.
static int make_map(CUtensorMap* map, const __nv_bfloat16* base, long long rows, long long cols, long long ld,
                    int box_rows, int box_cols = 64) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return VXB_E_CUDA;
  }
  if (((uintptr_t)base & 15) || ((ld * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", ld);
    return VXB_E_BADARG;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r, rows, cols, ld, box_rows);
    return VXB_E_CUDA;
  }
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------ split / pad kernels
__device__ __forceinline__ void split8(const float* f, uint4& hi, uint4& lo) {
  __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    h[t] = pl_from_float(f[t]);
    l[t] = pl_from_float(f[t] - pl_to_float(h[t]));
  }
  hi = *reinterpret_cast<const uint4*>(h);
  lo = *reinterpret_cast<const uint4*>(l);
}

// fp32 [rows, cols] (ldx) -> bf16 hi/lo planes [rows, ldp]; columns [cols, ldp) are zero-filled
static __global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols,
                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp) {
  const long long groups_per_row = ldp / 8;
  const long long total = rows * groups_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups_per_row;
    const int c = (int)(i % groups_per_row) * 8;
    float f[8];
    const float* src = x + r * ldx + c;
    if (c + 7 < cols && ((ldx & 3) == 0) && (((uintptr_t)x & 15) == 0)) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 b = *reinterpret_cast<const float4*>(src + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] = (c + t < cols) ? src[t] : 0.f;
    }
    uint4 h, l;
    split8(f, h, l);
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = h;
    *reinterpret_cast<uint4*>(lo + r * ldp + c) = l;
  }
}

// fp32 channels-last [B, V, V, V, C] -> 16-bit hi/lo planes of the replicate-padded grid [B, Vp, Vp, Vp, C], Vp = V + 2*pad
static __global__ void __launch_bounds__(256)
pad_split_kernel(const float* __restrict__ x, int B, int V, int pad, int C,
                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int Vp = V + 2 * pad;
  const int cg = C / 8;
  const long long total = (long long)B * Vp * Vp * Vp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 8;
    long long v = i / cg;
    const int pw = (int)(v % Vp); v /= Vp;
    const int ph = (int)(v % Vp); v /= Vp;
    const int pd = (int)(v % Vp);
    const int b = (int)(v / Vp);
    const int d = min(max(pd - pad, 0), V - 1), h = min(max(ph - pad, 0), V - 1), w = min(max(pw - pad, 0), V - 1);
    const float* src = x + ((((long long)b * V + d) * V + h) * V + w) * C + c;
    const float4 a = *reinterpret_cast<const float4*>(src);
    const float4 bb = *reinterpret_cast<const float4*>(src + 4);
    const float f[8] = {a.x, a.y, a.z, a.w, bb.x, bb.y, bb.z, bb.w};
    uint4 hh, ll;
    split8(f, hh, ll);
    const long long o = (i / cg) * C + c;
    *reinterpret_cast<uint4*>(hi + o) = hh;
    *reinterpret_cast<uint4*>(lo + o) = ll;
  }
}

// replicate-fill the halo of padded 16-bit planes in place: every halo voxel copies its nearest interior voxel.
// Only the halo voxels are enumerated (two full z planes + the border ring of every other plane), pad = 1 fast path.
static __global__ void __launch_bounds__(256)
halo_fill_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ third, int B, int V, int pad, int C) {
  const int Vp = V + 2 * pad;
  const int cg = C / 8;
  if (pad == 1) {
    const long long plane = (long long)Vp * Vp, ring = 4ll * Vp - 4;
    const long long per_b = 2 * plane + (long long)(Vp - 2) * ring;
    const long long total = (long long)B * per_b * cg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % cg) * 8;
      long long v = i / cg;
      const int b = (int)(v / per_b);
      v -= (long long)b * per_b;
      int pd, ph, pw;
      if (v < 2 * plane) {
        pd = v < plane ? 0 : Vp - 1;
        const long long r = v % plane;
        ph = (int)(r / Vp); pw = (int)(r % Vp);
      } else {
        const long long j = v - 2 * plane;
        pd = 1 + (int)(j / ring);
        const int r = (int)(j % ring);
        if (r < Vp) { ph = 0; pw = r; }
        else if (r < 2 * Vp) { ph = Vp - 1; pw = r - Vp; }
        else { const int rr = r - 2 * Vp; ph = 1 + rr / 2; pw = (rr & 1) ? Vp - 1 : 0; }
      }
      const int sd = min(max(pd, 1), V), sh = min(max(ph, 1), V), sw = min(max(pw, 1), V);
      const long long so = ((((long long)b * Vp + sd) * Vp + sh) * Vp + sw) * C + c;
      const long long o = ((((long long)b * Vp + pd) * Vp + ph) * Vp + pw) * C + c;
      *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hi + so);
      if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(lo + so);
      if (third) *reinterpret_cast<uint4*>(third + o) = *reinterpret_cast<const uint4*>(third + so);
    }
    return;
  }
  const long long total = (long long)B * Vp * Vp * Vp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 8;
    long long v = i / cg;
    const int pw = (int)(v % Vp); v /= Vp;
    const int ph = (int)(v % Vp); v /= Vp;
    const int pd = (int)(v % Vp);
    const int b = (int)(v / Vp);
    const int sd = min(max(pd, pad), V + pad - 1), sh = min(max(ph, pad), V + pad - 1), sw = min(max(pw, pad), V + pad - 1);
    if (sd == pd && sh == ph && sw == pw) continue;   // interior
    const long long so = ((((long long)b * Vp + sd) * Vp + sh) * Vp + sw) * C + c;
    const long long o = (i / cg) * C + c;
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hi + so);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(lo + so);
    if (third) *reinterpret_cast<uint4*>(third + o) = *reinterpret_cast<const uint4*>(third + so);
  }
}

int split_rows(const float* x, long long ldx, long long rows, int cols, Planes out, cudaStream_t st) {
  if (out.ld % 8 || out.ld < cols) {
    set_error("split_rows: plane ld must be a multiple of 8 and >= cols");
    return VXB_E_BADARG;
  }
  const long long total = rows * (out.ld / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  split_rows_kernel<<<blocks, 256, 0, st>>>(x, ldx, rows, cols, out.hi, out.lo, out.ld);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

int pad_split(const float* x, int B, int V, int pad, int C, Planes out, cudaStream_t st) {
  if (C % 8) {
    set_error("pad_split: C must be a multiple of 8");
    return VXB_E_BADARG;
  }
  const int Vp = V + 2 * pad;
  const long long total = (long long)B * Vp * Vp * Vp * (C / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  pad_split_kernel<<<blocks, 256, 0, st>>>(x, B, V, pad, C, out.hi, out.lo);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

int halo_fill(Planes p, int B, int V, int pad, int C, cudaStream_t st, __nv_bfloat16* third) {
  const int Vp = V + 2 * pad;
  long long total = (long long)B * Vp * Vp * Vp * (C / 8);
  if (pad == 1) total = (long long)B * (2ll * Vp * Vp + (long long)(Vp - 2) * (4ll * Vp - 4)) * (C / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  halo_fill_kernel<<<blocks, 256, 0, st>>>(p.hi, p.lo, third, B, V, pad, C);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// 8 consecutive values of a row -> plane stores.  f8a == 0: fp16 hi / lo planes.  f8a > 0 (f8c A operand, see umma_gemm.cuh
// terms == 2): hi = fp16(32 f8a x) (an exact power-of-two multiple of fp16(x)), lo plane = c8 blocks per 64 columns.
__device__ __forceinline__ void store8_planes(const float* f, __nv_bfloat16* hi, __nv_bfloat16* lo, long long row, long long ldp,
                                              int i, float f8a) {
  if (f8a == 0.f) {
    uint4 hh, ll;
    split8(f, hh, ll);
    *reinterpret_cast<uint4*>(hi + row * ldp + i) = hh;
    *reinterpret_cast<uint4*>(lo + row * ldp + i) = ll;
    return;
  }
  __align__(16) __nv_bfloat16 h[8];
  float l[8];
  const float s16 = f8a * 32.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float hf = pl_to_float(pl_from_float(f[t]));
    l[t] = f[t] - hf;
    h[t] = pl_from_float(hf * s16);
  }
  *reinterpret_cast<uint4*>(hi + row * ldp + i) = *reinterpret_cast<const uint4*>(h);
  uint8_t* rowb = reinterpret_cast<uint8_t*>(lo + row * ldp) + (i >> 6) * 128 + (i & 63);
  uint2 lo8, hi8;
  lo8.x = pl_e4m3x4(l[0], l[1], l[2], l[3], f8a * 2048.f); lo8.y = pl_e4m3x4(l[4], l[5], l[6], l[7], f8a * 2048.f);
  hi8.x = pl_e4m3x4(f[0], f[1], f[2], f[3], f8a); hi8.y = pl_e4m3x4(f[4], f[5], f[6], f[7], f8a);
  *reinterpret_cast<uint2*>(rowb) = lo8;
  *reinterpret_cast<uint2*>(rowb + 64) = hi8;
}

// LayerNorm (eps 1e-5, biased variance; reference PreNorm, perceiver_lang_io.py:56-71) -> planes; warp per row
static __global__ void __launch_bounds__(256)
layernorm_planes_kernel(const float* __restrict__ x, size_t x_batch_stride, int rows_per_batch,
                        const float* __restrict__ w, const float* __restrict__ b,
                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp, long long rows, int n, const float* __restrict__ f8alpha) {
  const float f8a = f8alpha ? __ldg(f8alpha) : 0.f;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)(row / rows_per_batch) * x_batch_stride + (size_t)(row % rows_per_batch) * n;
  float s = 0.f;
  for (int i = lane * 8; i < n; i += 256) {
    const float4 a = *reinterpret_cast<const float4*>(xr + i), c = *reinterpret_cast<const float4*>(xr + i + 4);
    s += (a.x + a.y + a.z + a.w) + (c.x + c.y + c.z + c.w);
  }
  s = warp_sum(s);
  const float mean = s / (float)n;
  float q = 0.f;
  for (int i = lane * 8; i < n; i += 256) {
    const float4 a = *reinterpret_cast<const float4*>(xr + i), c = *reinterpret_cast<const float4*>(xr + i + 4);
    const float d0 = a.x - mean, d1 = a.y - mean, d2 = a.z - mean, d3 = a.w - mean;
    const float d4 = c.x - mean, d5 = c.y - mean, d6 = c.z - mean, d7 = c.w - mean;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)n + 1e-5f);
  for (int i = lane * 8; i < n; i += 256) {
    float f[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) f[t] = (xr[i + t] - mean) * rstd * w[i + t] + b[i + t];
    store8_planes(f, hi, lo, row, ldp, i, f8a);
  }
}

// Same arithmetic with the row held in registers (n = CH * 256): x is read once instead of three times.
template <int CH>
static __global__ void __launch_bounds__(256)
layernorm_planes_reg_kernel(const float* __restrict__ x, size_t x_batch_stride, int rows_per_batch,
                            const float* __restrict__ w, const float* __restrict__ b,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp, long long rows,
                            const float* __restrict__ f8alpha) {
  constexpr int n = CH * 256;
  const float f8a = f8alpha ? __ldg(f8alpha) : 0.f;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)(row / rows_per_batch) * x_batch_stride + (size_t)(row % rows_per_batch) * n;
  float4 v[CH][2];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    v[c][0] = *reinterpret_cast<const float4*>(xr + c * 256 + lane * 8);
    v[c][1] = *reinterpret_cast<const float4*>(xr + c * 256 + lane * 8 + 4);
  }
#pragma unroll
  for (int c = 0; c < CH; ++c)
    s += (v[c][0].x + v[c][0].y + v[c][0].z + v[c][0].w) + (v[c][1].x + v[c][1].y + v[c][1].z + v[c][1].w);
  s = warp_sum(s);
  const float mean = s / (float)n;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const float d0 = v[c][0].x - mean, d1 = v[c][0].y - mean, d2 = v[c][0].z - mean, d3 = v[c][0].w - mean;
    const float d4 = v[c][1].x - mean, d5 = v[c][1].y - mean, d6 = v[c][1].z - mean, d7 = v[c][1].w - mean;
    q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3 + d4 * d4 + d5 * d5 + d6 * d6 + d7 * d7;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)n + 1e-5f);
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int i = c * 256 + lane * 8;
    const float4 w0 = *reinterpret_cast<const float4*>(w + i), w1 = *reinterpret_cast<const float4*>(w + i + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(b + i), b1 = *reinterpret_cast<const float4*>(b + i + 4);
    float f[8];
    f[0] = (v[c][0].x - mean) * rstd * w0.x + b0.x; f[1] = (v[c][0].y - mean) * rstd * w0.y + b0.y;
    f[2] = (v[c][0].z - mean) * rstd * w0.z + b0.z; f[3] = (v[c][0].w - mean) * rstd * w0.w + b0.w;
    f[4] = (v[c][1].x - mean) * rstd * w1.x + b1.x; f[5] = (v[c][1].y - mean) * rstd * w1.y + b1.y;
    f[6] = (v[c][1].z - mean) * rstd * w1.z + b1.z; f[7] = (v[c][1].w - mean) * rstd * w1.w + b1.w;
    store8_planes(f, hi, lo, row, ldp, i, f8a);
  }
}

static __global__ void ln_f8c_alpha_kernel(const float* __restrict__ w, const float* __restrict__ b, int n, float* __restrict__ out) {
  __shared__ unsigned int mx;
  if (threadIdx.x == 0) mx = 0u;
  __syncthreads();
  const float r = sqrtf((float)(n - 1));
  float m = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fmaf(fabsf(w[i]), r, fabsf(b[i])));
  atomicMax(&mx, __float_as_uint(m * 1.0001f));
  __syncthreads();
  if (threadIdx.x == 0) {
    const float bound = __uint_as_float(mx);
    int e = 0;
    if (bound > 0.f && isfinite(bound)) e = max(-60, min(60, ilogbf(240.f / bound)));
    out[0] = scalbnf(1.f, e);
  }
}
int layernorm_f8c_alpha(const float* w, const float* b, int n, float* alpha_out, cudaStream_t st) {
  ln_f8c_alpha_kernel<<<1, 256, 0, st>>>(w, b, n, alpha_out);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
static __global__ void f8c_unscale_kernel(const float* a, const float* b, float* o) { o[0] = 1.f / (a[0] * b[0]); }
int f8c_unscale(const float* alpha, const float* beta, float* out, cudaStream_t st) {
  f8c_unscale_kernel<<<1, 1, 0, st>>>(alpha, beta, out);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

int layernorm_planes(const float* x, size_t x_batch_stride, int rows_per_batch, const float* w, const float* b,
                     Planes out, long long rows, int n, cudaStream_t st, const float* f8alpha) {
  if (n % 8 || out.ld < n || (f8alpha && (n % 64 || out.ld != n))) {
    set_error("layernorm_planes: n must be a multiple of 8 (64 with ld == n for f8c planes) (n=%d)", n);
    return VXB_E_BADARG;
  }
  const bool aligned = !(((uintptr_t)w | (uintptr_t)b) & 15);
  if (n == 512 && aligned) {
    layernorm_planes_reg_kernel<2><<<cdiv(rows, 8), 256, 0, st>>>(x, x_batch_stride, rows_per_batch, w, b, out.hi, out.lo,
                                                                  out.ld, rows, f8alpha);
  } else if (n == 256 && aligned) {
    layernorm_planes_reg_kernel<1><<<cdiv(rows, 8), 256, 0, st>>>(x, x_batch_stride, rows_per_batch, w, b, out.hi, out.lo,
                                                                  out.ld, rows, f8alpha);
  } else {
    layernorm_planes_kernel<<<cdiv(rows, 8), 256, 0, st>>>(x, x_batch_stride, rows_per_batch, w, b, out.hi, out.lo, out.ld,
                                                          rows, n, f8alpha);
  }
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// GEGLU (perceiver_lang_io.py:74-77, exact erf gelu) -> planes
static __global__ void __launch_bounds__(256)
geglu_planes_kernel(const float* __restrict__ h, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                    long long ldp, long long rows, int n) {
  const long long total8 = rows * (long long)(n / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (n / 8);
    const int c = (int)(i % (n / 8)) * 8;
    const float* a = h + r * 2 * n + c;
    const float* g = a + n;
    float f[8];
#pragma unroll
    for (int t = 0; t < 8; t += 4) {
      const float4 av = *reinterpret_cast<const float4*>(a + t);
      const float4 gv = *reinterpret_cast<const float4*>(g + t);
      f[t + 0] = av.x * (0.5f * gv.x * (1.f + erff(gv.x * 0.70710678118654752f)));
      f[t + 1] = av.y * (0.5f * gv.y * (1.f + erff(gv.y * 0.70710678118654752f)));
      f[t + 2] = av.z * (0.5f * gv.z * (1.f + erff(gv.z * 0.70710678118654752f)));
      f[t + 3] = av.w * (0.5f * gv.w * (1.f + erff(gv.w * 0.70710678118654752f)));
    }
    uint4 hh, ll;
    split8(f, hh, ll);
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = hh;
    *reinterpret_cast<uint4*>(lo + r * ldp + c) = ll;
  }
}

int geglu_planes(const float* h, Planes out, long long rows, int n, cudaStream_t st) {
  if (n % 8 || out.ld < n) {
    set_error("geglu_planes: n must be a multiple of 8 (n=%d)", n);
    return VXB_E_BADARG;
  }
  const long long total8 = rows * (n / 8);
  geglu_planes_kernel<<<(int)std::min<long long>((total8 + 255) / 256, 148 * 16), 256, 0, st>>>(h, out.hi, out.lo, out.ld,
                                                                                                 rows, n);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

static __global__ void rowstat_init_kernel(float* __restrict__ mx, float* __restrict__ sum, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    mx[i] = -INFINITY;
    sum[i] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------ launch
template <int NT, int STAGES, int EPI>
static int launch_e(const CUtensorMap* maps, const Params& p, cudaStream_t st) {
  using L = SmemLayout<NT, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    VXB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<NT, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VXB_CUDA(cudaGetDevice(&dev));
    VXB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long tiles = (long long)p.m_tiles * p.n_tiles * p.batches;
  const int grid = (int)std::min<long long>(tiles, num_sms);
  umma_gemm_kernel<NT, STAGES, EPI><<<grid, gemm_threads(EPI), L::TOTAL, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// epilogue specialisation from the runtime description
template <int NT, int STAGES>
static int launch_t(const CUtensorMap* maps, const Params& p, cudaStream_t st) {
  const Epilogue& e = p.ep;
  if (e.row_mode == ROWS_PLAIN && !e.transpose_planes) {
    if (e.mode == EPI_ROWMAX) return launch_e<NT, STAGES, EPIK_ROWMAX>(maps, p, st);
    if (e.mode == EPI_EXP) return launch_e<NT, STAGES, EPIK_EXP>(maps, p, st);
    if (e.mode == EPI_GEGLU) return launch_e<NT, STAGES, EPIK_GEGLU>(maps, p, st);
    return launch_e<NT, STAGES, EPIK_PLAIN>(maps, p, st);
  }
  if (e.row_mode == ROWS_PHASE && e.mode == EPI_STORE && !e.residual && !e.row_div && !e.transpose_planes && e.N % 64 == 0 &&
      (!e.bias || (reinterpret_cast<uintptr_t>(e.bias) & 15) == 0) && (!e.out_f32 || ((e.ldc & 3) == 0 &&
      (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0)) && (!e.out_hi || (e.ldp & 3) == 0))
    return launch_e<NT, STAGES, EPIK_PHASE>(maps, p, st);
  return launch_e<NT, STAGES, EPIK_GENERIC>(maps, p, st);
}

static long long g_umma_launches = 0;
long long launches() { return g_umma_launches; }

int gemm(const Operand& A0, const Operand* A1, const Operand& W, int n_tile, Params p, cudaStream_t st) {
  CUtensorMap maps[6];
  ++g_umma_launches;
  VXB_TRY(make_map(&maps[0], A0.p.hi, A0.rows, A0.cols, A0.p.ld, BM));
  VXB_TRY(make_map(&maps[1], A0.p.lo, A0.rows, A0.cols, A0.p.ld, BM));
  const Operand& a1 = A1 ? *A1 : A0;
  VXB_TRY(make_map(&maps[2], a1.p.hi, a1.rows, a1.cols, a1.p.ld, BM));
  VXB_TRY(make_map(&maps[3], a1.p.lo, a1.rows, a1.cols, a1.p.ld, BM));
  VXB_TRY(make_map(&maps[4], W.p.hi, W.rows, W.cols, W.p.ld, n_tile));
  VXB_TRY(make_map(&maps[5], W.p.lo, W.rows, W.cols, W.p.ld, n_tile));
  if (p.batches <= 0) p.batches = 1;
  if (p.Hz <= 0) p.Hz = 1;
  switch (n_tile) {
    case 64: return launch_t<64, 4>(maps, p, st);
    case 128: return launch_t<128, 3>(maps, p, st);
    case 256: return launch_t<256, 2>(maps, p, st);
    default:
      set_error("umma::gemm: unsupported n_tile %d", n_tile);
      return VXB_E_UNSUPPORTED_SHAPE;
  }
}

}  // namespace umma
}  // namespace vxb

// ========================================================================================== fp32-in / fp32-out wrappers
// (operands are split into 16-bit hi/lo planes in `scratch`; weights may come pre-split from vxb_qnet_prepare)
namespace vxb {
namespace umma {

static Planes alloc_planes(Arena& a, long long rows, long long ld) {
  Planes p;
  p.hi = a.get<__nv_bfloat16>(plane_elems(rows, ld));
  p.lo = a.get<__nv_bfloat16>(plane_elems(rows, ld));
  p.ld = ld;
  return p;
}

static int pick_ntile(int N) { return N >= 256 ? 256 : (N > 64 ? 128 : 64); }
// small-M GEMMs (batch-1 acting: M = 2048 latent rows) leave most SMs idle with 128 x 256 tiles: halve the N tile until the
// launch has at least one tile per SM
static int pick_ntile_mn(long long m_tiles, int N) {
  int nt = pick_ntile(N);
  while (nt > 64 && m_tiles * cdiv(N, nt) < 148) nt >>= 1;
  return nt;
}

size_t linear_scratch_bytes(long long M, long long N, long long K, bool split_w) {
  Arena a(nullptr, 0);
  alloc_planes(a, M, pad8(K));
  if (split_w) alloc_planes(a, N, pad8(K));
  return a.off;
}

int linear_f32(const float* A, int lda, const float* W, int ldw, const Planes* Wpre, const float* bias,
               const float* residual, int res_rows, int ldr, float* C, int ldc, int M, int N, int K, float alpha,
               float act_slope, Arena& scratch, cudaStream_t st) {
  const long long Kp = pad8(K);
  Planes Ap = alloc_planes(scratch, M, Kp);
  Planes Wp;
  if (Wpre) Wp = *Wpre; else Wp = alloc_planes(scratch, N, Kp);
  if (!scratch.ok) {
    set_error("umma linear: scratch too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  VXB_TRY(split_rows(A, lda, M, K, Ap, st));
  if (!Wpre) VXB_TRY(split_rows(W, ldw, N, K, Wp, st));
  Params p;
  params_init(p);
  const int nt = pick_ntile(N);
  p.m_tiles = cdiv(M, BM);
  p.n_tiles = cdiv(N, nt);
  p.plan.num_kb = cdiv(K, BK);
  p.ep.M = M; p.ep.N = N; p.ep.row_mode = ROWS_PLAIN;
  p.ep.bias = bias; p.ep.alpha = alpha; p.ep.act_slope = act_slope;
  p.ep.residual = residual; p.ep.res_rows = res_rows > 0 ? res_rows : 1; p.ep.ldr = ldr;
  p.ep.out_f32 = C; p.ep.ldc = ldc;
  Operand a{Ap, M, K}, w{Wp, N, K};
  return gemm(a, nullptr, w, nt, p, st);
}

size_t conv3d_scratch_bytes(int B, int V, int C0, int C1, int k) {
  Arena a(nullptr, 0);
  const long long Vp = V + 2 * (k / 2);
  alloc_planes(a, (long long)B * Vp * Vp * Vp, C0);
  if (C1) alloc_planes(a, (long long)B * Vp * Vp * Vp, C1);
  return a.off;
}

// stride-1 convolution, replicate padding; x0 [B,V^3,C0] (+ x1 [B,V^3,C1]) fp32 compact channels-last,
// Wp planes of the tap-major weight [Co][k^3 * (C0+C1)], out fp32 compact [B,V^3,Co]
int conv3d_f32(const float* x0, const float* x1, int C0, int C1, const Planes& Wp, const float* bias, float* out,
               int B, int V, int Co, int k, float act_slope, Arena& scratch, cudaStream_t st) {
  const int pad = k / 2;
  const long long Vp = V + 2 * pad;
  const long long rows = (long long)B * Vp * Vp * Vp;
  if (C0 % 64 || C1 % 64 || Co > 256 || rows >= (1ll << 31)) {
    set_error("umma conv3d: unsupported channels/size (C0=%d C1=%d Co=%d rows=%lld)", C0, C1, Co, rows);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  Planes a0 = alloc_planes(scratch, rows, C0);
  Planes a1 = a0;
  if (C1) a1 = alloc_planes(scratch, rows, C1);
  if (!scratch.ok) {
    set_error("umma conv3d: scratch too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  VXB_TRY(pad_split(x0, B, V, pad, C0, a0, st));
  if (C1) VXB_TRY(pad_split(x1, B, V, pad, C1, a1, st));
  Params p;
  params_init(p);
  const int nt = pick_ntile(Co);
  // one batch entry per sample; the leading / trailing `pad` z planes of the padded grid hold no output voxel:
  // start at the first row that can be valid and stop after the last one
  const long long vp3 = Vp * Vp * Vp;
  const long long row_lo = ((long long)pad * Vp + pad) * Vp + pad;
  const long long row_hi = vp3 - row_lo;                       // one past the last interior row
  p.batches = B; p.a_row_zb = (int)vp3; p.a_row_off = (int)row_lo;
  p.m_tiles = cdiv(row_hi - row_lo, BM);
  p.n_tiles = cdiv(Co, nt);
  p.plan.taps = k; p.plan.Vp = (int)Vp; p.plan.cpb = (C0 + C1) / 64; p.plan.cb_src0 = C0 / 64;
  p.plan.num_kb = k * k * k * p.plan.cpb;
  p.ep.M = (int)(row_hi - row_lo); p.ep.N = Co; p.ep.row_mode = ROWS_CONV_FLAT; p.ep.Vp = (int)Vp; p.ep.pad = pad;
  p.ep.out_padded = 0;
  p.ep.bias = bias; p.ep.act_slope = act_slope;
  p.ep.out_f32 = out; p.ep.ldc = Co;
  Operand A0{a0, rows, C0}, A1{a1, rows, C1 ? C1 : C0};
  Operand w{Wp, Co, (long long)k * k * k * (C0 + C1)};
  return gemm(A0, C1 ? &A1 : nullptr, w, nt, p, st);
}

size_t upconv_scratch_bytes(int B, int S, int Ci) {
  Arena a(nullptr, 0);
  const long long Sp = S + 2;
  alloc_planes(a, (long long)B * Sp * Sp * Sp, Ci);
  return a.off;
}

// ---- f8c operands of the GEMM engine (terms == 2): 64-channel blocks of 128 bytes, [64 x lo8 | 64 x hi8] against [64 x w_hi8 | 64 x w_lo8]
constexpr float F8C_P16 = 32.f;      // the fp16 planes carry 2^5 x (A side) and 2^-5 x (W side) of the fp8 scales: both stay inside fp16
__device__ __forceinline__ uint32_t f8c_pack4(float a, float b, float c, float d, float sc) { return pl_e4m3x4(a, b, c, d, sc); }
// fp32 [B,V,V,V,C] -> planes of the replicate-padded grid: hi = fp16(32 alpha x), c8 as above with alpha, 2^11 alpha
static __global__ void __launch_bounds__(256)
pad_split_f8c64_kernel(const float* __restrict__ x, int B, int V, int C, __nv_bfloat16* __restrict__ hi, uint8_t* __restrict__ c8,
                       const float* __restrict__ alpha) {
  const int Vp = V + 2, cg = C / 4;
  const long long total = (long long)B * Vp * Vp * Vp * cg;
  const float fa = __ldg(alpha);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long v = i / cg;
    const long long row = v;
    const int pw = (int)(v % Vp); v /= Vp;
    const int ph = (int)(v % Vp); v /= Vp;
    const int pd = (int)(v % Vp);
    const int b = (int)(v / Vp);
    const int d = min(max(pd - 1, 0), V - 1), h = min(max(ph - 1, 0), V - 1), w = min(max(pw - 1, 0), V - 1);
    const float4 a = *reinterpret_cast<const float4*>(x + ((((long long)b * V + d) * V + h) * V + w) * C + g * 4);
    const __nv_bfloat162 u01 = pl2_from_floats(a.x, a.y), u23 = pl2_from_floats(a.z, a.w);      // unscaled split: x = hi + lo
    const float2 f01 = pl2_to_float2(u01), f23 = pl2_to_float2(u23);
    const float s16 = fa * F8C_P16;
    const __nv_bfloat162 h01 = pl2_from_floats(f01.x * s16, f01.y * s16), h23 = pl2_from_floats(f23.x * s16, f23.y * s16);   // exact
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(hi + row * C + g * 4) = hv;
    uint8_t* rowb = c8 + row * C * 2 + (g >> 4) * 128 + (g & 15) * 4;
    *reinterpret_cast<uint32_t*>(rowb) = f8c_pack4(a.x - f01.x, a.y - f01.y, a.z - f23.x, a.w - f23.y, fa * 2048.f);
    *reinterpret_cast<uint32_t*>(rowb + 64) = f8c_pack4(a.x, a.y, a.z, a.w, fa);
  }
}
static int pad_split_f8c64(const float* x, int B, int V, int C, Planes out, const float* alpha, cudaStream_t st) {
  if (C % 64) { set_error("pad_split_f8c64: C must be a multiple of 64"); return VXB_E_BADARG; }
  const long long total = (long long)B * (V + 2) * (V + 2) * (V + 2) * (C / 4);
  pad_split_f8c64_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(x, B, V, C, out.hi,
                                                                                                reinterpret_cast<uint8_t*>(out.lo), alpha);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
static __global__ void f8c_beta_kernel(const unsigned int* __restrict__ wmax, float* __restrict__ beta) {
  const float m = __uint_as_float(wmax[0]);
  int e = 0;
  if (m > 0.f && isfinite(m)) e = max(-60, min(60, ilogbf(240.f * 2048.f / m)));
  beta[0] = scalbnf(1.f, e);
}
// static weights [rows, cols] (cols % 64 == 0, ld = cols): hi = fp16(beta / 32 w), c8 blocks [64 x e4m3(2^-11 beta w) | 64 x e4m3(beta w_lo)]
static __global__ void __launch_bounds__(256)
split_rows_f8c_kernel(const float* __restrict__ x, long long rows, long long cols, __nv_bfloat16* __restrict__ hi,
                      uint8_t* __restrict__ c8, const float* __restrict__ beta, const uint8_t* __restrict__ perm) {
  const long long cg = cols / 4, total = rows * cg;
  const float fb = __ldg(beta);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int g = (int)(i % cg);
    const long long rs = perm ? (long long)perm[r >> 6] * 64 + (r & 63) : r;      // output row block j <- source row block perm[j]
    const float4 a = *reinterpret_cast<const float4*>(x + rs * cols + g * 4);
    const __nv_bfloat162 u01 = pl2_from_floats(a.x, a.y), u23 = pl2_from_floats(a.z, a.w);
    const float2 f01 = pl2_to_float2(u01), f23 = pl2_to_float2(u23);
    const float s16 = fb * (1.f / F8C_P16);
    const __nv_bfloat162 h01 = pl2_from_floats(f01.x * s16, f01.y * s16), h23 = pl2_from_floats(f23.x * s16, f23.y * s16);
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(hi + r * cols + g * 4) = hv;
    uint8_t* rowb = c8 + r * cols * 2 + (g >> 4) * 128 + (g & 15) * 4;
    *reinterpret_cast<uint32_t*>(rowb) = f8c_pack4(a.x, a.y, a.z, a.w, fb * (1.f / 2048.f));
    *reinterpret_cast<uint32_t*>(rowb + 64) = f8c_pack4(a.x - f01.x, a.y - f01.y, a.z - f23.x, a.w - f23.y, fb);
  }
}
int absmax_cols(const float* x, long long rows, int C, unsigned int* out, cudaStream_t st);
int upconv_f8c_prepare(const float* wfold, long long rows, long long cols, Planes out, float* beta_out, unsigned int* tmp,
                       cudaStream_t st, const uint8_t* perm) {
  if (perm && rows % 64) { set_error("upconv_f8c_prepare: a row-block permutation needs rows %% 64 == 0"); return VXB_E_BADARG; }
  if (cols % 64 || out.ld != cols) { set_error("upconv_f8c_prepare: cols must be a multiple of 64 with ld == cols"); return VXB_E_BADARG; }
  VXB_CUDA(cudaMemsetAsync(tmp, 0, sizeof(unsigned int), st));
  VXB_TRY(absmax_cols(wfold, rows * cols, 1, tmp, st));
  f8c_beta_kernel<<<1, 1, 0, st>>>(tmp, beta_out);
  VXB_LAUNCH_CHECK();
  split_rows_f8c_kernel<<<148 * 8, 256, 0, st>>>(wfold, rows, cols, out.hi, reinterpret_cast<uint8_t*>(out.lo), beta_out, perm);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ---- block sparsity of the folded up-convolution weights [P * 64][27 * 64] (P = s^3 phases): which (phase, tap) blocks hold a
// non-zero value (nz[phase], bit = tap), an order of the phases in which the phases of one N tile share their zero taps
// (greedy: seed = the remaining phase with the most taps, then the phases that grow the tile's union least), and the K-block
// mask of every N tile in that order (Params::kmask).  All on the device: nothing here depends on host-side knowledge of the
// up-sampling stencil, the masks are read off the weights.
static __global__ void __launch_bounds__(256)
upconv_block_nz_kernel(const float* __restrict__ wfold, uint32_t* __restrict__ nz) {
  const int nb = blockIdx.x, ph = blockIdx.y;
  const float* base = wfold + ((size_t)ph * 64) * (27 * 64) + (size_t)nb * 64;
  int any = 0;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) any |= base[(size_t)(i >> 6) * (27 * 64) + (i & 63)] != 0.f;
  any = __syncthreads_or(any);
  if (threadIdx.x == 0 && any) atomicOr(nz + ph, 1u << nb);
}
static __global__ void upconv_phase_order_kernel(const uint32_t* __restrict__ nz, int P, int per_tile, uint8_t* __restrict__ perm) {
  __shared__ uint8_t used[256];
  if (threadIdx.x) return;
  for (int i = 0; i < P; ++i) used[i] = 0;
  int n = 0;
  while (n < P) {
    int seed = -1, best = -1;
    for (int q = 0; q < P; ++q)
      if (!used[q] && __popc(nz[q]) > best) { best = __popc(nz[q]); seed = q; }
    used[seed] = 1; perm[n++] = (uint8_t)seed;
    uint32_t m = nz[seed];
    for (int k = 1; k < per_tile && n < P; ++k) {
      int c = -1, grow = 1 << 30, own = -1;
      for (int q = 0; q < P; ++q) {
        if (used[q]) continue;
        const int g = __popc(m | nz[q]), o = __popc(nz[q]);
        if (g < grow || (g == grow && o > own)) { grow = g; own = o; c = q; }
      }
      used[c] = 1; perm[n++] = (uint8_t)c;
      m |= nz[c];
    }
  }
}
static __global__ void upconv_tile_mask_kernel(const uint32_t* __restrict__ nz, const uint8_t* __restrict__ perm, int P, int per_tile,
                                               int n_tiles, uint32_t* __restrict__ kmask) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  uint32_t m = 0;
  for (int j = t * per_tile; j < min(P, (t + 1) * per_tile); ++j) m |= nz[perm ? perm[j] : j];
  kmask[t] = m ? m : 1u;            // a tile never skips everything: its accumulator must be written
}
int upconv_kmask_build(const float* wfold, int s, uint32_t* nz, uint8_t* perm, uint32_t* kmask_natural, uint32_t* kmask_perm,
                       cudaStream_t st) {
  const int P = s * s * s, per_tile = UPCONV_NT / 64, n_tiles = cdiv(P * 64, UPCONV_NT);
  if (P > 255 || n_tiles > UPCONV_MAX_TILES) { set_error("upconv_kmask_build: s=%d not supported", s); return VXB_E_UNSUPPORTED_SHAPE; }
  VXB_CUDA(cudaMemsetAsync(nz, 0, (size_t)P * sizeof(uint32_t), st));
  upconv_block_nz_kernel<<<dim3(27, P), 256, 0, st>>>(wfold, nz);
  VXB_LAUNCH_CHECK();
  upconv_phase_order_kernel<<<1, 32, 0, st>>>(nz, P, per_tile, perm);
  VXB_LAUNCH_CHECK();
  upconv_tile_mask_kernel<<<1, 256, 0, st>>>(nz, nullptr, P, per_tile, n_tiles, kmask_natural);
  VXB_LAUNCH_CHECK();
  upconv_tile_mask_kernel<<<1, 256, 0, st>>>(nz, perm, P, per_tile, n_tiles, kmask_perm);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// folded upsample-conv: low [B,S^3,Ci] fp32 -> out [B,(S*s)^3,64] fp32 and/or out_planes = hi/lo planes of the
// replicate-padded fine grid [B,(S*s+2)^3,64] (interior written here, halo by halo_fill); Wp planes of [s^3*64][27*Ci]
int upconv_f32(const float* low, const Planes& Wp, const float* bias, float* out, int B, int S, int Ci, int Co, int s,
               float act_slope, Arena& scratch, cudaStream_t st, const Planes* out_planes, const float* f8a,
               const F8cGemm* f8g, const UpconvSparsity* sp) {
  if (Ci % 64 || Co != 64) {
    set_error("umma upconv: needs Ci %% 64 == 0 and Co == 64");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  const long long Sp = S + 2;
  const long long rows = (long long)B * Sp * Sp * Sp;
  Planes a0 = alloc_planes(scratch, rows, Ci);
  if (!scratch.ok) {
    set_error("umma upconv: scratch too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  if (f8g) VXB_TRY(pad_split_f8c64(low, B, S, Ci, a0, f8g->alpha, st));
  else VXB_TRY(pad_split(low, B, S, 1, Ci, a0, st));
  Params p;
  params_init(p);
  if (f8g) { p.terms = 2; p.ep.alpha_dev = f8g->unscale; }
  const int N = s * s * s * 64;
  const int nt = UPCONV_NT;
  if (sp && sp->kmask && Ci == 64) { p.kmask = sp->kmask; p.ep.phase_perm = sp->phase_perm; }   // 27 K blocks = 27 taps
  // one batch entry per sample, restricted to the rows between the first and the last interior voxel (the two
  // all-halo z planes of the padded low-resolution grid are skipped)
  const long long sp3 = Sp * Sp * Sp;
  const long long row_lo = (Sp + 1) * Sp + 1, row_hi = sp3 - row_lo;
  p.batches = B; p.a_row_zb = (int)sp3; p.a_row_off = (int)row_lo;
  p.m_tiles = cdiv(row_hi - row_lo, BM);
  p.n_tiles = cdiv(N, nt);
  p.plan.taps = 3; p.plan.Vp = (int)Sp; p.plan.cpb = Ci / 64; p.plan.cb_src0 = Ci / 64;
  p.plan.num_kb = 27 * p.plan.cpb;
  p.ep.M = (int)(row_hi - row_lo); p.ep.N = N; p.ep.row_mode = ROWS_PHASE; p.ep.Vp = (int)Sp; p.ep.pad = 1;
  p.ep.phase_s = s;
  p.ep.bias = bias; p.ep.act_slope = act_slope;
  if (out_planes) {
    // fine voxel (d,h,w) -> row of the padded grid: the epilogue adds out_pad to every coordinate
    p.ep.out_Vp = S * s + 2; p.ep.out_pad = 1;
    p.ep.out_hi = out_planes->hi; p.ep.out_lo = out_planes->lo; p.ep.ldp = 64;
    p.ep.f8a = f8a;                       // non-null: out_planes->lo is the c8 plane of conv_f8c.cuh
  } else {
    p.ep.out_Vp = S * s; p.ep.out_pad = 0;
    p.ep.out_f32 = out; p.ep.ldc = 64;
  }
  Operand A0{a0, rows, Ci};
  Operand w{Wp, N, 27ll * Ci};
  VXB_TRY(gemm(A0, nullptr, w, nt, p, st));
  if (out_planes) VXB_TRY(halo_fill(*out_planes, B, S * s, 1, 64, st));
  return VXB_OK;
}



int linear_planes(const Planes& A, long long M, int K, const Planes& W, int N, const LinOut& o, cudaStream_t st) {
  Params p;
  params_init(p);
  const int nt = pick_ntile_mn(cdiv(M, (long long)BM), N);
  p.n_tiles = cdiv(N, nt);
  p.plan.num_kb = cdiv(K, BK);
  p.ep.N = N; p.ep.row_mode = ROWS_PLAIN;
  p.ep.bias = o.bias; p.ep.alpha = o.alpha; p.ep.act_slope = o.act_slope;
  p.terms = o.terms; p.ep.alpha_dev = o.alpha_dev;
  p.ep.residual = o.residual; p.ep.res_rows = o.res_rows > 0 ? o.res_rows : 1; p.ep.ldr = o.ldr;
  p.ep.out_f32 = o.out_f32; p.ep.ldc = o.ldc;
  if (o.out_planes) {
    p.ep.out_hi = o.out_planes->hi; p.ep.out_lo = o.out_planes->lo; p.ep.ldp = o.out_planes->ld;
    p.ep.transpose_planes = o.transposed;
  }
  if (o.geglu) {
    if (N % 128 || !o.out_planes || o.transposed || o.out_f32) {
      set_error("linear_planes: the GEGLU epilogue needs N %% 128 == 0 and a plane output only");
      return VXB_E_BADARG;
    }
    p.ep.mode = EPI_GEGLU;
  }
  if (o.transposed && o.batches > 1) {
    if (M % o.batches) {
      set_error("linear_planes: M=%lld not divisible by batches=%d", M, o.batches);
      return VXB_E_BADARG;
    }
    const long long rpb = M / o.batches;
    p.batches = o.batches;
    p.a_row_zb = (int)rpb;
    p.m_tiles = cdiv(rpb, BM);
    p.ep.M = (int)rpb;
    p.p_zb = (long long)N * o.out_planes->ld;
  } else {
    p.m_tiles = cdiv(M, BM);
    p.ep.M = (int)M;
  }
  Operand a{A, M, K}, w{W, N, K};
  return gemm(a, nullptr, w, nt, p, st);
}

int project_vt(const Planes& ctx, int B, int Nk, int K, const Planes& Wv, int inner, const Planes& vt, cudaStream_t st,
               int terms, const float* alpha_dev) {
  Params p;
  params_init(p);
  p.terms = terms; p.ep.alpha_dev = alpha_dev;
  const int nt = pick_ntile(Nk);
  p.m_tiles = cdiv(inner, BM);
  p.n_tiles = cdiv(Nk, nt);
  p.plan.num_kb = cdiv(K, BK);
  p.batches = B; p.Hz = 1;
  p.a_row_zb = 0;                 // the weights are shared by every batch
  p.w_row_zb = Nk;                // the context rows of batch b are the N operand
  p.p_zb = (long long)inner * vt.ld;
  p.ep.M = inner; p.ep.N = Nk; p.ep.row_mode = ROWS_PLAIN;
  p.ep.out_hi = vt.hi; p.ep.out_lo = vt.lo; p.ep.ldp = vt.ld;
  const Operand a{Wv, inner, K};
  const Operand w{ctx, (long long)B * Nk, K};
  return gemm(a, nullptr, w, nt, p, st);
}

int attention_planes(const Planes& Q, int q_batched, const Planes& K, const Planes& Vt, int B, int H, int Nq, int Nk,
                     int dh, float scale, float* rowmax, float* rowsum, const Planes& P, const Planes& O,
                     cudaStream_t st, const AttnDrop* drop) {
  if (dh != 64) {
    set_error("attention_planes: dim_head must be 64 (got %d)", dh);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  const long long nstat = (long long)B * H * Nq;
  rowstat_init_kernel<<<(int)std::min<long long>((nstat + 255) / 256, 148 * 8), 256, 0, st>>>(rowmax, rowsum, nstat);
  VXB_LAUNCH_CHECK();
  const Operand q{Q, (long long)(q_batched ? B : 1) * Nq, (long long)H * dh};
  const Operand k{K, (long long)B * Nk, (long long)H * dh};
  Params p;
  params_init(p);
  p.m_tiles = cdiv(Nq, BM);
  p.n_tiles = cdiv(Nk, 256);
  p.plan.num_kb = 1;
  p.batches = B * H; p.Hz = H;
  p.a_row_zb = q_batched ? Nq : 0; p.a_col_zh = dh;
  p.w_row_zb = Nk; p.w_col_zh = dh;
  p.rs_zb = (long long)H * Nq; p.rs_zh = Nq;
  p.ep.M = Nq; p.ep.N = Nk; p.ep.row_mode = ROWS_PLAIN;
  p.ep.alpha = scale * 1.4426950408889634f;   // scores in the log2 domain
  // (1) row max of the scores from the hi planes alone: a stabiliser, it cancels in p / sum(p)
  Params p1 = p;
  p1.terms = 1;
  p1.ep.mode = EPI_ROWMAX; p1.ep.row_stat = rowmax;
  VXB_TRY(gemm(q, nullptr, k, 256, p1, st));
  static int use_flash = -1;
  if (use_flash < 0) {
    const char* e = getenv("VXB_ATTN");
    use_flash = (e && !strcmp(e, "gemm")) ? 0 : 1;
  }
  if (use_flash) {
    // (2+3) fused: P never leaves the SM (flash_umma.cuh)
    FlashParams f;
    memset(&f, 0, sizeof(f));
    f.B = B; f.H = H; f.Nq = Nq; f.Nk = Nk; f.dh = dh; f.q_batched = q_batched;
    f.q_tiles = cdiv(Nq, 128); f.k_tiles = cdiv(Nk, FA_KT); f.items = B * H * f.q_tiles;
    f.alpha = scale * 1.4426950408889634f;
    f.rowmax = rowmax;
    f.out_hi = O.hi; f.out_lo = O.lo; f.ldo = O.ld;
    if (drop && drop->thresh) { f.drop_thresh = drop->thresh; f.drop_inv_keep = drop->inv_keep; f.drop_seed = drop->seed; f.drop_ld = drop->ld; }
    CUtensorMap maps[6];
    VXB_TRY(make_map(&maps[0], Q.hi, q.rows, q.cols, Q.ld, 128));
    VXB_TRY(make_map(&maps[1], Q.lo, q.rows, q.cols, Q.ld, 128));
    VXB_TRY(make_map(&maps[2], K.hi, k.rows, k.cols, K.ld, FA_KT));
    VXB_TRY(make_map(&maps[3], K.lo, k.rows, k.cols, K.ld, FA_KT));
    VXB_TRY(make_map(&maps[4], Vt.hi, (long long)B * H * dh, Nk, Vt.ld, 64));
    VXB_TRY(make_map(&maps[5], Vt.lo, (long long)B * H * dh, Nk, Vt.ld, 64));
    static bool attr_set = false;
    if (!attr_set) {
      VXB_CUDA(cudaFuncSetAttribute(flash_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      attr_set = true;
    }
    int dev = 0, sms = 148;
    VXB_CUDA(cudaGetDevice(&dev));
    VXB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ++g_umma_launches;
    flash_attn_kernel<<<std::min(f.items, sms), FA_THREADS, FA_SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], f);
    VXB_LAUNCH_CHECK();
    return VXB_OK;
  }
  if (drop && drop->thresh) {
    set_error("attention_planes: dropout needs the fused attention kernel (VXB_ATTN=gemm selects the three-GEMM form)");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  // (2) P = 2^(s - max) as planes, row sums
  Params p2 = p;
  p2.ep.mode = EPI_EXP; p2.ep.row_sub = rowmax; p2.ep.row_stat = rowsum;
  p2.ep.out_hi = P.hi; p2.ep.out_lo = P.lo; p2.ep.ldp = P.ld;
  p2.p_zb = (long long)H * Nq * P.ld; p2.p_zh = (long long)Nq * P.ld;
  VXB_TRY(gemm(q, nullptr, k, 256, p2, st));
  // (3) O = (P V) / rowsum
  Params p3;
  params_init(p3);
  p3.m_tiles = cdiv(Nq, BM);
  p3.n_tiles = 1;
  p3.plan.num_kb = cdiv(Nk, BK);
  p3.batches = B * H; p3.Hz = H;
  p3.a_row_zb = H * Nq; p3.a_row_zh = Nq;
  p3.w_row_zb = H * dh; p3.w_row_zh = dh;
  p3.rs_zb = (long long)H * Nq; p3.rs_zh = Nq;
  p3.ep.M = Nq; p3.ep.N = dh; p3.ep.row_mode = ROWS_PLAIN;
  p3.ep.row_div = rowsum;
  p3.ep.out_hi = O.hi; p3.ep.out_lo = O.lo; p3.ep.ldp = O.ld;
  p3.p_zb = (long long)Nq * O.ld; p3.p_zh = dh;
  const Operand pa{P, (long long)B * H * Nq, (long long)Nk};
  const Operand vt{Vt, (long long)B * H * dh, (long long)Nk};
  return gemm(pa, nullptr, vt, 64, p3, st);
}


// ---- fp32 wrapper of attention_planes for the per-op parity test (not used by the Q-network forward)
static __global__ void split_gather_kernel(const float* __restrict__ x, int ld, long long bs, int B, int rows, int cols,
                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp,
                                           int transposed) {
  // element (b, r, c) of x -> plane[(b*rows + r), c], or transposed plane[(b*cols + c), r]
  const long long total = (long long)B * rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)((i / cols) % rows);
    const int b = (int)(i / ((long long)cols * rows));
    const float f = x[b * bs + (long long)r * ld + c];
    const __nv_bfloat16 h = pl_from_float(f);
    const __nv_bfloat16 l = pl_from_float(f - pl_to_float(h));
    const long long o = transposed ? ((long long)b * cols + c) * ldp + r : ((long long)b * rows + r) * ldp + c;
    hi[o] = h; lo[o] = l;
  }
}
static __global__ void merge_planes_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                           long long ldp, int B, int rows, int cols, float* __restrict__ out, int ldo,
                                           long long obs) {
  const long long total = (long long)B * rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols);
    const int r = (int)((i / cols) % rows);
    const int b = (int)(i / ((long long)cols * rows));
    const long long o = ((long long)b * rows + r) * ldp + c;
    out[b * obs + (long long)r * ldo + c] = pl_to_float(hi[o]) + pl_to_float(lo[o]);
  }
}

struct AttnScratch { Planes q, k, vt, p, o; float *rmax, *rsum; };
static void carve_attn(Arena& a, int B, int H, int Nq, int Nk, int dh, AttnScratch& s) {
  s.q = alloc_planes(a, (long long)B * Nq, (long long)H * dh);
  s.k = alloc_planes(a, (long long)B * Nk, (long long)H * dh);
  s.vt = alloc_planes(a, (long long)B * H * dh, pad8(Nk));
  s.p = alloc_planes(a, (long long)B * H * Nq, pad8(Nk));
  s.o = alloc_planes(a, (long long)B * Nq, (long long)H * dh);
  s.rmax = a.get<float>((size_t)B * H * Nq);
  s.rsum = a.get<float>((size_t)B * H * Nq);
}
size_t attention_f32_scratch_bytes(int B, int H, int Nq, int Nk, int dh) {
  Arena a(nullptr, 0);
  AttnScratch s;
  carve_attn(a, B, H, Nq, Nk, dh, s);
  return a.off;
}
int attention_f32(const float* q, int ldq, long long qbs, const float* k, const float* v, int ldkv, long long kvbs,
                  float* out, int ldo, long long obs, int B, int H, int Nq, int Nk, int dh, float scale, Arena& scratch,
                  cudaStream_t st, const AttnDrop* drop) {
  AttnScratch s;
  carve_attn(scratch, B, H, Nq, Nk, dh, s);
  if (!scratch.ok) {
    set_error("umma attention: scratch too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  const int inner = H * dh;
  split_gather_kernel<<<148 * 4, 256, 0, st>>>(q, ldq, qbs, B, Nq, inner, s.q.hi, s.q.lo, s.q.ld, 0);
  split_gather_kernel<<<148 * 4, 256, 0, st>>>(k, ldkv, kvbs, B, Nk, inner, s.k.hi, s.k.lo, s.k.ld, 0);
  VXB_CUDA(cudaMemsetAsync(s.vt.hi, 0, plane_elems((long long)B * inner, s.vt.ld) * 2, st));
  VXB_CUDA(cudaMemsetAsync(s.vt.lo, 0, plane_elems((long long)B * inner, s.vt.ld) * 2, st));
  split_gather_kernel<<<148 * 4, 256, 0, st>>>(v, ldkv, kvbs, B, Nk, inner, s.vt.hi, s.vt.lo, s.vt.ld, 1);
  VXB_LAUNCH_CHECK();
  VXB_TRY(attention_planes(s.q, 1, s.k, s.vt, B, H, Nq, Nk, dh, scale, s.rmax, s.rsum, s.p, s.o, st, drop));
  merge_planes_kernel<<<148 * 4, 256, 0, st>>>(s.o.hi, s.o.lo, s.o.ld, B, Nq, inner, out, ldo, obs);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}


// ------------------------------------------------------------------------------------------ backward contractions
// max |x| of a strided matrix -> power-of-two scale that puts it near 2^12 (fp16 planes: max 65504, normals from 6e-5)
// (contiguous and 16-byte aligned: 128-bit loads, no index arithmetic; strided views: a warp per row -- the first version
// paid a 64-bit division and modulo per element, 18 ms of a training step over 82 launches)
static __global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, long long ld, long long rows, int cols, unsigned int* __restrict__ out) {
  float m = 0.f;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if (ld == cols && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long total = rows * cols, n4 = total >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    long long i = tid;
    for (; i + nthr < n4; i += 2 * nthr) {          // two independent loads in flight
      const float4 a = x4[i], b = x4[i + nthr];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      m = fmaxf(m, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
    }
    for (; i < n4; i += nthr) {
      const float4 a = x4[i];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
    for (long long j = (n4 << 2) + tid; j < total; j += nthr) m = fmaxf(m, fabsf(x[j]));
  } else {
    const int lane = threadIdx.x & 31;
    for (long long r = tid >> 5; r < rows; r += nthr >> 5) {
      const float* xr = x + r * ld;
      for (int c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(xr[c]));
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));     // non-negative floats order like their bit patterns
}
static __global__ void scale_from_amax_kernel(const unsigned int* __restrict__ amax, float* __restrict__ scale) {
  const float m = __uint_as_float(*amax);
  float s = 1.f;
  if (m > 0.f && m < INFINITY) {
    int e;
    frexpf(m, &e);                       // m = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.f, 12 - e);             // m * s in [2^11, 2^12)
  }
  *scale = s;
}
// device scalar scale for x (written to scale[0]); amax_tmp: one uint of scratch
static int operand_scale(const float* x, long long ld, long long rows, int cols, unsigned int* amax_tmp, float* scale,
                         cudaStream_t st) {
  VXB_CUDA(cudaMemsetAsync(amax_tmp, 0, sizeof(unsigned int), st));
  const long long total = rows * cols;
  amax_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 8), 256, 0, st>>>(x, ld, rows, cols, amax_tmp);
  scale_from_amax_kernel<<<1, 1, 0, st>>>(amax_tmp, scale);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// fp32 [rows, cols] (ldx) * scale -> planes [rows, ldp]
static __global__ void __launch_bounds__(256)
split_rows_scaled_kernel(const float* __restrict__ x, long long ldx, long long rows, int cols, const float* __restrict__ scale,
                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp) {
  const float sc = *scale;
  const long long groups_per_row = ldp / 8;
  const long long total = rows * groups_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / groups_per_row;
    const int c = (int)(i % groups_per_row) * 8;
    float f[8];
    const float* src = x + r * ldx + c;
#pragma unroll
    for (int t = 0; t < 8; ++t) f[t] = (c + t < cols) ? src[t] * sc : 0.f;
    uint4 h, l;
    split8(f, h, l);
    *reinterpret_cast<uint4*>(hi + r * ldp + c) = h;
    *reinterpret_cast<uint4*>(lo + r * ldp + c) = l;
  }
}
// fp32 x stored [R, Cc] (ldx) * scale -> TRANSPOSED planes [Cc, ldp] (ldp >= R, ldp % 8 == 0); 64 x 64 tiles through shared
// memory: 256-byte row reads, and a thread packs 8 consecutive output columns into one 128-bit store per plane (the first
// version moved 32 x 32 tiles with 2-byte stores: 11.5 ms of a training step over 114 launches)
static __global__ void __launch_bounds__(256)
split_transpose_scaled_kernel(const float* __restrict__ x, long long ldx, int R, int Cc, const float* __restrict__ scale,
                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ldp) {
  __shared__ float tile[64][65];
  const float sc = *scale;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int j = ty; j < 64; j += 4) {
    const int r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < R && c < Cc) ? x[(long long)r * ldx + c] * sc : 0.f;
  }
  __syncthreads();
  for (int id = threadIdx.x; id < 64 * 8; id += 256) {
    const int g = id & 7, j = id >> 3;
    const int c = c0 + j, r = r0 + g * 8;           // output row c, columns r .. r + 7
    if (c < Cc && r < ldp) {
      float f[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] = (r + t < R) ? tile[g * 8 + t][j] : 0.f;
      uint4 h, l;
      split8(f, h, l);
      *reinterpret_cast<uint4*>(hi + (long long)c * ldp + r) = h;
      *reinterpret_cast<uint4*>(lo + (long long)c * ldp + r) = l;
    }
  }
}
// C[m, n] = (accumulate ? C : 0) + tmp[m, n] / (sa * sw)   (tmp == C allowed when not accumulating)
// (nsplit > 1: tmp holds split-K partials [nsplit][M][ldt], summed here)
static __global__ void __launch_bounds__(256)
unscale_kernel(float* __restrict__ C, int ldc, const float* __restrict__ tmp, int ldt, long long M, int N,
               const float* __restrict__ sa, const float* __restrict__ sw, int accumulate, int nsplit, long long split_stride) {
  const float inv = 1.f / (*sa * *sw);
  const long long total = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N;
    const int n = (int)(i % N);
    float acc = tmp[m * ldt + n];
    for (int sp = 1; sp < nsplit; ++sp) acc += tmp[sp * split_stride + m * ldt + n];
    const float v = acc * inv;
    float* d = C + m * ldc + n;
    *d = accumulate ? *d + v : v;
  }
}
// Split-K for the weight-gradient shapes (M x N = a weight matrix, K = every token of the batch): dW of a 512 x 512 projection
// is 8 tiles of 128 x 256 over K = 32 768 -- one wave on 8 of 148 SMs, ~1 ms whatever the matrix; with the K range cut into
// `splits` batch entries (the engine's per-batch K offsets) the launch fills the chip and the partials are summed by unscale_kernel
static int pick_gemm_ksplits(int M, int N, int nt, int K) {
  const long long tiles = (long long)cdiv(M, BM) * cdiv(N, nt);
  const int nkb = cdiv(K, BK);
  if (tiles * 2 > 148 || nkb < 32) return 1;
  const long long s = std::min<long long>(std::min<long long>(148 / tiles, nkb / 8), 32);
  if (s <= 1) return 1;
  const int kc = cdiv(nkb, (int)s);
  return cdiv(nkb, kc);
}

size_t gemm_any_scratch_bytes(int M, int N, int K, bool accumulate) {
  Arena a(nullptr, 0);
  alloc_planes(a, M, pad8(K));
  alloc_planes(a, N, pad8(K));
  a.get<float>(64);
  if (accumulate) a.get<float>((size_t)M * N);
  const int ns = pick_gemm_ksplits(M, N, pick_ntile(N), K);
  if (ns > 1) a.get<float>((size_t)ns * M * N);
  return a.off;
}

static __global__ void set_one_kernel(float* p) { *p = 1.f; }
static __global__ void inv_prod_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) { *out = 1.f / (*a * *b); }

// ---- the two K = dim_head contractions of the attention backward (bwd_ops.cuh attention_bwd): for every (batch, head)
//   out[b, h] (fp32 [Nq, ldc]) = alpha * A[b, :, h dh : (h+1) dh] W[b, :, h dh : (h+1) dh]^T
// (probability logits scale q k^T with A = q, W = k; gA = go v^T with A = go scaled per tensor, W = v).  One batched launch of
// the GEMM engine on split-fp16 planes of A and W, exactly the score GEMM of attention_planes with the plain fp32 store
// epilogue; the FFMA form of these two products took 29 ms of a 350 ms training step.  abs_ == 0: A shared by all batches.
static __global__ void recip_kernel(const float* __restrict__ s, float* __restrict__ out) { *out = 1.f / *s; }
size_t attn_scores_scratch_bytes(int B, int H, int Nq, int Nk, int dh) {
  Arena a(nullptr, 0);
  alloc_planes(a, (long long)B * Nq, (long long)H * dh);
  alloc_planes(a, (long long)B * Nk, (long long)H * dh);
  a.get<float>(64);
  return a.off;
}
int attn_scores_f32(const float* A, int lda, long long abs_, const float* W, int ldw, long long wbs, float* out, long long ldc,
                    int B, int H, int Nq, int Nk, int dh, float alpha, bool a_dynamic, Arena& scratch, cudaStream_t st) {
  const bool shared = abs_ == 0;
  if (dh != BK || (!shared && abs_ != (long long)Nq * lda) || wbs != (long long)Nk * ldw || (ldc & 3) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return VXB_E_UNSUPPORTED_SHAPE;                 // the caller keeps its FFMA form
  const long long inner = (long long)H * dh, arows = shared ? Nq : (long long)B * Nq, wrows = (long long)B * Nk;
  Planes Ap = alloc_planes(scratch, arows, inner), Wp = alloc_planes(scratch, wrows, inner);
  float* sc = scratch.get<float>(64);               // [0] scale of A, [1] 1, [2] 1 / scale of A, [3] amax scratch
  if (!scratch.ok) return VXB_E_WORKSPACE_TOO_SMALL;
  if (a_dynamic) {
    VXB_TRY(operand_scale(A, lda, arows, (int)inner, reinterpret_cast<unsigned int*>(sc + 3), sc, st));
  } else {
    set_one_kernel<<<1, 1, 0, st>>>(sc);
  }
  set_one_kernel<<<1, 1, 0, st>>>(sc + 1);
  recip_kernel<<<1, 1, 0, st>>>(sc, sc + 2);
  split_rows_scaled_kernel<<<(int)std::min<long long>((arows * (Ap.ld / 8) + 255) / 256, 148 * 16), 256, 0, st>>>(
      A, lda, arows, (int)inner, sc, Ap.hi, Ap.lo, Ap.ld);
  split_rows_scaled_kernel<<<(int)std::min<long long>((wrows * (Wp.ld / 8) + 255) / 256, 148 * 16), 256, 0, st>>>(
      W, ldw, wrows, (int)inner, sc + 1, Wp.hi, Wp.lo, Wp.ld);
  VXB_LAUNCH_CHECK();
  const Operand a{Ap, arows, inner}, w{Wp, wrows, inner};
  Params p;
  params_init(p);
  p.m_tiles = cdiv(Nq, BM);
  p.n_tiles = cdiv(Nk, 256);
  p.plan.num_kb = 1;
  p.batches = B * H; p.Hz = H;
  p.a_row_zb = shared ? 0 : Nq; p.a_col_zh = dh;
  p.w_row_zb = Nk; p.w_col_zh = dh;
  p.ep.M = Nq; p.ep.N = Nk; p.ep.row_mode = ROWS_PLAIN;
  p.ep.alpha = alpha; p.ep.alpha_dev = sc + 2;
  p.ep.out_f32 = out; p.ep.ldc = ldc;
  p.c_zb = (long long)H * Nq * ldc; p.c_zh = (long long)Nq * ldc;
  return gemm(a, nullptr, w, 256, p, st);
}


int gemm_any_f32(const float* A, long long lda, bool a_trans, const float* W, long long ldw, bool w_trans, float* C, int ldc,
                 int M, int N, int K, bool accumulate, Arena& scratch, cudaStream_t st, bool w_dynamic_scale) {
  const long long Kp = pad8(K);
  Planes Ap = alloc_planes(scratch, M, Kp), Wp = alloc_planes(scratch, N, Kp);
  float* sc = scratch.get<float>(64);           // [0] sa, [1] sw, [2..3] amax temporaries
  float* tmp = accumulate ? scratch.get<float>((size_t)M * N) : C;
  if (!scratch.ok) return VXB_E_WORKSPACE_TOO_SMALL;
  int nsplit = pick_gemm_ksplits(M, N, pick_ntile(N), K);
  float* part = nullptr;
  if (nsplit > 1) {
    if (scratch.off + align_up((size_t)nsplit * M * N * sizeof(float), 256) <= scratch.cap) part = scratch.get<float>((size_t)nsplit * M * N);
    else nsplit = 1;                                // short scratch: one K range, as before
  }
  auto split_op = [&](const float* X, long long ldx, bool trans, int rows /*M or N*/, Planes P, float* scale, unsigned int* amax,
                      bool dynamic) -> int {
    // stored extent: [rows, K] (plain) or [K, rows] (transposed)
    if (dynamic) {
      VXB_TRY(operand_scale(X, ldx, trans ? K : rows, trans ? rows : K, amax, scale, st));
    } else {
      set_one_kernel<<<1, 1, 0, st>>>(scale);       // weights / forward activations: O(1), as in the inference path
    }
    if (!trans) {
      const long long total = (long long)rows * (P.ld / 8);
      split_rows_scaled_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(X, ldx, rows, K, scale, P.hi,
                                                                                                        P.lo, P.ld);
    } else {
      split_transpose_scaled_kernel<<<dim3(cdiv(rows, 64), cdiv(P.ld, 64)), 256, 0, st>>>(X, ldx, K, rows, scale, P.hi, P.lo, P.ld);
    }
    VXB_LAUNCH_CHECK();
    return VXB_OK;
  };
  VXB_TRY(split_op(A, lda, a_trans, M, Ap, sc, reinterpret_cast<unsigned int*>(sc + 2), true));
  VXB_TRY(split_op(W, ldw, w_trans, N, Wp, sc + 1, reinterpret_cast<unsigned int*>(sc + 3), w_dynamic_scale));
  Params p;
  params_init(p);
  const int nt = pick_ntile(N);
  p.m_tiles = cdiv(M, BM);
  p.n_tiles = cdiv(N, nt);
  p.plan.num_kb = cdiv(K, BK);
  p.ep.M = M; p.ep.N = N; p.ep.row_mode = ROWS_PLAIN;
  p.ep.out_f32 = tmp; p.ep.ldc = accumulate ? N : ldc;
  if (nsplit > 1) {
    const int kc = cdiv(p.plan.num_kb, nsplit);     // K blocks per split; the last split's excess columns read as zero (TMA)
    p.plan.num_kb = kc;
    p.batches = nsplit; p.Hz = nsplit;
    p.a_col_zh = kc * BK; p.w_col_zh = kc * BK;
    p.ep.out_f32 = part; p.ep.ldc = N;
    p.c_zh = (long long)M * N;
  }
  else {
    // one K range: the epilogue un-scales by the device scalar 1 / (sa sw) and, when accumulating, adds C in place (each
    // element is read and written by the same thread) -- no separate pass over the output
    inv_prod_kernel<<<1, 1, 0, st>>>(sc, sc + 1, sc + 4);
    p.ep.alpha_dev = sc + 4;
    p.ep.out_f32 = C; p.ep.ldc = ldc;
    if (accumulate) { p.ep.residual = C; p.ep.res_rows = M; p.ep.ldr = ldc; }
  }
  Operand a{Ap, M, K}, w{Wp, N, K};
  VXB_TRY(gemm(a, nullptr, w, nt, p, st));
  if (nsplit > 1) {
    const long long total = (long long)M * N;
    unscale_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(C, ldc, part, N, M, N, sc, sc + 1,
                                                                                           accumulate ? 1 : 0, nsplit, (long long)M * N);
  }
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// channels-last fp32 x [B, V^3, 64] * scale -> CHANNEL-MAJOR planes [64][ldp] over the flat wgrad grid
// (B, V+2, V+2, Vx), Vx = V+2 rounded up to a multiple of 8 (TMA needs 16-byte aligned K offsets, so the (dz, dy) taps
// become K shifts of whole x-rows and the dx taps are taken from `dxs`-shifted copies of the gradient).
// Element (b, pz, py, px) = padded-grid value at (pz, py, px - dxs): halo = nearest interior voxel (replicate = 1) or zero.
static __global__ void __launch_bounds__(256)
pad_transpose_split_kernel(const float* __restrict__ x, int B, int V, int Vx, int dxs, int replicate,
                           const float* __restrict__ scale, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                           int ctot, int choff, long long rows) {
  __shared__ float tile[64][65];
  __shared__ long long src[64];
  const float sc = *scale;
  const int Vp = V + 2;
  const long long r0 = (long long)blockIdx.x * 64;
  if (threadIdx.x < 64) {
    const long long r = r0 + threadIdx.x;
    long long s = -1;
    if (r < rows) {
      long long v = r;
      const int pw = (int)(v % Vx) - dxs; v /= Vx;
      const int ph = (int)(v % Vp); v /= Vp;
      const int pd = (int)(v % Vp);
      const int b = (int)(v / Vp);
      const int d = min(max(pd - 1, 0), V - 1), h = min(max(ph - 1, 0), V - 1), w = min(max(pw - 1, 0), V - 1);
      const bool interior = (d == pd - 1) && (h == ph - 1) && (w == pw - 1);
      if (interior || replicate) s = (((long long)b * V + d) * V + h) * V + w;
    }
    src[threadIdx.x] = s;
  }
  __syncthreads();
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  for (int j = rl; j < 64; j += 4) {
    const long long s = src[j];
    tile[j][c] = s >= 0 ? x[s * 64 + c] * sc : 0.f;
  }
  __syncthreads();
  // K-blocked output [K / 64][ctot channels][64]: this block of 64 K positions, channels choff .. choff + 63; a thread packs
  // 8 consecutive K positions of one channel into one 128-bit store per plane (the first version stored element by element)
  for (int id = threadIdx.x; id < 64 * 8; id += 256) {
    const int g = id & 7, ch = id >> 3;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (r0 + g * 8 + j < rows) ? tile[g * 8 + j][ch] : 0.f;
    uint4 h, l;
    split8(f, h, l);
    const long long o = ((long long)blockIdx.x * ctot + choff + ch) * 64 + g * 8;
    *reinterpret_cast<uint4*>(hi + o) = h;
    *reinterpret_cast<uint4*>(lo + o) = l;
  }
}
// The three x-shifted copies of the zero-padded gradient (dxs = -1, 0, +1 into channel blocks 0, 64, 128 of a 192-channel
// K-blocked plane pair) from ONE read of gz: copy dxs at flat row r is the base copy at row r - dxs (rows that would cross an
// x-row boundary land on halo columns of the base copy, which are zero), so a block loads rows r0 - 1 .. r0 + 64 once and
// writes three 64-row slices (three launches of pad_transpose_split_kernel read the 4 GB gradient three times).
static __global__ void __launch_bounds__(256)
pad_transpose_split3_kernel(const float* __restrict__ x, int B, int V, int Vx, const float* __restrict__ scale,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ctot, long long rows) {
  __shared__ float tile[66][65];
  __shared__ long long src[66];
  const float sc = *scale;
  const int Vp = V + 2;
  const long long r0 = (long long)blockIdx.x * 64;
  if (threadIdx.x < 66) {
    const long long r = r0 - 1 + threadIdx.x;
    long long s = -1;
    if (r >= 0 && r < rows) {
      long long v = r;
      const int pw = (int)(v % Vx); v /= Vx;
      const int ph = (int)(v % Vp); v /= Vp;
      const int pd = (int)(v % Vp);
      const int b = (int)(v / Vp);
      if (pd >= 1 && pd <= V && ph >= 1 && ph <= V && pw >= 1 && pw <= V && b < B)
        s = (((long long)b * V + (pd - 1)) * V + (ph - 1)) * V + (pw - 1);
    }
    src[threadIdx.x] = s;
  }
  __syncthreads();
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  for (int j = rl; j < 66; j += 4) {
    const long long s = src[j];
    tile[j][c] = s >= 0 ? x[s * 64 + c] * sc : 0.f;
  }
  __syncthreads();
  for (int id = threadIdx.x; id < 3 * 64 * 8; id += 256) {
    const int copy = id >> 9, rem = id & 511;       // copy i holds dxs = i - 1: out[rr] = base[rr - dxs] = tile[rr + 2 - i]
    const int g = rem & 7, ch = rem >> 3;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (r0 + g * 8 + j < rows) ? tile[g * 8 + j + 2 - copy][ch] : 0.f;
    uint4 h, l;
    split8(f, h, l);
    const long long o = ((long long)blockIdx.x * ctot + copy * 64 + ch) * 64 + g * 8;
    *reinterpret_cast<uint4*>(hi + o) = h;
    *reinterpret_cast<uint4*>(lo + o) = l;
  }
}
// dwt[(tap, ci)][co] = inv * sum_split partial[tap][split][ci][co]
static __global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, const float* __restrict__ sa, const float* __restrict__ sw,
                    float* __restrict__ dwt, int taps, int mn) {
  const float inv = 1.f / (*sa * *sw);
  const long long total = (long long)taps * mn;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / mn), e = (int)(i % mn);
    const float* p = partial + ((long long)t * nsplit) * mn + e;
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += p[(long long)s * mn];
    dwt[i] = acc * inv;
  }
}
static __global__ void max2_kernel(unsigned int* a, const unsigned int* b) { if (*b > *a) *a = *b; }

constexpr int kWgradSplits = 296;
static long long wgrad_rows(int B, int V, long long* Vx_out) {
  const long long Vp = V + 2, Vx = (Vp + 63) / 64 * 64;      // x-rows padded to whole K blocks: every tap shift is a multiple of 64
  if (Vx_out) *Vx_out = Vx;
  return (long long)B * Vp * Vp * Vx;
}
size_t conv3_wgrad_scratch_bytes(int B, int V) {
  Arena a(nullptr, 0);
  const long long ld = wgrad_rows(B, V, nullptr);
  alloc_planes(a, 128, ld);
  alloc_planes(a, 192, ld);
  a.get<float>((size_t)27 * kWgradSplits * 128 * 64);
  a.get<float>(64);
  return a.off;
}

int conv3_wgrad_f32(const float* x0, const float* x1, const float* gz, float* dwt, int B, int V, Arena& scratch, cudaStream_t st) {
  long long Vx;
  const long long Vp = V + 2, rows = wgrad_rows(B, V, &Vx), ld = rows;      // rows is a multiple of 64
  if (rows >= (1ll << 31)) {
    set_error("conv3_wgrad: grid too large");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  Planes xt = alloc_planes(scratch, 128, ld);
  const Planes gall = alloc_planes(scratch, 192, ld);                    // gradient shifted by dx = -1, 0, +1 along x: 3 x 64 rows
  float* partial = scratch.get<float>((size_t)27 * kWgradSplits * 128 * 64);
  float* sc = scratch.get<float>(64);          // [0] sx, [1] sg, [2..4] amax temporaries
  if (!scratch.ok) return VXB_E_WORKSPACE_TOO_SMALL;
  unsigned int* am = reinterpret_cast<unsigned int*>(sc + 2);
  const long long nvox = (long long)B * V * V * V;
  // one scale for both input halves (they share the A operand), one for the gradient
  VXB_CUDA(cudaMemsetAsync(am, 0, 3 * sizeof(unsigned int), st));
  const int ab = (int)std::min<long long>((nvox * 64 + 255) / 256, 148 * 8);
  amax_kernel<<<ab, 256, 0, st>>>(x0, 64, nvox, 64, am);
  amax_kernel<<<ab, 256, 0, st>>>(x1, 64, nvox, 64, am + 1);
  amax_kernel<<<ab, 256, 0, st>>>(gz, 64, nvox, 64, am + 2);
  max2_kernel<<<1, 1, 0, st>>>(am, am + 1);
  scale_from_amax_kernel<<<1, 1, 0, st>>>(am, sc);
  scale_from_amax_kernel<<<1, 1, 0, st>>>(am + 2, sc + 1);
  VXB_LAUNCH_CHECK();
  const int tb = (int)((ld + 63) / 64);
  pad_transpose_split_kernel<<<tb, 256, 0, st>>>(x0, B, V, (int)Vx, 0, 1, sc, xt.hi, xt.lo, 128, 0, rows);
  pad_transpose_split_kernel<<<tb, 256, 0, st>>>(x1, B, V, (int)Vx, 0, 1, sc, xt.hi, xt.lo, 128, 64, rows);
  // sum_r X[r + S + dx] G[r] = sum_r' X[r' + S] G[r' - dx]: the copy for tap dx holds the gradient moved by +dx
  pad_transpose_split3_kernel<<<tb, 256, 0, st>>>(gz, B, V, (int)Vx, sc + 1, gall.hi, gall.lo, 192, rows);
  VXB_LAUNCH_CHECK();
  const long long Kc = (rows + kWgradSplits - 1) / kWgradSplits;
  const long long Kcb = (Kc + BK - 1) / BK * BK;
  const int nsplit = (int)((rows + Kcb - 1) / Kcb);
  {
    // ONE launch: tile = (K chunk, tap); the 27 taps of a chunk run side by side, so x and g stream from HBM once (the first
    // version launched one split-K GEMM per tap and read both operands 27 times: 84 ms of a 480 ms step at B=16)
    Params p;
    params_init(p);
    p.m_tiles = 27; p.n_tiles = 1;
    p.plan.num_kb = (int)(Kcb / BK);
    p.batches = nsplit; p.Hz = nsplit;
    p.a_col_zh = (int)Kcb; p.w_col_zh = (int)Kcb;
    p.tap_m = 1;
    p.kblk_a = 128; p.kblk_w = 192;
    for (int tap = 0; tap < 27; ++tap) {
      const int dz = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
      p.tap_acol[tap] = (int)(((long long)dz * Vp + dy) * Vx);    // multiple of 8 elements; out-of-range columns read as zero
      p.tap_wrow[tap] = (dx + 1) * 64;
      // rotate the K walk by -shift so that, K block for K block, every tap reads the same A columns
      const long long nkb = Kcb / BK;
      long long rot = (-(long long)p.tap_acol[tap] / BK) % nkb;
      if (rot < 0) rot += nkb;
      p.tap_rot[tap] = (int)rot;
    }
    p.c_zh = 27ll * 128 * 64;
    p.ep.M = 27 * 128; p.ep.N = 64; p.ep.row_mode = ROWS_PLAIN;
    p.ep.out_f32 = partial; p.ep.ldc = 64;                        // partial[split][tap][ci][co]
    const long long nblk = rows / BK;
    Operand a{Planes{xt.hi, xt.lo, BK}, nblk * 128, BK}, w{Planes{gall.hi, gall.lo, BK}, nblk * 192, BK};
    VXB_TRY(gemm(a, nullptr, w, 64, p, st));
  }
  if (nsplit < kWgradSplits)      // unused split slots must not contribute
    VXB_CUDA(cudaMemsetAsync(partial + (size_t)nsplit * 27 * 128 * 64, 0, (size_t)(kWgradSplits - nsplit) * 27 * 128 * 64 * sizeof(float), st));
  wgrad_reduce_kernel<<<148 * 2, 256, 0, st>>>(partial, kWgradSplits, sc, sc + 1, dwt, 1, 27 * 128 * 64);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// gz [B, V^3, C] * scale -> planes of the zero-embedded grid [B, (V + 2e)^3, C] (e rings of zeros around the data)
static __global__ void __launch_bounds__(256)
embed_split_scaled_kernel(const float* __restrict__ x, int B, int V, int e, int C, const float* __restrict__ scale,
                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const float sc = *scale;
  const int Vp = V + 2 * e;
  const int cg = C / 8;
  const long long total = (long long)B * Vp * Vp * Vp * cg;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 8;
    long long v = i / cg;
    const int pw = (int)(v % Vp) - e; v /= Vp;
    const int ph = (int)(v % Vp) - e; v /= Vp;
    const int pd = (int)(v % Vp) - e;
    const int b = (int)(v / Vp);
    uint4 hh = make_uint4(0, 0, 0, 0), ll = make_uint4(0, 0, 0, 0);
    if ((unsigned)pd < (unsigned)V && (unsigned)ph < (unsigned)V && (unsigned)pw < (unsigned)V) {
      const float* src = x + ((((long long)b * V + pd) * V + ph) * V + pw) * C + c;
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 bb = *reinterpret_cast<const float4*>(src + 4);
      const float f[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, bb.x * sc, bb.y * sc, bb.z * sc, bb.w * sc};
      split8(f, hh, ll);
    }
    const long long o = (i / cg) * C + c;
    *reinterpret_cast<uint4*>(hi + o) = hh;
    *reinterpret_cast<uint4*>(lo + o) = ll;
  }
}

size_t conv_dgrad_scratch_bytes(int B, int V, int Cz, int Cx, int k) {
  Arena a(nullptr, 0);
  const long long Vp = V + 4 * (k / 2);
  alloc_planes(a, (long long)B * Vp * Vp * Vp, Cz);
  a.get<float>(64);
  a.get<__nv_bfloat16>(std::max<size_t>(conv3_weight_elems(64), (size_t)2 * 64 * k * k * k * Cz));
  a.get<float>(64);
  return a.off;
}

int conv_dgrad_f32(const float* gz, int Cz, const float* wd, int Cx, float* gxp, int B, int V, int k, float* scale_out,
                   Arena& scratch, cudaStream_t st) {
  const int pad = k / 2;
  const int Vo = V + 2 * pad;                 // output (padded-gradient) grid
  const long long Vp = Vo + 2 * pad;          // zero-embedded input grid = replicate-padded view of the zero ring
  const long long rows = (long long)B * Vp * Vp * Vp;
  const long long Kt = (long long)k * k * k * Cz;
  if (Cz % 64 || Cx % 64 || rows >= (1ll << 31)) {
    set_error("conv_dgrad: unsupported channels/size (Cz=%d Cx=%d)", Cz, Cx);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  Planes zp = alloc_planes(scratch, rows, Cz);
  unsigned int* amax = reinterpret_cast<unsigned int*>(scratch.get<float>(64));
  const bool fast = (k == 3 && Cz == 64 && Vp <= 180);
  __nv_bfloat16* wbuf = scratch.get<__nv_bfloat16>(std::max<size_t>(conv3_weight_elems(64), (size_t)2 * 64 * Kt));
  float* zero_bias = scratch.get<float>(64);
  if (!scratch.ok) return VXB_E_WORKSPACE_TOO_SMALL;
  VXB_TRY(operand_scale(gz, Cz, (long long)B * V * V * V, Cz, amax, scale_out, st));
  {
    const long long total = rows * (Cz / 8);
    embed_split_scaled_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(gz, B, V, 2 * pad, Cz, scale_out,
                                                                                                      zp.hi, zp.lo);
    VXB_LAUNCH_CHECK();
  }
  VXB_CUDA(cudaMemsetAsync(zero_bias, 0, 64 * sizeof(float), st));
  const size_t block_elems = (size_t)B * Vo * Vo * Vo * 64;
  for (int j = 0; j < Cx / 64; ++j) {
    const float* wj = wd + (size_t)j * 64 * Kt;          // rows [64 j, 64 j + 64) of the tap-major dgrad weights
    float* out = gxp + (size_t)j * block_elems;
    if (fast) {
      VXB_TRY(conv3_prepare_weights(wj, 64, wbuf, st));
      VXB_TRY(conv3_planes(zp, nullptr, 64, 0, wbuf, zero_bias, -1.f, out, B, Vo, st, nullptr));
    } else {
      Planes wp{wbuf, wbuf + (size_t)64 * Kt, Kt};
      VXB_TRY(split_rows(wj, Kt, 64, (int)Kt, wp, st));
      Params p;
      params_init(p);
      const long long vp3 = Vp * Vp * Vp;
      const long long row_lo = ((long long)pad * Vp + pad) * Vp + pad, row_hi = vp3 - row_lo;
      p.batches = B; p.a_row_zb = (int)vp3; p.a_row_off = (int)row_lo;
      p.m_tiles = cdiv(row_hi - row_lo, BM);
      p.n_tiles = 1;
      p.plan.taps = k; p.plan.Vp = (int)Vp; p.plan.cpb = Cz / 64; p.plan.cb_src0 = Cz / 64;
      p.plan.num_kb = k * k * k * p.plan.cpb;
      p.ep.M = (int)(row_hi - row_lo); p.ep.N = 64; p.ep.row_mode = ROWS_CONV_FLAT; p.ep.Vp = (int)Vp; p.ep.pad = pad;
      p.ep.out_padded = 0;
      p.ep.out_f32 = out; p.ep.ldc = 64;
      Operand A0{zp, rows, Cz};
      Operand w{wp, 64, Kt};
      VXB_TRY(gemm(A0, nullptr, w, 64, p, st));
    }
  }
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------ conv3 (conv_umma.cuh)
size_t conv3_weight_elems(int Cin) { return (size_t)(Cin / CV_KC) * 27 * 2 * 64 * CV_KC; }

static __global__ void conv3_weight_kernel(const float* __restrict__ w /*[64][27][Cin]*/, int Cin,
                                           __nv_bfloat16* __restrict__ wc /*[ncb][27][2][64][32]*/) {
  const long long total = (long long)(Cin / CV_KC) * 27 * 64 * CV_KC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % CV_KC);
    const int co = (int)((i / CV_KC) % 64);
    const int tap = (int)((i / (CV_KC * 64)) % 27);
    const int cb = (int)(i / ((long long)CV_KC * 64 * 27));
    const float f = w[((long long)co * 27 + tap) * Cin + cb * CV_KC + kc];
    const __nv_bfloat16 h = pl_from_float(f);
    const long long o = (((long long)(cb * 27 + tap) * 2) * 64 + co) * CV_KC + kc;
    wc[o] = h;
    wc[o + 64 * CV_KC] = pl_from_float(f - pl_to_float(h));
  }
}

int conv3_prepare_weights(const float* w_tapmajor, int Cin, __nv_bfloat16* wc, cudaStream_t st) {
  if (Cin % CV_KC) {
    set_error("conv3: input channels must be a multiple of %d", CV_KC);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  conv3_weight_kernel<<<148 * 4, 256, 0, st>>>(w_tapmajor, Cin, wc);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

template <int CL, bool PAIR>
static int conv3_launch(const CUtensorMap* maps, const ConvParams& p, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    VXB_CUDA(cudaFuncSetAttribute(conv3_umma_kernel<CL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VXB_CUDA(cudaGetDevice(&dev));
    VXB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int max_clusters = num_sms / CL;
  if (CL > 1) {
    cfg.gridDim = dim3(num_sms / CL * CL);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv3_umma_kernel<CL, PAIR>, &cfg) == cudaSuccess && n > 0) max_clusters = std::min(max_clusters, n);
    else cudaGetLastError();
  }
  const int groups = p.items / CL;
  cfg.gridDim = dim3(std::min(groups, max_clusters) * CL);
  VXB_CUDA(cudaLaunchKernelEx(&cfg, conv3_umma_kernel<CL, PAIR>, maps[0], maps[1], maps[2], maps[3], maps[4], p));
  return VXB_OK;
}

static int conv3_cluster() {
  static int cl = 0;
  if (!cl) {
    const char* e = getenv("VXB_CONV_CLUSTER");
    cl = e ? atoi(e) : 2;
    if (cl != 1 && cl != 2 && cl != 4) cl = 2;
  }
  return cl;
}
// z chunking: each chunk of lz output planes stages lz + 2 input planes; pick the chunk count that minimises
// rounds(per cluster) x (lz + 2), i.e. halo re-reads against load imbalance of the persistent schedule
static void conv3_plan(int B, int V, int cl, int& tiles, int& lz, int& zchunks) {
  const int Vp = V + 2;
  tiles = cdiv((long long)Vp * Vp, 128);
  const int clusters = std::max(1, (cl == 4 ? 132 : 148) / cl);
  const long long ncols = cdiv((long long)B * tiles, cl) * cl;
  long long best = -1;
  int best_lz = V;
  for (int zch = 1; zch <= std::max(1, V / 4); ++zch) {
    const int l = cdiv(V, zch);
    const int zc = cdiv(V, l);
    const long long rounds = cdiv(ncols * zc / cl, clusters);
    const long long cost = rounds * (l + 2);
    if (best < 0 || cost < best) { best = cost; best_lz = l; }
  }
  lz = best_lz;
  zchunks = cdiv(V, lz);
}

constexpr int kTailMergeSplits = 32;
size_t conv3_tail_partial_floats(int B, int V) {
  int tiles, lz, zchunks;
  conv3_plan(B, V, conv3_cluster(), tiles, lz, zchunks);
  return (size_t)B * (zchunks * tiles * 16 + kTailMergeSplits) * 6 * 64;   // chunk partials + the first merge level
}

// trans[b, v] = bias + sum_t ptap[b][t][clamp(v + offset_t)]   (replicate padding of the 3x3x3 stencil)
static __global__ void __launch_bounds__(256)
trans_gather_kernel(const float* __restrict__ ptap, const float* __restrict__ bias, float* __restrict__ y, int B, int V) {
  const size_t V3 = (size_t)V * V * V;
  const size_t total = (size_t)B * V3;
  const float bv = bias[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % V), yy = (int)((i / V) % V), z = (int)((i / ((size_t)V * V)) % V);
    const int b = (int)(i / V3);
    const float* pb = ptap + (size_t)b * 27 * V3;
    float acc = bv;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = min(max(z + dz, 0), V - 1);
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yc = min(max(yy + dy, 0), V - 1);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int xc = min(max(x + dx, 0), V - 1);
          const int tp = (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1);
          acc += __ldg(pb + (size_t)tp * V3 + ((size_t)zz * V + yc) * V + xc);
        }
      }
    }
    y[i] = acc;
  }
}

int conv3_planes(const Planes& x0, const Planes* x1, int C0, int C1, const __nv_bfloat16* wc, const float* bias,
                 float act_slope, float* out, int B, int V, cudaStream_t st, const ConvTail* tail) {
  const int Vp = V + 2;
  const long long rows = (long long)B * Vp * Vp * Vp;
  if (C0 % CV_KC || C1 % CV_KC || x0.ld != 64 || (x1 && x1->ld != 64) || C0 > 64 || C1 > 64 || rows >= (1ll << 31) || Vp > 180) {
    set_error("conv3_planes: unsupported geometry (C0=%d C1=%d V=%d)", C0, C1, V);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  const int cl = conv3_cluster();
  static int desc_mode = -1;
  if (desc_mode < 0) {
    const char* d = getenv("VXB_CONV_DESC_MODE");
    desc_mode = d ? atoi(d) : 0;   // measured on B200: the swizzle is a function of the absolute smem address, base offset 0
  }
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.V = V; p.Vp = Vp;
  p.ncb = (C0 + C1) / CV_KC; p.cb_src0 = C0 / CV_KC;
  conv3_plan(B, V, cl, p.tiles, p.lz, p.zchunks);
  const int cols = B * p.tiles;
  const int cols_pad = cdiv(cols, cl) * cl;
  p.items = cols_pad * p.zchunks;
  p.items_real = cols * p.zchunks;
  p.slab_rows = 128 + 2 * (Vp + 1);
  p.box_rows = (cdiv(p.slab_rows, 2) + 7) / 8 * 8;
  p.base_off_mode = desc_mode;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VXB_CONV_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; }
    p.debug_skip = dbg;
  }
  p.bias = bias; p.act_slope = act_slope; p.out = out;
  if (tail) {
    p.tail_w = tail->tail_w; p.ptap = tail->ptap; p.ss_partial = tail->ss_partial; p.out = nullptr;
    p.tail_w2 = tail->tail_w2; p.ptap2 = tail->ptap2;
  }
  CUtensorMap maps[5];
  VXB_TRY(make_map(&maps[0], x0.hi, rows, 64, 64, p.box_rows, CV_KC));
  VXB_TRY(make_map(&maps[1], x0.lo, rows, 64, 64, p.box_rows, CV_KC));
  const Planes& xb = x1 ? *x1 : x0;
  VXB_TRY(make_map(&maps[2], xb.hi, rows, 64, 64, p.box_rows, CV_KC));
  VXB_TRY(make_map(&maps[3], xb.lo, rows, 64, 64, p.box_rows, CV_KC));
  static int pair = -1;
  if (pair < 0) {
    const char* e = getenv("VXB_CONV_PAIR");
    pair = e ? atoi(e) : 0;   // measured: the kernel is limited by the chip's sustained tensor throughput either way
  }
  const bool use_pair = pair && cl == 2;
  VXB_TRY(make_map(&maps[4], wc, (long long)p.ncb * 27 * 128, CV_KC, CV_KC, use_pair ? 32 : 128 / cl, CV_KC));
  const size_t smem = (size_t)CV_SLABS * 4 * p.box_rows * 64 + (size_t)CV_WSTAGES * (use_pair ? CV_PAIR_WBYTES : CV_WBYTES) + 1024 +
                      ((tail && !use_pair) ? 1024 + 2 * CV_TAILW_BYTES : 0);   // + the tail weights as a tcgen05 B operand
  ++g_umma_launches;
  int rc;
  if (use_pair) {
    rc = conv3_launch<2, true>(maps, p, smem, st);
  } else {
    switch (cl) {
      case 1: rc = conv3_launch<1, false>(maps, p, smem, st); break;
      case 4: rc = conv3_launch<4, false>(maps, p, smem, st); break;
      default: rc = conv3_launch<2, false>(maps, p, smem, st); break;
    }
  }
  return rc;
}

// second half of the fused tail: gather the 27 tap products per voxel into q_trans, merge the ss_final partials
int conv3_tail_finish(const ConvTail& tail, int B, int V, cudaStream_t st) {
  int tiles, lz, zchunks;
  conv3_plan(B, V, conv3_cluster(), tiles, lz, zchunks);
  const size_t total = (size_t)B * V * V * V;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  trans_gather_kernel<<<blocks, 256, 0, st>>>(tail.ptap, tail.tail_b, tail.q_trans, B, V);
  VXB_LAUNCH_CHECK();
  if (tail.tail_w2) {
    trans_gather_kernel<<<blocks, 256, 0, st>>>(tail.ptap2, tail.tail_b2, tail.q_trans2, B, V);
    VXB_LAUNCH_CHECK();
  }
  const int chunks = zchunks * tiles * 16;
  // two-level merge: 32 slices of the chunk list per sample, then the slices
  float* level1 = tail.ss_partial + (size_t)B * chunks * 6 * 64;
  ss_merge_kernel<<<dim3(cdiv(64, 32), B, kTailMergeSplits), 256, 0, st>>>(tail.ss_partial, chunks, 64, nullptr, 0, nullptr, 0, level1, nullptr);
  VXB_LAUNCH_CHECK();
  ss_merge_kernel<<<dim3(cdiv(64, 32), B), 256, 0, st>>>(level1, kTailMergeSplits, 64, tail.ss, tail.ss_stride, tail.mx, tail.mx_stride, nullptr, nullptr);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------ f16 + fp8-corrected conv (conv_f8c.cuh)
// Device scalars f8s[F8S_COUNT] of one convolution call (all powers of two except the bounds):
//   [0] alpha of source 0, [1] alpha of source 1 (activation -> e4m3 scale of the A_hi8 half; the A_lo8 half uses 2^11 alpha)
//   [2] 2^-s (scale of the fp8 accumulator), [3] beta of source 0, [4] beta of source 1 (W_lo -> e4m3; W_hi8 uses 2^-11 beta)
//   [5], [6] the activation bounds the alphas were derived from (diagnostics)
size_t conv3_f8c_w16_elems(int Cin) { return (size_t)(Cin / CV_KC) * 9 * F8_WROWS * CV_KC; }
size_t conv3_f8c_w8_bytes(int Cin) { return (size_t)(Cin / CV_KC) * 9 * F8_WROWS * 64; }

// static half of the weights: fp16 W_hi, rows [cb][tap9][j = 0,1,2 <-> dz = +1,0,-1 <-> tap index dzc = 2,1,0][co], 32 channels each
static __global__ void f8c_w16_kernel(const float* __restrict__ w /*[64][27][Cin]*/, int Cin, __nv_bfloat16* __restrict__ w16) {
  const long long total = (long long)(Cin / CV_KC) * 9 * F8_WROWS * CV_KC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % CV_KC);
    const int co = (int)((i / CV_KC) % 64);
    const int j = (int)((i / (CV_KC * 64)) % 3);
    const int tap9 = (int)((i / (CV_KC * F8_WROWS)) % 9);
    const int cb = (int)(i / ((long long)CV_KC * F8_WROWS * 9));
    const int tap = (2 - j) * 9 + tap9;
    w16[i] = pl_from_float(w[((long long)co * 27 + tap) * Cin + cb * CV_KC + kc]);
  }
}
// max |w| per source (channels [0, C0) and [C0, Cin)) as float bits; out[2] zeroed by the caller
static __global__ void f8c_wmax_kernel(const float* __restrict__ w, long long rows, int Cin, int C0, unsigned int* __restrict__ out) {
  float m0 = 0.f, m1 = 0.f;
  const long long total = rows * Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float a = fabsf(w[i]);
    if ((int)(i % Cin) < C0) m0 = fmaxf(m0, a); else m1 = fmaxf(m1, a);
  }
  m0 = warp_max(m0); m1 = warp_max(m1);
  if ((threadIdx.x & 31) == 0) { atomicMax(out, __float_as_uint(m0)); atomicMax(out + 1, __float_as_uint(m1)); }
}
__device__ __forceinline__ uint8_t f8c_e4m3(float x) {
  unsigned short a;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(a) : "f"(0.f), "f"(x));
  return (uint8_t)(a & 0xff);
}
// per-call half of the weights: rows as above, 64 bytes each = [32 x e4m3(2^-11 beta W) | 32 x e4m3(beta W_lo)]
static __global__ void f8c_w8_kernel(const float* __restrict__ w, int Cin, int cb_src0, const float* __restrict__ f8s,
                                     uint8_t* __restrict__ w8) {
  const long long total = (long long)(Cin / CV_KC) * 9 * F8_WROWS * CV_KC;
  const float beta0 = f8s[3], beta1 = f8s[4];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kc = (int)(i % CV_KC);
    const long long row = i / CV_KC;
    const int co = (int)(row % 64);
    const int j = (int)((row / 64) % 3);
    const int tap9 = (int)((row / F8_WROWS) % 9);
    const int cb = (int)(row / (F8_WROWS * 9));
    const int tap = (2 - j) * 9 + tap9;
    const float f = w[((long long)co * 27 + tap) * Cin + cb * CV_KC + kc];
    const float lo = f - pl_to_float(pl_from_float(f));
    const float beta = cb >= cb_src0 ? beta1 : beta0;
    w8[row * 64 + kc] = f8c_e4m3(f * (beta * (1.f / 2048.f)));
    w8[row * 64 + 32 + kc] = f8c_e4m3(lo * beta);
  }
}

// per-column max |x| of a row-major [rows, C] fp32 matrix as float bits (out[C] zeroed by the caller): every thread keeps to
// one column phase (the grid stride is a multiple of C), 128-bit loads when the layout allows
static __global__ void __launch_bounds__(256)
absmax_cols_kernel(const float* __restrict__ x, long long total, int C, unsigned int* __restrict__ out, int vec4) {
  extern __shared__ unsigned int am_s[];
  for (int i = threadIdx.x; i < C; i += blockDim.x) am_s[i] = 0u;
  __syncthreads();
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const long long stride = (nthreads + C - 1) / C * C;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec4) {
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    const long long n4 = total >> 2;
    for (long long i = t0; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      m[0] = fmaxf(m[0], fabsf(v.x)); m[1] = fmaxf(m[1], fabsf(v.y));
      m[2] = fmaxf(m[2], fabsf(v.z)); m[3] = fmaxf(m[3], fabsf(v.w));
    }
    const int c0 = (int)((4 * t0) % C);
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicMax(&am_s[(c0 + k) % C], __float_as_uint(m[k]));
  } else {
    float m = 0.f;
    for (long long i = t0; i < total; i += stride) m = fmaxf(m, fabsf(__ldg(x + i)));
    atomicMax(&am_s[(int)(t0 % C)], __float_as_uint(m));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) if (am_s[i]) atomicMax(out + i, am_s[i]);
}
int absmax_cols(const float* x, long long rows, int C, unsigned int* out, cudaStream_t st) {
  if (C <= 0 || C > 4096) { set_error("absmax_cols: C=%d", C); return VXB_E_BADARG; }
  const long long total = rows * C;
  const int vec4 = (total % 4 == 0) && (((uintptr_t)x & 15) == 0);
  const int blocks = (int)std::min<long long>((total / (vec4 ? 4 : 1) + 255) / 256 + 1, 148 * 8);
  absmax_cols_kernel<<<blocks, 256, C * sizeof(unsigned int), st>>>(x, total, C, out, vec4);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

__device__ __forceinline__ int f8c_floor_log2(float x) {      // floor(log2 x) clamped to a range that keeps every product finite
  if (!(x > 0.f) || !isfinite(x)) return 0;
  return max(-60, min(60, ilogbf(x)));
}
// alpha of one source from a bound of |x|: the largest power of two with alpha * bound <= 240 (e4m3 saturates at 448)
__device__ __forceinline__ float f8c_alpha(float bound) { return bound > 0.f ? scalbnf(1.f, f8c_floor_log2(240.f / bound)) : 1.f; }

// bound of |act(W g + b)| for the 1x1 input_preprocess convolution from the per-channel maxima of the voxel grid
static __global__ void f8c_bound_ipp_kernel(const unsigned int* __restrict__ gmax /*[CIN]*/, const float* __restrict__ w /*[C][CIN]*/,
                                            const float* __restrict__ bias, int CIN, int C, float* __restrict__ f8s) {
  __shared__ unsigned int mx;
  if (threadIdx.x == 0) mx = 0u;
  __syncthreads();
  for (int o = threadIdx.x; o < C; o += blockDim.x) {
    float bnd = fabsf(bias[o]);
    for (int c = 0; c < CIN; ++c) bnd = fmaf(fabsf(w[o * CIN + c]), __uint_as_float(gmax[c]), bnd);
    atomicMax(&mx, __float_as_uint(bnd * 1.0001f));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float bnd = __uint_as_float(mx);
    f8s[5] = bnd;
    f8s[0] = f8c_alpha(bnd);
  }
}
// bound of the folded up-convolution output from the per-channel maxima of its low-resolution input and the per-(phase, co, ci)
// absolute tap sums of the folded weights; then the joint scales of the two sources of the final convolution:
//   beta_max(src) = largest power of two with beta * 2^-11 max|W_src| <= 240;  s = min_src log2(alpha_src beta_max(src));
//   beta_src = 2^s / alpha_src (<= beta_max: nothing saturates);  f8s[2] = 2^-s
// (a) one warp per weight row: bound of that output channel and phase, max over rows into *bmax (float bits, zeroed by the caller)
static __global__ void __launch_bounds__(256)
f8c_bound_rows_kernel(const unsigned int* __restrict__ lmax /*[Ci]*/, const float* __restrict__ S /*[rows][Ci]*/, int rows, int Ci,
                      const float* __restrict__ bias /*[64]*/, unsigned int* __restrict__ bmax) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float* sr = S + (size_t)r * Ci;
    float a = 0.f;
    for (int c = lane; c < Ci; c += 32) a = fmaf(sr[c], __uint_as_float(lmax[c]), a);
    a = warp_sum(a) + fabsf(bias[r & 63]);
    m = fmaxf(m, a);
  }
  if (lane == 0) atomicMax(bmax, __float_as_uint(m * 1.001f));
}
// (b) the scalars
static __global__ void f8c_bound_up_kernel(const unsigned int* __restrict__ lmax /*[Ci]*/, int Ci, const unsigned int* __restrict__ bmax,
                                           const unsigned int* __restrict__ wmax /*[2]*/, float* __restrict__ f8s,
                                           const float* __restrict__ up_beta /*beta of the folded up-conv weights, or null*/) {
  if (threadIdx.x == 0) {
    const float bnd = __uint_as_float(bmax[0]);
    f8s[6] = bnd;
    const float a0 = f8s[0], a1 = f8c_alpha(bnd);
    f8s[1] = a1;
    const int b0 = f8c_floor_log2(240.f * 2048.f / fmaxf(__uint_as_float(wmax[0]), 1e-30f));
    const int b1 = f8c_floor_log2(240.f * 2048.f / fmaxf(__uint_as_float(wmax[1]), 1e-30f));
    const int s = min(ilogbf(a0) + b0, ilogbf(a1) + b1);
    f8s[2] = scalbnf(1.f, -s);
    f8s[3] = scalbnf(1.f, s - ilogbf(a0));
    f8s[4] = scalbnf(1.f, s - ilogbf(a1));
    if (up_beta) {
      // operands of the folded up-convolution GEMM itself (terms == 2): alpha of `low` from its exact maximum
      float lmx = 0.f;
      for (int c = 0; c < Ci; ++c) lmx = fmaxf(lmx, __uint_as_float(lmax[c]));
      const float al = f8c_alpha(lmx * 1.0001f);
      f8s[8] = al;
      f8s[9] = 1.f / (al * up_beta[0]);
    }
  }
}
// per-op entry: exact maxima of the two sources (amax[0], amax[1]) instead of bounds
static __global__ void f8c_scales_exact_kernel(const unsigned int* __restrict__ amax, const unsigned int* __restrict__ wmax,
                                               int two_src, float* __restrict__ f8s) {
  const float a0 = f8c_alpha(__uint_as_float(amax[0]) * 1.0001f);
  const float a1 = two_src ? f8c_alpha(__uint_as_float(amax[1]) * 1.0001f) : a0;
  const int b0 = f8c_floor_log2(240.f * 2048.f / fmaxf(__uint_as_float(wmax[0]), 1e-30f));
  const int b1 = two_src ? f8c_floor_log2(240.f * 2048.f / fmaxf(__uint_as_float(wmax[1]), 1e-30f)) : b0;
  const int s = min(ilogbf(a0) + b0, ilogbf(a1) + b1);
  f8s[0] = a0; f8s[1] = a1;
  f8s[2] = scalbnf(1.f, -s);
  f8s[3] = scalbnf(1.f, s - ilogbf(a0));
  f8s[4] = scalbnf(1.f, s - ilogbf(a1));
  f8s[5] = __uint_as_float(amax[0]); f8s[6] = __uint_as_float(amax[two_src ? 1 : 0]);
}
// S[row][ci] = sum over the 27 taps of |wfold[row][tap][ci]|  (row = phase * 64 + co)
static __global__ void f8c_fold_abs_kernel(const float* __restrict__ wfold, long long rows, int Ci, float* __restrict__ S) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows * Ci; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Ci;
    const int c = (int)(i % Ci);
    float a = 0.f;
    for (int t = 0; t < 27; ++t) a += fabsf(wfold[(r * 27 + t) * Ci + c]);
    S[i] = a;
  }
}

int conv3_f8c_prepare(const float* w_tapmajor, int Cin, int C0, __nv_bfloat16* w16, unsigned int* wmax, cudaStream_t st) {
  if (Cin % CV_KC || C0 % CV_KC) { set_error("conv3_f8c: channels must be multiples of %d", CV_KC); return VXB_E_UNSUPPORTED_SHAPE; }
  f8c_w16_kernel<<<148 * 4, 256, 0, st>>>(w_tapmajor, Cin, w16);
  VXB_LAUNCH_CHECK();
  VXB_CUDA(cudaMemsetAsync(wmax, 0, 2 * sizeof(unsigned int), st));
  f8c_wmax_kernel<<<64, 256, 0, st>>>(w_tapmajor, 64ll * 27, Cin, C0, wmax);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
int conv3_f8c_fold_abs(const float* wfold, long long rows, int Ci, float* S, cudaStream_t st) {
  f8c_fold_abs_kernel<<<148 * 4, 256, 0, st>>>(wfold, rows, Ci, S);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
int conv3_f8c_bound_ipp(const float* grid, long long rows, int CIN, const float* w, const float* bias, int C, unsigned int* gmax,
                        float* f8s, cudaStream_t st) {
  VXB_CUDA(cudaMemsetAsync(gmax, 0, (size_t)CIN * sizeof(unsigned int), st));
  VXB_TRY(absmax_cols(grid, rows, CIN, gmax, st));
  f8c_bound_ipp_kernel<<<1, 64, 0, st>>>(gmax, w, bias, CIN, C, f8s);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
int conv3_f8c_bound_up(const float* low, long long rows, int Ci, const float* S, long long srows, const float* bias,
                       const unsigned int* wmax, unsigned int* lmax, float* f8s, cudaStream_t st, const float* up_beta) {
  if (Ci > 256) { set_error("conv3_f8c: Ci=%d", Ci); return VXB_E_UNSUPPORTED_SHAPE; }
  VXB_CUDA(cudaMemsetAsync(lmax, 0, (size_t)(Ci + 1) * sizeof(unsigned int), st));      // lmax[Ci] = max row bound
  VXB_TRY(absmax_cols(low, rows, Ci, lmax, st));
  f8c_bound_rows_kernel<<<(int)std::min<long long>((srows + 7) / 8, 148 * 8), 256, 0, st>>>(lmax, S, (int)srows, Ci, bias, lmax + Ci);
  VXB_LAUNCH_CHECK();
  f8c_bound_up_kernel<<<1, 32, 0, st>>>(lmax, Ci, lmax + Ci, wmax, f8s, up_beta);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}
int conv3_f8c_quantize_weights(const float* w_tapmajor, int Cin, int C0, const float* f8s, uint8_t* w8, cudaStream_t st) {
  f8c_w8_kernel<<<148 * 2, 256, 0, st>>>(w_tapmajor, Cin, C0 / CV_KC, f8s, w8);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

template <int CL>
static int conv3_f8c_launch(const CUtensorMap* maps, const ConvParams& p, size_t smem, cudaStream_t st) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    VXB_CUDA(cudaFuncSetAttribute(conv3_f8c_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VXB_CUDA(cudaGetDevice(&dev));
    VXB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int max_clusters = num_sms / CL;
  if (CL > 1) {
    cfg.gridDim = dim3(num_sms / CL * CL);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv3_f8c_kernel<CL>, &cfg) == cudaSuccess && n > 0) max_clusters = std::min(max_clusters, n);
    else cudaGetLastError();
  }
  const int groups = p.items / CL;
  cfg.gridDim = dim3(std::min(groups, max_clusters) * CL);
  VXB_CUDA(cudaLaunchKernelEx(&cfg, conv3_f8c_kernel<CL>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p));
  return VXB_OK;
}

// x0 / x1: Planes{hi, c8} of the replicate-padded grids (the `lo` member holds the c8 plane); w16 / w8: conv3_f8c_prepare /
// conv3_f8c_quantize_weights; f8s: the device scalars both were derived from
int conv3_f8c_planes(const Planes& x0, const Planes* x1, int C0, int C1, const __nv_bfloat16* w16, const uint8_t* w8, const float* f8s,
                     const float* bias, float act_slope, float* out, int B, int V, cudaStream_t st, const ConvTail* tail) {
  const int Vp = V + 2;
  const long long rows = (long long)B * Vp * Vp * Vp;
  if (C0 % CV_KC || C1 % CV_KC || x0.ld != 64 || (x1 && x1->ld != 64) || C0 > 64 || C1 > 64 || rows >= (1ll << 31) || Vp > 180) {
    set_error("conv3_f8c_planes: unsupported geometry (C0=%d C1=%d V=%d)", C0, C1, V);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  int cl = conv3_cluster();
  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.V = V; p.Vp = Vp;
  p.ncb = (C0 + C1) / CV_KC; p.cb_src0 = C0 / CV_KC;
  conv3_plan(B, V, cl, p.tiles, p.lz, p.zchunks);
  const int cols = B * p.tiles;
  const int cols_pad = cdiv(cols, cl) * cl;
  p.items = cols_pad * p.zchunks;
  p.items_real = cols * p.zchunks;
  p.slab_rows = 128 + 2 * (Vp + 1);
  p.box_rows = (cdiv(p.slab_rows, 2) + 7) / 8 * 8;
  p.bias = bias; p.act_slope = act_slope; p.out = out; p.f8s = f8s;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VXB_CONV_DEBUG_SKIP"); dbg = e ? atoi(e) : 0; }   // timing experiments only (wrong results)
    p.debug_skip = dbg;
  }
  if (tail) {
    p.tail_w = tail->tail_w; p.ptap = tail->ptap; p.ss_partial = tail->ss_partial; p.out = nullptr;
    p.tail_w2 = tail->tail_w2; p.ptap2 = tail->ptap2;
  }
  CUtensorMap maps[6];
  VXB_TRY(make_map(&maps[0], x0.hi, rows, 64, 64, p.box_rows, CV_KC));
  VXB_TRY(make_map(&maps[1], x0.lo, rows, 64, 64, p.box_rows, CV_KC));
  const Planes& xb = x1 ? *x1 : x0;
  VXB_TRY(make_map(&maps[2], xb.hi, rows, 64, 64, p.box_rows, CV_KC));
  VXB_TRY(make_map(&maps[3], xb.lo, rows, 64, 64, p.box_rows, CV_KC));
  const long long wrows = (long long)p.ncb * 9 * F8_WROWS;
  VXB_TRY(make_map(&maps[4], w16, wrows, CV_KC, CV_KC, F8_WCHUNK_ROWS, CV_KC));
  VXB_TRY(make_map(&maps[5], reinterpret_cast<const __nv_bfloat16*>(w8), wrows, CV_KC, CV_KC, F8_WCHUNK_ROWS, CV_KC));
  const size_t smem = (size_t)CV_SLABS * 4 * p.box_rows * 64 + (size_t)CV_WSTAGES * F8_WBYTES + 1024 + 1024 + 2 * CV_TAILW_BYTES;
  ++g_umma_launches;
  switch (cl) {
    case 1: return conv3_f8c_launch<1>(maps, p, smem, st);
    case 4: return conv3_f8c_launch<4>(maps, p, smem, st);
    default: return conv3_f8c_launch<2>(maps, p, smem, st);
  }
}

// fp32 [B,V,V,V,64] -> Planes{hi, c8} of the replicate-padded grid (per-op entry of the f8c convolution); alpha = f8s[src]
static __global__ void __launch_bounds__(256)
pad_split_c8_kernel(const float* __restrict__ x, int B, int V, __nv_bfloat16* __restrict__ hi, uint8_t* __restrict__ c8,
                    const float* __restrict__ alpha) {
  const int Vp = V + 2;
  const long long total = (long long)B * Vp * Vp * Vp * 16;
  const float fa = __ldg(alpha);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i & 15);
    long long v = i >> 4;
    const long long row = v;
    const int pw = (int)(v % Vp); v /= Vp;
    const int ph = (int)(v % Vp); v /= Vp;
    const int pd = (int)(v % Vp);
    const int b = (int)(v / Vp);
    const int d = min(max(pd - 1, 0), V - 1), h = min(max(ph - 1, 0), V - 1), w = min(max(pw - 1, 0), V - 1);
    const float4 a = *reinterpret_cast<const float4*>(x + ((((long long)b * V + d) * V + h) * V + w) * 64 + g * 4);
    const __nv_bfloat162 h01 = pl2_from_floats(a.x, a.y), h23 = pl2_from_floats(a.z, a.w);
    const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
    uint2 hv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(hi + row * 64 + g * 4) = hv;
    uint8_t* rowb = c8 + row * 128 + (g >> 3) * 64 + (g & 7) * 4;
    *reinterpret_cast<uint32_t*>(rowb) = pl_e4m3x4(a.x - f01.x, a.y - f01.y, a.z - f23.x, a.w - f23.y, fa * 2048.f);
    *reinterpret_cast<uint32_t*>(rowb + 32) = pl_e4m3x4(a.x, a.y, a.z, a.w, fa);
  }
}
size_t conv3_f8c_scratch_bytes(int B, int V) {
  const size_t prow = (size_t)B * (V + 2) * (V + 2) * (V + 2);
  return 2 * align_up(prow * 128, 256) + align_up(conv3_f8c_w16_elems(64) * 2, 256) + align_up(conv3_f8c_w8_bytes(64), 256) + 4096;
}
// per-op entry (tests): y = act(conv3(x) + bias), x fp32 [B,V,V,V,64], tap-major weights [64][27][64]
int conv3_f8c_f32(const float* x, const float* w_tapmajor, const float* bias, float act_slope, float* out, int B, int V,
                  Arena& scratch, cudaStream_t st) {
  const long long prow = (long long)B * (V + 2) * (V + 2) * (V + 2);
  Planes xp;
  xp.hi = scratch.get<__nv_bfloat16>((size_t)prow * 64);
  xp.lo = scratch.get<__nv_bfloat16>((size_t)prow * 64);
  xp.ld = 64;
  __nv_bfloat16* w16 = scratch.get<__nv_bfloat16>(conv3_f8c_w16_elems(64));
  uint8_t* w8 = scratch.get<uint8_t>(conv3_f8c_w8_bytes(64));
  float* f8s = scratch.get<float>(16);
  unsigned int* stat = scratch.get<unsigned int>(8);
  if (!scratch.ok) { set_error("conv3_f8c: workspace too small"); return VXB_E_WORKSPACE_TOO_SMALL; }
  VXB_TRY(conv3_f8c_prepare(w_tapmajor, 64, 64, w16, stat, st));                 // stat[0..1] = max |w|
  VXB_CUDA(cudaMemsetAsync(stat + 2, 0, 2 * sizeof(unsigned int), st));
  VXB_TRY(absmax_cols(x, (long long)B * V * V * V * 64, 1, stat + 2, st));       // stat[2] = max |x|
  f8c_scales_exact_kernel<<<1, 1, 0, st>>>(stat + 2, stat, 0, f8s);
  VXB_LAUNCH_CHECK();
  VXB_TRY(conv3_f8c_quantize_weights(w_tapmajor, 64, 64, f8s, w8, st));
  const long long total = prow * 16;
  pad_split_c8_kernel<<<(int)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(x, B, V, xp.hi, reinterpret_cast<uint8_t*>(xp.lo), f8s);
  VXB_LAUNCH_CHECK();
  return conv3_f8c_planes(xp, nullptr, 64, 0, w16, w8, f8s, bias, act_slope, out, B, V, st, nullptr);
}

// ------------------------------------------------------------------------------------------ patchify (patchify_umma.cuh)
size_t patchify_weight_elems(int k) { return (size_t)k * k * k * 2 * 64 * 64; }

static __global__ void patchify_weight_kernel(const float* __restrict__ w /*[64][k3][64]*/, int k3,
                                              __nv_bfloat16* __restrict__ wc /*[k3][2][64][64]*/) {
  const long long total = (long long)k3 * 64 * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % 64);
    const int co = (int)((i / 64) % 64);
    const int tap = (int)(i / 4096);
    const float f = w[((long long)co * k3 + tap) * 64 + ci];
    const __nv_bfloat16 h = pl_from_float(f);
    const long long o = ((long long)tap * 2 * 64 + co) * 64 + ci;
    wc[o] = h;
    wc[o + 64 * 64] = pl_from_float(f - pl_to_float(h));
  }
}

int patchify_prepare_weights(const float* w_tapmajor, int k, __nv_bfloat16* wc, cudaStream_t st) {
  patchify_weight_kernel<<<148 * 2, 256, 0, st>>>(w_tapmajor, k * k * k, wc);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

int patchify_f32(const float* x, const __nv_bfloat16* wc, const float* bias, float act_slope, float* out, int B, int V,
                 int k, int s, cudaStream_t st, const Planes* xplanes) {
  const int pad = k / 2;
  const int S = (V + 2 * pad - k) / s + 1;
  if (S <= 0 || (!xplanes && (!x || ((uintptr_t)x & 15))) || ((uintptr_t)out & 15) || ((uintptr_t)bias & 15) ||
      (xplanes && (xplanes->ld != 64 || V > 1023))) {      // plane gather packs voxel coordinates in 10 bits
    set_error("patchify: bad geometry or unaligned pointers");
    return VXB_E_BADARG;
  }
  PatchifyParams p;
  memset(&p, 0, sizeof(p));
  p.x = xplanes ? nullptr : x; p.bias = bias; p.out = out;
  if (xplanes) { p.xhi = xplanes->hi; p.xlo = xplanes->lo; }
  p.B = B; p.V = V; p.S = S; p.k = k; p.s = s; p.pad = pad;
  p.tokens = B * S * S * S;
  p.tiles = cdiv(p.tokens, 128);
  p.act_slope = act_slope;
  CUtensorMap map;
  VXB_TRY(make_map(&map, wc, (long long)k * k * k * 128, 64, 64, 128));
  static bool attr_set = false;
  if (!attr_set) {
    VXB_CUDA(cudaFuncSetAttribute(patchify_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    VXB_CUDA(cudaFuncSetAttribute(patchify_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PF_SMEM));
    attr_set = true;
  }
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    VXB_CUDA(cudaGetDevice(&dev));
    VXB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  ++g_umma_launches;
  if (xplanes) patchify_umma_kernel<true><<<std::min(p.tiles, num_sms), PF_THREADS, PF_SMEM, st>>>(map, p);
  else patchify_umma_kernel<false><<<std::min(p.tiles, num_sms), PF_THREADS, PF_SMEM, st>>>(map, p);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

}  // namespace umma
}  // namespace vxb

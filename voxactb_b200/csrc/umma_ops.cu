// Host side of the tcgen05 split-bf16 engine: TMA tensor maps, operand splitting / padding kernels
// and the launch wrappers used by dispatch.cuh.
#include "umma_gemm.cuh"
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
                    int box_rows) {
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
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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
    h[t] = __float2bfloat16_rn(f[t]);
    l[t] = __float2bfloat16_rn(f[t] - __bfloat162float(h[t]));
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

// fp32 channels-last [B, V, V, V, C] -> bf16 planes of the replicate-padded grid [B, Vp, Vp, Vp, C], Vp = V + 2*pad
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

// replicate-fill the halo of padded bf16 planes in place: every halo voxel copies its nearest interior voxel
static __global__ void __launch_bounds__(256)
halo_fill_kernel(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int B, int V, int pad, int C) {
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
    const int sd = min(max(pd, pad), V + pad - 1), sh = min(max(ph, pad), V + pad - 1), sw = min(max(pw, pad), V + pad - 1);
    if (sd == pd && sh == ph && sw == pw) continue;   // interior
    const long long so = ((((long long)b * Vp + sd) * Vp + sh) * Vp + sw) * C + c;
    const long long o = (i / cg) * C + c;
    *reinterpret_cast<uint4*>(hi + o) = *reinterpret_cast<const uint4*>(hi + so);
    *reinterpret_cast<uint4*>(lo + o) = *reinterpret_cast<const uint4*>(lo + so);
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

int halo_fill(Planes p, int B, int V, int pad, int C, cudaStream_t st) {
  const int Vp = V + 2 * pad;
  const long long total = (long long)B * Vp * Vp * Vp * (C / 8);
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  halo_fill_kernel<<<blocks, 256, 0, st>>>(p.hi, p.lo, B, V, pad, C);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------ launch
template <int NT, int STAGES>
static int launch_t(const CUtensorMap* maps, const Params& p, cudaStream_t st) {
  using L = SmemLayout<NT, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    VXB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
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
  umma_gemm_kernel<NT, STAGES><<<grid, THREADS, L::TOTAL, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
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
// (operands are split into bf16 planes in `scratch`; weights may come pre-split from vxb_qnet_prepare)
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
  p.m_tiles = cdiv(rows, BM);
  p.n_tiles = cdiv(Co, nt);
  p.plan.taps = k; p.plan.Vp = (int)Vp; p.plan.cpb = (C0 + C1) / 64; p.plan.cb_src0 = C0 / 64;
  p.plan.num_kb = k * k * k * p.plan.cpb;
  p.ep.M = (int)rows; p.ep.N = Co; p.ep.row_mode = ROWS_CONV_FLAT; p.ep.Vp = (int)Vp; p.ep.pad = pad;
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

// folded upsample-conv: low [B,S^3,Ci] fp32 -> out [B,(S*s)^3,64] fp32; Wp planes of [s^3*64][27*Ci]
int upconv_f32(const float* low, const Planes& Wp, const float* bias, float* out, int B, int S, int Ci, int Co, int s,
               float act_slope, Arena& scratch, cudaStream_t st) {
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
  VXB_TRY(pad_split(low, B, S, 1, Ci, a0, st));
  Params p;
  params_init(p);
  const int N = s * s * s * 64;
  const int nt = 256;
  p.m_tiles = cdiv(rows, BM);
  p.n_tiles = cdiv(N, nt);
  p.plan.taps = 3; p.plan.Vp = (int)Sp; p.plan.cpb = Ci / 64; p.plan.cb_src0 = Ci / 64;
  p.plan.num_kb = 27 * p.plan.cpb;
  p.ep.M = (int)rows; p.ep.N = N; p.ep.row_mode = ROWS_PHASE; p.ep.Vp = (int)Sp; p.ep.pad = 1;
  p.ep.phase_s = s; p.ep.out_Vp = S * s; p.ep.out_pad = 0;
  p.ep.bias = bias; p.ep.act_slope = act_slope;
  p.ep.out_f32 = out; p.ep.ldc = 64;
  Operand A0{a0, rows, Ci};
  Operand w{Wp, N, 27ll * Ci};
  return gemm(A0, nullptr, w, nt, p, st);
}

}  // namespace umma
}  // namespace vxb

// K1: point cloud -> dense voxel grid (scatter-mean), B200 native.
//
// Replaces VoxelGrid.coords_to_bounding_voxel_grid (reference peract/voxel/voxel_grid.py:148-198,
// with _scatter_nd :127-146 and _scatter_mean :106-125).
//
// Design (HBM-bound; algorithmic bytes/sample = N*(12+4F) read + V^3*(7+F)*4 write):
//   1. scatter: one thread per point computes the reference's clamped (V+2)-grid index with the
//      reference's exact fp32 operation order (sub, div, floor; no FMA contraction, no
//      reciprocal), drops points that fall in the cropped border, and accumulates
//      [count, xyz, features] into a per-sample open-addressing hash table that stays L2-resident
//      (2N slots x 36 B = 4.5 MB/sample at N=65536) -- so the dense (V+2)^3 x 7 accumulation buffer
//      of the reference (2 x 29.7 MB zero-fill + read-back per sample) never exists.  A point costs one
//      atomicCAS (slot claim) and two 128-bit vector reductions (red.global.add.v4.f32).  Points of one
//      warp that hit the same voxel are first combined with match_any + shuffles, which bounds the
//      contention of degenerate clouds (every point in one voxel); on the surface-heavy 4-camera input of
//      SURVEY.md section 8d neither that nor a shared-memory per-block bin removes traffic: 64 122 in-grid
//      points of a sample fall into 48 017 voxels, and the duplicates come from DIFFERENT cameras --
//      32-point groups hold 64 083 distinct (group, voxel) pairs, 1024-point blocks 62 885 (-2 %).
//   2. fill: one pass writes the dense [B,V,V,V,7+F] output, a warp per (x, y) row with 128-bit stores; the
//      94+ % empty voxels depend only on their position (index-grid channels from a per-block table of i / V,
//      zeros elsewhere) and are told apart by an occupancy bitmap (1 bit/voxel) without touching the hash
//      table (the first version staged 32 voxels per warp through shared memory with three fp32 divisions
//      and two integer divisions per voxel: 290 instructions per 32 voxels, issue-bound at 48 % of the HBM peak).
#include "common.cuh"
#include <algorithm>

namespace vxb {

struct VoxEntry {  // 32 bytes for F<=3; generic F uses stride_f floats
  int key;         // flat cropped voxel id + 1, 0 = empty
  float cnt;
  float sum[6];
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

// Reference index arithmetic, voxel_grid.py:152-163 (fp32, this exact operation order).
struct AxisMap {
  float shifted;  // bb_min - res
  float denom;    // res + 1e-12
};
__device__ __forceinline__ void make_axis_maps(const float* __restrict__ bnd, int V, AxisMap m[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float mn = bnd[a], mx = bnd[3 + a];
    float range = __fsub_rn(mx, mn);
    // dims_orig.float() + 1e-12 is evaluated in fp32 by torch -> exactly V
    float res = __fdiv_rn(range, __fadd_rn((float)V, 1e-12f));
    m[a].shifted = __fsub_rn(mn, res);
    m[a].denom = __fadd_rn(res, 1e-12f);
  }
}
__device__ __forceinline__ int axis_index(float p, const AxisMap& m, int V) {
  float q = __fdiv_rn(__fsub_rn(p, m.shifted), m.denom);
  float f = floorf(q);
  f = fminf(fmaxf(f, 0.f), (float)(V + 1));  // == min(.,V+1), max(.,0) on the int (voxel_grid.py:162-163)
  return (int)f;
}

// Raw-depth source: the world-frame point of pixel (x, y) of camera `cam` is formed here instead
// of on the host -- VisionSensor.pointcloud_from_depth_and_camera_params (reference PyRep/pyrep/objects/vision_sensor.py:
// 155-175 with _create_uniform_pixel_coords_image / _pixel_to_world_coords :381-412): pc = (x d, y d, d) in the depth dtype
// (fp32), then world = inv(K [R^T | -R^T C])[0:3] . (pc, 1) in float64, stored as fp32.  The 3x4 float64 matrix per
// (sample, camera) comes from the host (a 4x4 inverse per camera per step, as in the reference).
struct DepthSrc {
  const float* depth;       // [B, cams, H, W] metres
  const double* minv;       // [B, cams, 3, 4]
  const float* rgb;         // [B, cams, F, H, W] planar image features, or null when F == 0
  float* out_points;        // optional [B, N, 3] back-projected points (parity tests), N = cams * H * W
  int cams, H, W;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int F> struct VoxTab { static constexpr int EF = (1 + 3 + F + 3) / 4 * 4; };   // floats per entry: count, xyz, features

template <int F, bool DEPTH>
__global__ void __launch_bounds__(256)
vox_scatter_kernel(const float* __restrict__ coords, const float* __restrict__ feats,
                   const float* __restrict__ bounds, int Bb, int N, int V,
                   int* __restrict__ tkeys, float* __restrict__ tvals, int slots,
                   uint32_t* __restrict__ bitmap, int bitmap_words,
                   int32_t* __restrict__ out_idx, const DepthSrc ds) {
  const int b = blockIdx.y;
  __shared__ AxisMap maps[3];
  if (threadIdx.x == 0) make_axis_maps(bounds + (Bb == 1 ? 0 : b) * 6, V, maps);
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = n < N;
  float val[3 + F];
  int key = 0;
  if (in_range) {
    if constexpr (DEPTH) {
      const int hw = ds.H * ds.W;
      const int cam = n / hw, pix = n - cam * hw;
      const int y = pix / ds.W, x = pix - y * ds.W;
      const size_t img = (size_t)b * ds.cams + cam;
      const float d = ds.depth[img * hw + pix];
      const float xd = __fmul_rn((float)x, d), yd = __fmul_rn((float)y, d);     // upc * depth in fp32 (vision_sensor.py:165)
      const double* M = ds.minv + img * 12;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double w = fma(M[a * 4 + 2], (double)d, fma(M[a * 4 + 1], (double)yd, fma(M[a * 4 + 0], (double)xd, M[a * 4 + 3])));
        val[a] = (float)w;
      }
      if (ds.out_points) {
        float* o = ds.out_points + ((size_t)b * N + n) * 3;
        o[0] = val[0]; o[1] = val[1]; o[2] = val[2];
      }
#pragma unroll
      for (int f = 0; f < F; ++f) val[3 + f] = ds.rgb[(img * F + f) * hw + pix];
    } else {
    const float* c = coords + ((size_t)b * N + n) * 3;
    val[0] = c[0]; val[1] = c[1]; val[2] = c[2];
#pragma unroll
    for (int f = 0; f < F; ++f) val[3 + f] = feats[((size_t)b * N + n) * F + f];
    }
    int ix = axis_index(val[0], maps[0], V);
    int iy = axis_index(val[1], maps[1], V);
    int iz = axis_index(val[2], maps[2], V);
    if (out_idx) {
      int32_t* o = out_idx + ((size_t)b * N + n) * 3;
      o[0] = ix; o[1] = iy; o[2] = iz;
    }
    // border cells (index 0 or V+1) are cropped by vox[:, 1:-1, 1:-1, 1:-1] (voxel_grid.py:184)
    if (ix >= 1 && ix <= V && iy >= 1 && iy <= V && iz >= 1 && iz <= V)
      key = ((ix - 1) * V + (iy - 1)) * V + (iz - 1) + 1;
  }
  // ---- per-warp binning: combine lanes that target the same voxel
  const unsigned lane = threadIdx.x & 31u;
  unsigned peers = __match_any_sync(0xffffffffu, key);
  const unsigned leader = __ffs(peers) - 1;
  float cnt = 1.f;
  // warp-uniform loop over the lanes that are not the first of their voxel group; the group
  // leader adds their contribution (deterministic lane order), everybody else just feeds shuffles.
  unsigned followers = __ballot_sync(0xffffffffu, key != 0 && lane != leader);
  while (followers) {
    const int src = __ffs(followers) - 1;
    followers &= followers - 1;
    const bool take = (lane == leader) && ((peers >> src) & 1u);
#pragma unroll
    for (int j = 0; j < 3 + F; ++j) {
      float v = __shfl_sync(0xffffffffu, val[j], src);
      if (take) val[j] += v;
    }
    if (take) cnt += 1.f;
  }
  if (key == 0 || lane != leader) return;

  constexpr int EF = VoxTab<F>::EF;
  int* keys = tkeys + (size_t)b * slots;
  uint32_t h = hash_u32((uint32_t)key) & (uint32_t)(slots - 1);
  while (true) {
    int prev = atomicCAS(keys + h, 0, key);
    if (prev == 0 || prev == key) break;
    h = (h + 1) & (uint32_t)(slots - 1);
  }
  float* e = tvals + ((size_t)b * slots + h) * EF;
  red_add_v4(e, cnt, val[0], val[1], val[2]);
  if constexpr (F == 3) red_add_v4(e + 4, val[3], val[4], val[5], 0.f);
  static_assert(F == 0 || F == 3, "feature sizes compiled: 0 and 3");
  const int vid = key - 1;
  atomicOr(bitmap + (size_t)b * bitmap_words + (vid >> 5), 1u << (vid & 31));
}

// exact n / d for n < 2^32 with one 64-bit high multiply: m = ceil(2^64 / d)
struct FastDiv { unsigned long long m; unsigned int d; };
__device__ __forceinline__ unsigned int fdiv_u32(unsigned int n, const FastDiv& f) { return (unsigned int)__umul64hi((unsigned long long)n, f.m); }

// Dense writer, one warp per (x, y) row of V voxels = V * (7+F) contiguous floats, 128-bit stores (needs V * (7+F) % 4 == 0).
// Empty voxels (94+ %) depend only on their position: channels [0, 3+F) = 0, [3+F, 6+F) = voxel index / V (index_grid[:, :-2,
// :-2, :-2] / voxel_d, voxel_grid.py:197; exact fp32 division, tabulated once per block), 6+F = occupancy 0.  The occupancy
// bitmap (1 bit per voxel, L1-resident per row) tells them apart without touching the table; a lane whose float4 overlaps an
// occupied voxel looks its entry up (mean = sum / clamp(count, 1), voxel_grid.py:119; occupancy 1, :192).
template <int F>
__device__ __forceinline__ void vox_lookup(const int* __restrict__ keys, const float* __restrict__ vals, int slots, int key, float (&m)[3 + F]) {
  constexpr int EF = VoxTab<F>::EF;
  uint32_t h = hash_u32((uint32_t)key) & (uint32_t)(slots - 1);
#pragma unroll 1
  for (int probe = 0; probe < slots; ++probe) {
    const int kk = keys[h];
    if (kk == key) break;
    if (kk == 0) return;   // cannot happen for a set bit
    h = (h + 1) & (uint32_t)(slots - 1);
  }
  const float4* e = reinterpret_cast<const float4*>(vals + (size_t)h * EF);
  const float4 a = e[0];
  m[0] = a.y; m[1] = a.z; m[2] = a.w;
  if constexpr (F == 3) {
    const float4 c = e[1];
    m[3] = c.x; m[4] = c.y; m[5] = c.z;
  }
  if (a.x > 1.f) {          // three of four occupied voxels of the 4-camera input hold ONE point: mean == sum, no division
#pragma unroll
    for (int j = 0; j < 3 + F; ++j) m[j] = __fdiv_rn(m[j], a.x);
  }
}
template <int F>
__global__ void __launch_bounds__(256)
vox_fill_rows_kernel(const int* __restrict__ tkeys, const float* __restrict__ tvals, int slots,
                     const uint32_t* __restrict__ bitmap, int bitmap_words, int V, float* __restrict__ out) {
  constexpr int CH = 7 + F;
  extern __shared__ __align__(16) float vf_smem[];
  const int lut_floats = (V + 3) & ~3, row_floats = V * CH;      // row_floats % 4 == 0 (checked by the launcher)
  float* lut = vf_smem;
  for (int i = threadIdx.x; i < V; i += blockDim.x) lut[i] = __fdiv_rn((float)i, (float)V);
  __syncthreads();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rowbuf = vf_smem + lut_floats + warp * (row_floats + lut_floats);
  int* list = reinterpret_cast<int*>(rowbuf + row_floats);         // occupied voxels of the row (<= V entries)
  const int rows = V * V, row_f4 = row_floats / 4;
  const int* keys = tkeys + (size_t)b * slots;
  const float* vals = tvals + (size_t)b * slots * VoxTab<F>::EF;
  const uint32_t* bm = bitmap + (size_t)b * bitmap_words;
  int row = blockIdx.x * 8 + warp;
  // occupancy words of a row: bits [v0, v0 + V) of the bitmap, <= 32 words (launcher: V <= 960), one per lane
  auto load_bits = [&](int r) -> uint32_t {
    if (r >= rows) return 0u;
    const int v0 = r * V, w0 = v0 >> 5, nw = ((v0 + V - 1) >> 5) - w0 + 1;
    return lane < nw ? bm[w0 + lane] : 0u;
  };
  uint32_t wnext = load_bits(row);
  for (; row < rows; row += gridDim.x * 8) {
    uint32_t wbits = wnext;
    wnext = load_bits(row + gridDim.x * 8);                        // the next row's words are in flight while this row is built
    const int ix = row / V, iy = row - ix * V;
    const float fx = lut[ix], fy = lut[iy];
    const int v0 = row * V, w0 = v0 >> 5, nw = ((v0 + V - 1) >> 5) - w0 + 1;
    if (lane == 0) wbits &= ~0u << (v0 & 31);
    const int endbit = (v0 + V) - ((w0 + nw - 1) << 5);            // valid bits of the last word: 1..32
    if (lane == nw - 1 && endbit < 32) wbits &= (1u << endbit) - 1u;
    const int cnt = __popc(wbits);
    int pre = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if ((int)lane >= o) pre += t;
    }
    const int total = __shfl_sync(0xffffffffu, pre, 31);
    pre -= cnt;
    // background: the position-only values of every voxel of the row
    for (int iz = lane; iz < V; iz += 32) {
      float* r = rowbuf + iz * CH;
      float v[CH];
#pragma unroll
      for (int j = 0; j < 3 + F; ++j) v[j] = 0.f;
      v[3 + F] = fx; v[4 + F] = fy; v[5 + F] = lut[iz]; v[CH - 1] = 0.f;
      if constexpr (CH % 2 == 0) {
#pragma unroll
        for (int j = 0; j < CH; j += 2) *reinterpret_cast<float2*>(r + j) = make_float2(v[j], v[j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) r[j] = v[j];
      }
    }
    if (total) {                                                   // warp-uniform
      // occupied voxels of the row, compacted: ONE look-up round per 32 of them (a row of the 4-camera input holds ~5)
      uint32_t w = wbits;
      int pos = pre;
      while (w) {
        const int bpos = __ffs(w) - 1;
        w &= w - 1;
        list[pos++] = ((w0 + lane) << 5) + bpos - v0;
      }
      __syncwarp();
      for (int t = lane; t < total; t += 32) {
        const int iz = list[t];
        float m[3 + F];
#pragma unroll
        for (int j = 0; j < 3 + F; ++j) m[j] = 0.f;
        vox_lookup<F>(keys, vals, slots, v0 + iz + 1, m);
        float* r = rowbuf + iz * CH;
#pragma unroll
        for (int j = 0; j < 3 + F; ++j) r[j] = m[j];
        r[CH - 1] = 1.f;
      }
    }
    __syncwarp();
    // the row leaves with 128-bit stores
    float4* o = reinterpret_cast<float4*>(out + ((size_t)b * rows + row) * row_floats);
    const float4* src = reinterpret_cast<const float4*>(rowbuf);
    for (int q = lane; q < row_f4; q += 32) __stcs(o + q, src[q]);   // streaming stores: the 580 MB of output must not evict the table from L2
    __syncwarp();
  }
}

// generic geometry (V * (7+F) not a multiple of 4): scalar position pattern + per-slot patch
template <int CH>
__global__ void __launch_bounds__(256)
vox_background_scalar_kernel(float* __restrict__ out, int V, unsigned int V3) {
  float* o = out + (size_t)blockIdx.y * V3 * CH;
  for (unsigned int v = blockIdx.x * blockDim.x + threadIdx.x; v < V3; v += gridDim.x * blockDim.x) {
    const int ix = v / (V * V), iy = (v / V) % V, iz = v % V;
#pragma unroll
    for (int j = 0; j < CH; ++j) o[(size_t)v * CH + j] = 0.f;
    o[(size_t)v * CH + CH - 4] = __fdiv_rn((float)ix, (float)V);
    o[(size_t)v * CH + CH - 3] = __fdiv_rn((float)iy, (float)V);
    o[(size_t)v * CH + CH - 2] = __fdiv_rn((float)iz, (float)V);
  }
}

// One thread per table slot: mean = sum / clamp(count, 1) (voxel_grid.py:119), occupancy = 1 (voxel_grid.py:192)
template <int F>
__global__ void __launch_bounds__(256)
vox_occupied_kernel(const int* __restrict__ tkeys, const float* __restrict__ tvals, int slots, int V3, float* __restrict__ out) {
  constexpr int CH = 7 + F, EF = VoxTab<F>::EF;
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= slots) return;
  const int key = tkeys[(size_t)b * slots + s];
  if (key == 0) return;
  const float4* e = reinterpret_cast<const float4*>(tvals + ((size_t)b * slots + s) * EF);
  const float4 a = e[0];
  const float cnt = fmaxf(a.x, 1.f);
  float* o = out + ((size_t)b * V3 + (key - 1)) * CH;
  o[0] = __fdiv_rn(a.y, cnt); o[1] = __fdiv_rn(a.z, cnt); o[2] = __fdiv_rn(a.w, cnt);
  if constexpr (F == 3) {
    const float4 c = e[1];
    o[3] = __fdiv_rn(c.x, cnt); o[4] = __fdiv_rn(c.y, cnt); o[5] = __fdiv_rn(c.z, cnt);
  }
  o[CH - 1] = 1.f;
}

static int table_slots(int N) {
  int s = 1024;
  while (s < 2 * N) s <<= 1;
  return s;
}
static int entry_floats_for(int F) { return (1 + 3 + F + 3) / 4 * 4; }

}  // namespace vxb

using namespace vxb;

static size_t vox_bitmap_words(int V) { return align_up(((size_t)V * V * V + 31) / 32 + 1, 64); }

extern "C" size_t vxb_voxelize_workspace_bytes(int B, int N, int V, int F) {
  if (B <= 0 || N <= 0 || V <= 0 || F < 0) return 0;
  const size_t slots = (size_t)table_slots(N);
  return align_up((size_t)B * slots * sizeof(int), 256) + align_up((size_t)B * slots * entry_floats_for(F) * sizeof(float), 256) +
         align_up((size_t)B * vox_bitmap_words(V) * 4, 256);
}

extern "C" int vxb_voxelize_launches(void) { return 2; }

template <int F, bool DEPTH>
static int voxelize_impl(const float* coords, const float* feats, const float* bounds, int Bb,
                         int B, int N, int V, float* out, int32_t* out_idx, void* ws,
                         cudaStream_t st, const DepthSrc& ds) {
  constexpr int CH = 7 + F;
  const int slots = table_slots(N);
  const size_t key_bytes = align_up((size_t)B * slots * sizeof(int), 256);
  const size_t val_bytes = align_up((size_t)B * slots * entry_floats_for(F) * sizeof(float), 256);
  const int words = (int)vox_bitmap_words(V);
  int* tkeys = (int*)ws;
  float* tvals = (float*)((char*)ws + key_bytes);
  uint32_t* bitmap = (uint32_t*)((char*)ws + key_bytes + val_bytes);
  VXB_CUDA(cudaMemsetAsync(ws, 0, key_bytes + val_bytes + (size_t)B * words * 4, st));
  dim3 g1(cdiv(N, 256), B);
  vox_scatter_kernel<F, DEPTH><<<g1, 256, 0, st>>>(coords, feats, bounds, Bb, N, V, tkeys, tvals, slots, bitmap, words, out_idx, ds);
  VXB_LAUNCH_CHECK();
  const long long V3 = (long long)V * V * V;
  const size_t fill_smem = ((size_t)((V + 3) & ~3) * 9 + 8 * (size_t)V * CH) * sizeof(float);
  if ((V * CH) % 4 == 0 && (((uintptr_t)out) & 15) == 0 && fill_smem <= 200 * 1024 && V <= 960) {
    static size_t attr = 0;
    if (fill_smem > 48 * 1024 && fill_smem > attr) {
      VXB_CUDA(cudaFuncSetAttribute(vox_fill_rows_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = 200 * 1024;
    }
    // persistent blocks (B rows of them), 8 voxel rows per block and iteration
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(fill_smem, 1)));
    const int bx = (int)std::max<long long>(1, std::min<long long>(cdiv((long long)V * V, 8), cdiv(148 * per_sm, B)));
    vox_fill_rows_kernel<F><<<dim3(bx, B), 256, fill_smem, st>>>(tkeys, tvals, slots, bitmap, words, V, out);
    VXB_LAUNCH_CHECK();
  } else {
    // generic geometry: position pattern, then one thread per table slot patches the occupied voxels
    vox_background_scalar_kernel<CH><<<dim3((int)std::min<long long>(cdiv(V3, 256), 148 * 8), B), 256, 0, st>>>(out, V, (unsigned int)V3);
    VXB_LAUNCH_CHECK();
    vox_occupied_kernel<F><<<dim3(cdiv(slots, 256), B), 256, 0, st>>>(tkeys, tvals, slots, (int)V3, out);
    VXB_LAUNCH_CHECK();
  }
  return VXB_OK;
}

extern "C" int vxb_voxelize_f32(const float* coords, const float* feats, const float* bounds,
                                int Bb, int B, int N, int F, int V, float* out, int layout,
                                int32_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(coords && bounds && out && ws, "voxelize: null pointer argument");
  VXB_CHECK_ARG(B > 0 && N > 0 && V > 0, "voxelize: B, N, V must be positive");
  VXB_CHECK_ARG(Bb == 1 || Bb == B, "voxelize: bounds batch must be 1 or B (got %d, B=%d)", Bb, B);
  VXB_CHECK_ARG(layout == VXB_LAYOUT_CHANNELS_LAST, "voxelize: unknown layout %d", layout);
  VXB_CHECK_ARG(F == 0 || feats, "voxelize: feats is null but F=%d", F);
  if ((long long)V * V * V >= (1ll << 30)) {
    set_error("voxelize: V=%d too large", V);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (ws_bytes < vxb_voxelize_workspace_bytes(B, N, V, F)) {
    set_error("voxelize: workspace too small (%zu < %zu)", ws_bytes,
              vxb_voxelize_workspace_bytes(B, N, V, F));
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  DepthSrc none;
  memset(&none, 0, sizeof(none));
  switch (F) {
    case 0: return voxelize_impl<0, false>(coords, feats, bounds, Bb, B, N, V, out, out_idx, ws, st, none);
    case 3: return voxelize_impl<3, false>(coords, feats, bounds, Bb, B, N, V, out, out_idx, ws, st, none);
    default:
      set_error("voxelize: feature_size %d not compiled (0 or 3)", F);
      return VXB_E_UNSUPPORTED_SHAPE;
  }
}

extern "C" int vxb_voxelize_depth_f32(const float* depth, const double* proj_inv, const float* rgb, const float* bounds, int Bb,
                                      int B, int cams, int H, int W, int F, int V, float* out, int layout, float* out_points,
                                      int32_t* out_idx, void* ws, size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(depth && proj_inv && bounds && out && ws, "voxelize_depth: null pointer argument");
  VXB_CHECK_ARG(B > 0 && cams > 0 && H > 0 && W > 0 && V > 0, "voxelize_depth: sizes must be positive");
  VXB_CHECK_ARG((long long)cams * H * W < (1ll << 30), "voxelize_depth: too many pixels");
  VXB_CHECK_ARG(Bb == 1 || Bb == B, "voxelize_depth: bounds batch must be 1 or B (got %d, B=%d)", Bb, B);
  VXB_CHECK_ARG(layout == VXB_LAYOUT_CHANNELS_LAST, "voxelize_depth: unknown layout %d", layout);
  VXB_CHECK_ARG(F == 0 || rgb, "voxelize_depth: rgb is null but F=%d", F);
  const int N = cams * H * W;
  if ((long long)V * V * V >= (1ll << 30)) {
    set_error("voxelize_depth: V=%d too large", V);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (ws_bytes < vxb_voxelize_workspace_bytes(B, N, V, F)) {
    set_error("voxelize_depth: workspace too small (%zu < %zu)", ws_bytes, vxb_voxelize_workspace_bytes(B, N, V, F));
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  DepthSrc ds;
  ds.depth = depth; ds.minv = proj_inv; ds.rgb = rgb; ds.out_points = out_points; ds.cams = cams; ds.H = H; ds.W = W;
  cudaStream_t st = (cudaStream_t)stream;
  switch (F) {
    case 0: return voxelize_impl<0, true>(nullptr, nullptr, bounds, Bb, B, N, V, out, out_idx, ws, st, ds);
    case 3: return voxelize_impl<3, true>(nullptr, nullptr, bounds, Bb, B, N, V, out, out_idx, ws, st, ds);
    default:
      set_error("voxelize_depth: feature_size %d not compiled (0 or 3)", F);
      return VXB_E_UNSUPPORTED_SHAPE;
  }
}

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
//      (2N slots x 32 B = 4 MB/sample at N=65536; one 32-byte sector per entry, key included) -- so the dense (V+2)^3 x 7 accumulation buffer
//      of the reference (2 x 29.7 MB zero-fill + read-back per sample) never exists.  A point costs one
//      atomicCAS (slot claim) and three vector / scalar reductions on the same sector (red.global.add.v4 / .v2 / .f32).  Points of one
//      warp that hit the same voxel are first combined with match_any + shuffles, which bounds the
//      contention of degenerate clouds (every point in one voxel); on the surface-heavy 4-camera input of
//      SURVEY.md section 8d neither that nor a shared-memory per-block bin removes traffic: 64 122 in-grid
//      points of a sample fall into 48 017 voxels, and the duplicates come from DIFFERENT cameras --
//      32-point groups hold 64 083 distinct (group, voxel) pairs, 1024-point blocks 62 885 (-2 %).
//   2. fill: one pass writes the dense [B,V,V,V,7+F] output, a warp per (x, y) row with 128-bit stores; the
//      94+ % empty voxels depend only on their position (a template row per block, two channels patched in
//      registers) and are told apart by an occupancy bitmap (1 bit/voxel) without touching the hash table;
//      the look-ups of the occupied voxels are in flight while the warp streams the row's background out
//      (history: round 1 staged 32 voxels per warp through shared memory with three fp32 and two integer
//      divisions per voxel, 290 instructions per 32 voxels, issue-bound at 48 % of the HBM peak; the second
//      version built whole rows in per-warp buffers and was latency-bound at 40 warps / SM).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace vxb {

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

__device__ __forceinline__ void red_add_v2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
// One table entry = ONE 32-byte sector: [count, x, y, z | f0, f1, f2, key] (sums; key = flat cropped voxel id + 1 as an int,
// 0 = empty).  Claiming, accumulating and looking an entry up all touch that one sector (the first layout kept keys and
// values in two arrays: two sectors and two DRAM bursts per look-up).
// The table is re-read by the fill while 580 MB of output stream through L2: its accesses carry an evict_last policy
// (measured at B=16: 233 -> 219 us per call; a persisting-L2 window on the table was also tried and starves the stores: 420 us).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_v4_keep(float* addr, float a, float b, float c, float d, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_keep(const float4* addr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr), "l"(pol));
  return v;
}
constexpr int VOX_EF = 8, VOX_KEY = 7;

template <int F, bool DEPTH>
__global__ void __launch_bounds__(256)
vox_scatter_kernel(const float* __restrict__ coords, const float* __restrict__ feats,
                   const float* __restrict__ bounds, int Bb, int N, int V,
                   float* __restrict__ table, int slots,
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

  float* tab = table + (size_t)b * slots * VOX_EF;
  uint32_t h = hash_u32((uint32_t)key) & (uint32_t)(slots - 1);
  while (true) {
    int prev = atomicCAS(reinterpret_cast<int*>(tab + (size_t)h * VOX_EF) + VOX_KEY, 0, key);
    if (prev == 0 || prev == key) break;
    h = (h + 1) & (uint32_t)(slots - 1);
  }
  float* e = tab + (size_t)h * VOX_EF;
  red_add_v4_keep(e, cnt, val[0], val[1], val[2], l2_policy_evict_last());
  if constexpr (F == 3) {
    red_add_v2(e + 4, val[3], val[4]);
    atomicAdd(e + 6, val[5]);
  }
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
__device__ __forceinline__ void vox_entry_mean(const float4& a, const float4& c, float (&m)[3 + F]) {
  m[0] = a.y; m[1] = a.z; m[2] = a.w;
  if constexpr (F == 3) { m[3] = c.x; m[4] = c.y; m[5] = c.z; }
  if (a.x > 1.f) {          // three of four occupied voxels of the 4-camera input hold ONE point: mean == sum, no division
#pragma unroll
    for (int j = 0; j < 3 + F; ++j) m[j] = __fdiv_rn(m[j], a.x);
  }
}
template <int F>
__device__ __forceinline__ void vox_lookup(const float* __restrict__ tab, int slots, int key, float (&m)[3 + F]) {
  uint32_t h = hash_u32((uint32_t)key) & (uint32_t)(slots - 1);
#pragma unroll 1
  for (int probe = 0; probe < slots; ++probe) {
    const float4* e = reinterpret_cast<const float4*>(tab + (size_t)h * VOX_EF);
    const float4 c = e[1];
    const int kk = __float_as_int(c.w);
    if (kk == key) {
      vox_entry_mean<F>(e[0], c, m);
      return;
    }
    if (kk == 0) return;   // cannot happen for a set bit
    h = (h + 1) & (uint32_t)(slots - 1);
  }
}
// Fill, second design: NO per-warp row buffer.  The 94+ % background of a row differs from every other row only in two
// channels (x / V and y / V), so the block keeps ONE template row in shared memory (z / V and zeros) plus a byte per float4
// saying which of its elements are the x / y channels; a warp streams its row out as template + two register patches.  The
// look-ups of the row's occupied voxels (compacted from the occupancy bitmap, ~5 of 100) are issued BEFORE the background stores
// -- first-probe key and both value vectors speculatively, a mismatch falls back to the probing loop -- so the two dependent L2
// round trips of a look-up hide behind the warp's own stores instead of serialising with them; the occupied voxels are then
// overwritten in place (same warp, after __syncwarp: the sectors are still in L2 and merge there).  The first design built
// every row in a 4 KB per-warp buffer: ~13 k cycles of dependent steps per row at 40 warps / SM made the fill latency-bound
// (189 us at B=16, profiles/ncu_r02_voxelize_after_summary.json); without the buffer a block needs 8 KB and the SM holds 64 warps.
template <int F>
__global__ void __launch_bounds__(256, 6)       // 40 registers, 48 warps / SM (64 warps at 32 registers and 40 at 48 measured the same)
vox_fill_rows_kernel(const float* __restrict__ table, int slots,
                     const uint32_t* __restrict__ bitmap, int bitmap_words, int V, float* __restrict__ out) {
  constexpr int CH = 7 + F;
  extern __shared__ __align__(16) float vf_smem[];
  const int lut_floats = (V + 3) & ~3, row_floats = V * CH, row_f4 = row_floats / 4;   // row_floats % 4 == 0 (launcher)
  float* lut = vf_smem;                                            // i / V
  float* tmpl = lut + lut_floats;                                  // one background row with x = y = 0
  int* lists = reinterpret_cast<int*>(tmpl + row_floats);          // 8 x V: occupied voxels of each warp's row
  uint8_t* sel = reinterpret_cast<uint8_t*>(lists + 8 * V);        // per float4: (element of the x channel + 1) | (y ... + 1) << 4
  for (int i = threadIdx.x; i < V; i += blockDim.x) lut[i] = __fdiv_rn((float)i, (float)V);
  __syncthreads();
  for (int p = threadIdx.x; p < row_floats; p += blockDim.x) {
    const int iz = p / CH, ch = p - iz * CH;
    tmpl[p] = ch == 5 + F ? lut[iz] : 0.f;
  }
  // PAIR (F = 3): the x and y channels are floats 10 iz + 6, 7 -- an aligned pair that lands in .xy or .zw of one float4
  constexpr bool PAIR = CH % 4 == 2 && (3 + F) % 2 == 0;
  for (int q = threadIdx.x; q < row_f4; q += blockDim.x) {
    int s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = (4 * q + j) % CH;
      if constexpr (PAIR) {
        if (ch == 3 + F) s = j == 0 ? 1 : 2;
      } else {
        if (ch == 3 + F) s |= j + 1;
        if (ch == 4 + F) s |= (j + 1) << 4;
      }
    }
    sel[q] = (uint8_t)s;
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* list = lists + warp * V;
  const int rows = V * V;
  const float* tab = table + (size_t)b * slots * VOX_EF;
  const uint32_t* bm = bitmap + (size_t)b * bitmap_words;
  const float4* tmpl4 = reinterpret_cast<const float4*>(tmpl);
  int row = blockIdx.x * 8 + warp;
  // occupancy words of a row: bits [v0, v0 + V) of the bitmap, <= 32 words (launcher: V <= 960), one per lane
  auto load_bits = [&](int r) -> uint32_t {
    if (r >= rows) return 0u;
    const int v0 = r * V, w0 = v0 >> 5, nw = ((v0 + V - 1) >> 5) - w0 + 1;
    return lane < nw ? bm[w0 + lane] : 0u;
  };
  uint32_t wnext = load_bits(row);
  const int step = gridDim.x * 8, step_x = step / V, step_y = step - step_x * V;
  int ix = row / V, iy = row - ix * V;                             // advanced with the row: no division per row
  for (; row < rows; row += step) {
    uint32_t wbits = wnext;
    wnext = load_bits(row + gridDim.x * 8);                        // the next row's words are in flight while this row leaves
    const float fx = lut[ix], fy = lut[iy];
    ix += step_x; iy += step_y;
    if (iy >= V) { iy -= V; ++ix; }
    const int v0 = row * V, w0 = v0 >> 5, nw = ((v0 + V - 1) >> 5) - w0 + 1;
    if (lane == 0) wbits &= ~0u << (v0 & 31);
    const int endbit = (v0 + V) - ((w0 + nw - 1) << 5);            // valid bits of the last word: 1..32
    if (lane == nw - 1 && endbit < 32) wbits &= (1u << endbit) - 1u;
    int total = 0;
    // first look-up round, issued ahead of the stores
    int iz0 = -1;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), c0 = a0;
    if (__any_sync(0xffffffffu, wbits != 0u)) {
      const int cnt = __popc(wbits);
      int pre = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pre, o);
        if ((int)lane >= o) pre += t;
      }
      total = __shfl_sync(0xffffffffu, pre, 31);
      pre -= cnt;
      uint32_t w = wbits;
      int pos = pre;
      while (w) {
        const int bpos = __ffs(w) - 1;
        w &= w - 1;
        list[pos++] = ((w0 + lane) << 5) + bpos - v0;
      }
      __syncwarp();
      if (lane < total) {
        iz0 = list[lane];
        const uint32_t h0 = hash_u32((uint32_t)(v0 + iz0 + 1)) & (uint32_t)(slots - 1);
        const float4* e = reinterpret_cast<const float4*>(tab + (size_t)h0 * VOX_EF);
        const uint64_t keep = l2_policy_evict_last();
        a0 = ld_keep(e, keep);
        c0 = ld_keep(e + 1, keep);
      }
    }
    // background: template + this row's x / V and y / V, 128-bit streaming stores (the 580 MB of output must not evict the
    // table from L2)
    float* orow = out + ((size_t)b * rows + row) * row_floats;
    float4* o4 = reinterpret_cast<float4*>(orow);
    for (int q = lane; q < row_f4; q += 32) {
      float4 v = tmpl4[q];
      const int s = sel[q];
      if constexpr (PAIR) {
        if (s == 1) { v.x = fx; v.y = fy; }
        if (s == 2) { v.z = fx; v.w = fy; }
      } else {
        const int jx = s & 15, jy = s >> 4;
        if (jx == 1) v.x = fx; else if (jx == 2) v.y = fx; else if (jx == 3) v.z = fx; else if (jx == 4) v.w = fx;
        if (jy == 1) v.x = fy; else if (jy == 2) v.y = fy; else if (jy == 3) v.z = fy; else if (jy == 4) v.w = fy;
      }
      __stcs(o4 + q, v);
    }
    if (total) {                                                   // warp-uniform
      __syncwarp();                                                // orders the background stores before the overwrites below
      for (int t = lane; t < total; t += 32) {
        int iz;
        float m[3 + F];
        if (t == lane && __float_as_int(c0.w) == v0 + iz0 + 1) {   // first round, first probe hit (the common case)
          iz = iz0;
          vox_entry_mean<F>(a0, c0, m);
        } else {
          iz = list[t];
#pragma unroll
          for (int j = 0; j < 3 + F; ++j) m[j] = 0.f;
          vox_lookup<F>(tab, slots, v0 + iz + 1, m);
        }
        float v[CH];
#pragma unroll
        for (int j = 0; j < 3 + F; ++j) v[j] = m[j];
        v[3 + F] = fx; v[4 + F] = fy; v[5 + F] = lut[iz]; v[CH - 1] = 1.f;
        float* r = orow + iz * CH;
        if constexpr (CH % 2 == 0) {
#pragma unroll
          for (int j = 0; j < CH; j += 2) __stcs(reinterpret_cast<float2*>(r + j), make_float2(v[j], v[j + 1]));
        } else {
#pragma unroll
          for (int j = 0; j < CH; ++j) __stcs(r + j, v[j]);
        }
      }
      __syncwarp();                                                // the list is rebuilt by the next row
    }
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
vox_occupied_kernel(const float* __restrict__ table, int slots, int V3, float* __restrict__ out) {
  constexpr int CH = 7 + F;
  const int b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= slots) return;
  const float4* e = reinterpret_cast<const float4*>(table + ((size_t)b * slots + s) * VOX_EF);
  const float4 c = e[1];
  const int key = __float_as_int(c.w);
  if (key == 0) return;
  const float4 a = e[0];
  const float cnt = fmaxf(a.x, 1.f);
  float* o = out + ((size_t)b * V3 + (key - 1)) * CH;
  o[0] = __fdiv_rn(a.y, cnt); o[1] = __fdiv_rn(a.z, cnt); o[2] = __fdiv_rn(a.w, cnt);
  if constexpr (F == 3) {
    o[3] = __fdiv_rn(c.x, cnt); o[4] = __fdiv_rn(c.y, cnt); o[5] = __fdiv_rn(c.z, cnt);
  }
  o[CH - 1] = 1.f;
}

static int table_slots(int N) {
  int s = 1024;
  while (s < 2 * N) s <<= 1;
  return s;
}

}  // namespace vxb

using namespace vxb;

static size_t vox_bitmap_words(int V) { return align_up(((size_t)V * V * V + 31) / 32 + 1, 64); }

extern "C" size_t vxb_voxelize_workspace_bytes(int B, int N, int V, int F) {
  if (B <= 0 || N <= 0 || V <= 0 || F < 0) return 0;
  const size_t slots = (size_t)table_slots(N);
  return align_up((size_t)B * slots * VOX_EF * sizeof(float), 256) + align_up((size_t)B * vox_bitmap_words(V) * 4, 256);
}

extern "C" int vxb_voxelize_launches(void) { return 2; }

template <int F, bool DEPTH>
static int voxelize_impl(const float* coords, const float* feats, const float* bounds, int Bb,
                         int B, int N, int V, float* out, int32_t* out_idx, void* ws,
                         cudaStream_t st, const DepthSrc& ds) {
  constexpr int CH = 7 + F;
  const int slots = table_slots(N);
  const size_t tab_bytes = align_up((size_t)B * slots * VOX_EF * sizeof(float), 256);
  const int words = (int)vox_bitmap_words(V);
  float* table = (float*)ws;
  uint32_t* bitmap = (uint32_t*)((char*)ws + tab_bytes);
  VXB_CUDA(cudaMemsetAsync(ws, 0, tab_bytes + (size_t)B * words * 4, st));
  dim3 g1(cdiv(N, 256), B);
  vox_scatter_kernel<F, DEPTH><<<g1, 256, 0, st>>>(coords, feats, bounds, Bb, N, V, table, slots, bitmap, words, out_idx, ds);
  VXB_LAUNCH_CHECK();
  const long long V3 = (long long)V * V * V;
  // i / V table + template row + 8 occupied-voxel lists + one selector byte per float4 of the row
  const size_t fill_smem = align_up(((size_t)((V + 3) & ~3) + (size_t)V * CH + 8 * (size_t)V) * sizeof(float) + (size_t)V * CH / 4, 16);
  if ((V * CH) % 4 == 0 && (((uintptr_t)out) & 15) == 0 && fill_smem <= 200 * 1024 && V <= 960) {
    auto kern = vox_fill_rows_kernel<F>;
    static size_t attr = 0;
    if (fill_smem > 48 * 1024 && fill_smem > attr) {
      VXB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = 200 * 1024;
    }
    // persistent blocks (B rows of them), 8 voxel rows per block and iteration
    static int per_sm = 0;
    static size_t per_sm_smem = (size_t)-1;
    if (per_sm_smem != fill_smem) {
      VXB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, fill_smem));
      per_sm = std::max(1, per_sm);
      per_sm_smem = fill_smem;
    }
    const int bx = (int)std::max<long long>(1, std::min<long long>(cdiv((long long)V * V, 8), cdiv(148 * per_sm, B)));
    kern<<<dim3(bx, B), 256, fill_smem, st>>>(table, slots, bitmap, words, V, out);
    VXB_LAUNCH_CHECK();
  } else {
    // generic geometry: position pattern, then one thread per table slot patches the occupied voxels
    vox_background_scalar_kernel<CH><<<dim3((int)std::min<long long>(cdiv(V3, 256), 148 * 8), B), 256, 0, st>>>(out, V, (unsigned int)V3);
    VXB_LAUNCH_CHECK();
    vox_occupied_kernel<F><<<dim3(cdiv(slots, 256), B), 256, 0, st>>>(table, slots, (int)V3, out);
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

// Input-stationary 3x3x3 convolution on tcgen05, "f16 + fp8-corrected" arithmetic (sm_100a).
//
// Same data flow as conv_umma.cuh (one slab per (z plane, 32-channel block) staged once by TMA, the nine in-plane taps as
// row-shifted shared-memory descriptors, the CTA marching along z with three live output planes in tensor memory, the
// trans_decoder / ss_final tail fused into the epilogue), but a different split of the fp32 product.  With x = x_hi + x_lo
// (x_hi = fp16(x)) the three-term product  A_hi W_hi + A_hi W_lo + A_lo W_hi  of conv_umma.cuh spends two of its three
// fp16 MMAs on terms that are 2^-11 of the result.  Those two correction terms only need ~4 significant bits, so here they
// run as ONE kind::f8f6f4 (E4M3) MMA at twice the fp16 rate, K-concatenated:
//
//     F += A_hi . W_hi^T                                  (kind::f16,    K = 16 channels per instruction)
//     E += [A_lo8 | A_hi8] . [W_hi8 | W_lo8]^T            (kind::f8f6f4, K = 32 = 16 channels of both terms)
//     out = F + 2^-s E
//
// A_hi8 = e4m3(alpha A), A_lo8 = e4m3(2^11 alpha A_lo), W_lo8 = e4m3(beta W_lo), W_hi8 = e4m3(2^-11 beta W) with power-of-two
// scales, alpha beta = 2^s for every source tensor (alpha from a device-side bound of the activation, beta re-derived per
// call, see f8c_* in umma_ops.cu) -- E carries one absolute scale.  Measured on the reference goldens (tools/sim_two_term.py
// and tools/report_errors.py): the fp8 rounding of the correction terms adds 2e-5 .. 5e-5 to the Q-value error, against
// 3e-4 .. 1.5e-3 when either correction term is dropped (the "2-term" variants, rejected).
// 2 MMA units per product instead of 3; the activation planes keep their size (the fp16 lo plane is replaced by the c8
// plane: per 32-channel block 32 bytes of A_lo8 followed by 32 bytes of A_hi8).
//
// The three output planes an input plane feeds (dz = +1, 0, -1) share the A operand, so they are ONE MMA with N = 192
// (weights of the three dz taps stacked along N, accumulators of consecutive output planes in adjacent TMEM columns):
// the A tile is read from shared memory once per 192 columns instead of once per 64/128 -- the three-term kernel was
// reading 149 B/clk of operands at full MMA rate against the 128 B/clk the shared-memory pipe delivers.  The accumulator
// ring has four slots, so two of four z phases wrap and issue N = 128 + N = 64 instead.
//
// TMEM: F slots at columns [0,256) (4 x 64), E slots at [256,512).  Weight ring stage (24 KB) = the three dz taps of one
// in-plane tap: 192 rows x 64 B of fp16 W_hi, then 192 rows x 64 B of [W_hi8 | W_lo8]; rows ordered dz = +1, 0, -1
// (ascending output plane).  Warp roles as in conv_umma.cuh.
#pragma once
#include "conv_umma.cuh"

namespace vxb {
namespace umma {

constexpr int F8_WROWS = 192;                         // rows of one weight part (3 dz taps x 64 output channels)
constexpr int F8_WPART = F8_WROWS * 64;               // bytes of one part (fp16 or fp8) of a stage
constexpr int F8_WBYTES = 2 * F8_WPART;               // 24 KB per stage
constexpr int F8_WCHUNK_ROWS = 96;                    // TMA box rows for the weights (4 chunks per stage)

__device__ __forceinline__ void tc_mma_f8_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDesc64Hi) : "memory");
}
// kind::f16 (A = B = fp16) and kind::f8f6f4 (A = B = E4M3) share the descriptor bits: D = f32, formats 0, K-major, M = 128
__device__ __forceinline__ uint32_t f8c_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// column segments of one MMA step: the valid dz taps, merged while their accumulator slots are adjacent
struct F8Segs {
  int n;                 // number of segments (<= 3)
  uint32_t col[3];       // first accumulator column (slot * 64)
  uint32_t brow[3];      // first weight row, in descriptor units (rows * 64 B >> 4)
  uint32_t cols[3];      // N of the MMA (64, 128 or 192)
};
__device__ __forceinline__ F8Segs f8c_segments(int s2, uint32_t vmask, bool split_first, int dbg = 0) {
  F8Segs sg;
  sg.n = 0;
  int prev = -2;
#pragma unroll
  for (int j = 0; j < 3; ++j) {                        // j = 0, 1, 2 <-> dz tap index dzc = 2, 1, 0 <-> output plane zi-1, zi, zi+1
    if (!(vmask & (1u << (2 - j)))) continue;
    const int slot = (s2 + j) & 3;
    if (sg.n > 0 && prev == j - 1 && (slot != 0 || (dbg & 8)) && !(dbg & 4) && !(split_first && j == 2)) {   // dbg: timing experiments
      sg.cols[sg.n - 1] += 64;
    } else {
      sg.col[sg.n] = (uint32_t)slot * 64u;
      sg.brow[sg.n] = (uint32_t)j * (64u * 64u >> 4);
      sg.cols[sg.n] = 64;
      ++sg.n;
    }
    prev = j;
  }
  return sg;
}

template <int CL>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3_f8c_kernel(const __grid_constant__ CUtensorMap mapA0h, const __grid_constant__ CUtensorMap mapA0c,
                 const __grid_constant__ CUtensorMap mapA1h, const __grid_constant__ CUtensorMap mapA1c,
                 const __grid_constant__ CUtensorMap mapW16, const __grid_constant__ CUtensorMap mapW8, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(16) uint8_t epi_smem[4 * EPI_BYTES_PER_WARP];
  __shared__ __align__(8) uint64_t bars[2 * CV_SLABS + 2 * CV_WSTAGES + 8 + 4];
  __shared__ __align__(16) float tail_sw[64];
  __shared__ uint32_t tmem_base_smem;

  const int plane_bytes = 2 * p.box_rows * 64;                 // one plane (hi or c8) of a slab
  const int slab_bytes = 2 * plane_bytes;
  uint8_t* slab_base = smem;
  uint8_t* w_base = smem + CV_SLABS * slab_bytes;
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = bars + CV_SLABS;
  uint64_t* w_full = bars + 2 * CV_SLABS;
  uint64_t* w_empty = w_full + CV_WSTAGES;
  uint64_t* acc_full = w_empty + CV_WSTAGES;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* tail_done = acc_empty + 4;
  uint8_t* tailw_s = (uint8_t*)(((uintptr_t)(w_base + CV_WSTAGES * F8_WBYTES) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0h); tma_prefetch_desc(&mapA0c);
    tma_prefetch_desc(&mapA1h); tma_prefetch_desc(&mapA1c);
    tma_prefetch_desc(&mapW16); tma_prefetch_desc(&mapW8);
    for (int i = 0; i < CV_SLABS; ++i) { mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); }
    for (int i = 0; i < CV_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], CL); }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); mbar_init(&tail_done[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.tail_w) {
    // the 27 x 64 tail weights as fp16 hi / lo rows of a tcgen05 B operand (tap rows 27..31 are zero), as in conv_umma.cuh
    for (int i = threadIdx.x; i < 2 * 32 * 64; i += CV_THREADS) {
      const int set = i >> 11, tap = (i >> 6) & 31, ch = i & 63;
      const float* src = set ? p.tail_w2 : p.tail_w;
      const float f = (src && tap < 27) ? src[tap * 64 + ch] : 0.f;
      const __nv_bfloat16 h = pl_from_float(f);
      const __nv_bfloat16 l = pl_from_float(f - pl_to_float(h));
      uint8_t* base = tailw_s + set * CV_TAILW_BYTES;
      const int rl = 32 + tap;
      *reinterpret_cast<__nv_bfloat16*>(base + tap * 128 + (((ch >> 3) ^ (tap & 7)) << 4) + (ch & 7) * 2) = h;
      *reinterpret_cast<__nv_bfloat16*>(base + rl * 128 + (((ch >> 3) ^ (rl & 7)) << 4) + (ch & 7) * 2) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 64; i += CV_THREADS) tail_sw[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int groups = p.items / CL;
  const int Vp = p.Vp, Vp2 = Vp * Vp;

  auto decode = [&](int item, int& b, int& t, int& z0, int& lz) {
    const int per_chunk = p.items / p.zchunks;
    const int zc = item / per_chunk;
    int rem = item - zc * per_chunk;
    const int col = min(rem, p.B * p.tiles - 1);               // padded items repeat the last column (not stored)
    b = col / p.tiles;
    t = col - b * p.tiles;
    z0 = zc * p.lz;
    lz = min(p.lz, p.V - z0);
  };

  if (warp == 0) {
    // ===================================================== slab producer
    if (lane == 0) {
      int sb = 0, dbg_n = 0;
      uint32_t sph = 0;
      for (int g = cluster_id; g < groups; g += num_clusters) {
        int b, t, z0, lz;
        decode(g * CL + rank, b, t, z0, lz);
        for (int zi = z0 - 1; zi <= z0 + lz; ++zi) {
          const int row0 = (b * Vp + (zi + 1)) * Vp2 + t * 128 - (Vp + 1);
          for (int cb = 0; cb < p.ncb; ++cb) {
            const bool s1 = cb >= p.cb_src0;
            const int col = (s1 ? cb - p.cb_src0 : cb) * CV_KC;
            const CUtensorMap* mh = s1 ? &mapA1h : &mapA0h;
            const CUtensorMap* mc = s1 ? &mapA1c : &mapA0c;
            mbar_wait(&slab_empty[sb], sph ^ 1);
            uint8_t* s = slab_base + sb * slab_bytes;
            if ((p.debug_skip & 2) && dbg_n >= CV_SLABS) {           // timing experiment: never refill the slabs (wrong results)
              mbar_expect_tx(&slab_full[sb], 0);
              if (++sb == CV_SLABS) { sb = 0; sph ^= 1; }
              continue;
            }
            ++dbg_n;
            mbar_expect_tx(&slab_full[sb], slab_bytes);
            tma_load_2d(mh, &slab_full[sb], s, col, row0);
            tma_load_2d(mh, &slab_full[sb], s + p.box_rows * 64, col, row0 + p.box_rows);
            tma_load_2d(mc, &slab_full[sb], s + plane_bytes, col, row0);
            tma_load_2d(mc, &slab_full[sb], s + plane_bytes + p.box_rows * 64, col, row0 + p.box_rows);
            if (++sb == CV_SLABS) { sb = 0; sph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== weight producer: 4 chunks of 96 rows per stage, 4 / CL per CTA, multicast
    if (lane == 0) {
      int ws = 0, dbg_n = 0;
      uint32_t wph = 0;
      constexpr int CHUNKS = 4 / CL;
      const uint16_t mask = (uint16_t)((1u << CL) - 1);
      for (int g = cluster_id; g < groups; g += num_clusters) {
        int b, t, z0, lz;
        decode(g * CL + rank, b, t, z0, lz);
        for (int zi = z0 - 1; zi <= z0 + lz; ++zi) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            for (int tap9 = 0; tap9 < 9; ++tap9) {
              mbar_wait(&w_empty[ws], wph ^ 1);
              uint8_t* s = w_base + ws * F8_WBYTES;
              if ((p.debug_skip & 1) && dbg_n >= CV_WSTAGES) {       // timing experiment: never refill the ring (wrong results)
                mbar_expect_tx(&w_full[ws], 0);
                if (++ws == CV_WSTAGES) { ws = 0; wph ^= 1; }
                continue;
              }
              ++dbg_n;
              mbar_expect_tx(&w_full[ws], F8_WBYTES);
              const int r0 = (cb * 9 + tap9) * F8_WROWS;
#pragma unroll
              for (int c = 0; c < CHUNKS; ++c) {
                const int ch = (int)rank * CHUNKS + c;             // 0, 1: fp16 part; 2, 3: fp8 part
                const CUtensorMap* m = ch < 2 ? &mapW16 : &mapW8;
                uint8_t* d = s + ch * (F8_WCHUNK_ROWS * 64);
                const int row = r0 + (ch & 1) * F8_WCHUNK_ROWS;
                if (CL > 1) tma_load_2d_mc(m, &w_full[ws], d, 0, row, mask);
                else tma_load_2d(m, &w_full[ws], d, 0, row);
              }
              if (++ws == CV_WSTAGES) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== MMA issuer (whole warp runs the uniform loops, one elected lane issues)
    uint32_t leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint16_t mask = (uint16_t)((1u << CL) - 1);
    int sb = 0, ws = 0;
    uint32_t sph = 0, wph = 0;
    uint32_t acc_ph = 0;
    uint32_t w_ready = 0;
    for (int g = cluster_id; g < groups; g += num_clusters) {
      int b, t, z0, lz;
      decode(g * CL + rank, b, t, z0, lz);
      for (int zi = z0 - 1; zi <= z0 + lz; ++zi) {
        uint32_t vmask = 0;
        for (int dzc = 0; dzc < 3; ++dzc) {
          const int zo = zi - (dzc - 1);
          if (zo >= z0 && zo < z0 + lz) vmask |= 1u << dzc;
        }
        const int s2 = (zi - 1 - z0) & 3;                       // slot of output plane zi - 1 (dz tap index 2)
        const F8Segs seg = f8c_segments(s2, vmask, false, p.debug_skip);
        const F8Segs seg0 = f8c_segments(s2, vmask, true, p.debug_skip);      // first step of the plane: the new output plane zi + 1 starts from zero
        for (int cb = 0; cb < p.ncb; ++cb) {
          mbar_wait(&slab_full[sb], sph);
          tc_fence_after();
          const uint32_t a_hi = desc64_lo(smem_u32(slab_base + sb * slab_bytes));
          const uint32_t a_c8 = desc64_lo(smem_u32(slab_base + sb * slab_bytes + plane_bytes));
          if (cb == 0 && (vmask & 1u)) {
            const int slot = (zi + 1 - z0) & 3;
            mbar_wait(&acc_empty[slot], ((acc_ph >> slot) & 1u) ^ 1u);
            acc_ph ^= 1u << slot;
            tc_fence_after();
          }
#pragma unroll 1
          for (int tap9 = 0; tap9 < 9; ++tap9) {
            if (!w_ready) mbar_wait(&w_full[ws], wph);
            tc_fence_after();
            const uint32_t w16 = desc64_lo(smem_u32(w_base + ws * F8_WBYTES));
            const uint32_t w8 = w16 + (uint32_t)(F8_WPART >> 4);
            const uint32_t arow = (uint32_t)((tap9 / 3) * Vp + (tap9 % 3)) * 4u;
            const bool first = (cb == 0 && tap9 == 0 && (vmask & 1u));
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < CV_KC / 16; ++ks) {
                const uint32_t aoff = arow + (uint32_t)(ks * 2);
                const F8Segs& sg = (first && ks == 0) ? seg0 : seg;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                  if (i >= sg.n) break;
                  // the segment that holds only the new plane (slot of zi + 1, weight rows 128..191) starts its accumulators
                  const uint32_t acc_on = (first && ks == 0 && sg.brow[i] == 2u * (64u * 64u >> 4)) ? 0u : 1u;
                  const uint32_t id = f8c_idesc((int)sg.cols[i]);
                  tc_mma_bf16_w(tmem_u + sg.col[i], a_hi + aoff, w16 + sg.brow[i] + (uint32_t)(ks * 2), id, acc_on);
                  tc_mma_f8_w(tmem_u + 256u + sg.col[i], a_c8 + aoff, w8 + sg.brow[i] + (uint32_t)(ks * 2), id, acc_on);
                }
              }
              if (CL > 1) tc_commit_mc(&w_empty[ws], mask);
              else tc_commit(&w_empty[ws]);
            }
            __syncwarp();
            if (++ws == CV_WSTAGES) { ws = 0; wph ^= 1; }
            w_ready = mbar_test(&w_full[ws], wph);
          }
          if (leader) tc_commit(&slab_empty[sb]);
          __syncwarp();
          if (++sb == CV_SLABS) { sb = 0; sph ^= 1; }
        }
        const int zdone = zi - 1;
        if (zdone >= z0 && zdone < z0 + lz && leader) tc_commit(&acc_full[(zdone - z0) & 3]);
        __syncwarp();
      }
    }
  } else {
    // ===================================================== epilogue (warps 3..6)
    const int q = warp & 3;
    float* stage = reinterpret_cast<float*>(epi_smem + q * EPI_BYTES_PER_WARP);
    RowInfo* ri = reinterpret_cast<RowInfo*>(stage + 32 * EPI_STAGE_LD);
    const int tr = lane >> 3, tc = (lane & 7) * 4;
    const float slope = p.act_slope >= 0.f ? p.act_slope : 1.f;
    const float escale = p.f8s[2];                      // 2^-s: scale of the fp8 correction accumulator
    uint32_t full_ph = 0;
    uint32_t tail_ph = 0;
    const int V = p.V;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    for (int g = cluster_id; g < groups; g += num_clusters) {
      int b, t, z0, lz;
      const int item = g * CL + rank;
      decode(item, b, t, z0, lz);
      const int per_chunk = p.items / p.zchunks;
      const bool real = (item % per_chunk) < p.B * p.tiles;
      if (p.tail_w) {
        TailRowInfo* tri = reinterpret_cast<TailRowInfo*>(ri);
        const int rr = t * 128 + q * 32 + lane;
        const int yp = rr / Vp, xp = rr - yp * Vp;
        const bool ok = real && rr < Vp2 && yp >= 1 && yp <= V && xp >= 1 && xp <= V;
        const long long yx = ok ? ((long long)(yp - 1) * V + (xp - 1)) : -1;
        tri->orow[lane] = yx;
        tri->px[lane] = ok ? ss_lin_coord(yp - 1, V) : 0.f;
        tri->pz[lane] = ok ? ss_lin_coord(xp - 1, V) : 0.f;
        __syncwarp();
        SSState st[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) st[k] = {-INFINITY, 0.f, 0.f, 0.f, 0.f, -INFINITY};
        const size_t V3 = (size_t)V * V * V;
        for (int zo = z0; zo < z0 + lz; ++zo) {
          const int slot = (zo - z0) & 3;
          const uint32_t t_f = tmem_base + (uint32_t)(slot * 64);          // F slot (then: U as the tail's TMEM A operand)
          const uint32_t t_e = tmem_base + 256u + (uint32_t)(slot * 64);   // E slot (then: the tail's tap products)
          const float py = ss_lin_coord(zo, V);
          mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
          full_ph ^= 1u << slot;
          tc_fence_after();
          float pt[27], pt2[27];
#pragma unroll
          for (int tp = 0; tp < 27; ++tp) { pt[tp] = 0.f; pt2[tp] = 0.f; }
          uint32_t uh[32], ul[32];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = cc * 32;
            uint32_t v0[32], v1[32];
            tc_ld32(t_f + t_lane + (uint32_t)c0, v0);
            tc_ld32(t_e + t_lane + (uint32_t)c0, v1);
            float u[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(tail_sw + c0 + j);
              u[j] = fmaf(__uint_as_float(v1[j]), escale, __uint_as_float(v0[j])) + bv.x;
              u[j + 1] = fmaf(__uint_as_float(v1[j + 1]), escale, __uint_as_float(v0[j + 1])) + bv.y;
              u[j + 2] = fmaf(__uint_as_float(v1[j + 2]), escale, __uint_as_float(v0[j + 2])) + bv.z;
              u[j + 3] = fmaf(__uint_as_float(v1[j + 3]), escale, __uint_as_float(v0[j + 3])) + bv.w;
#pragma unroll
              for (int k = 0; k < 4; ++k) u[j + k] = fmaxf(u[j + k], u[j + k] * slope);
              *reinterpret_cast<float4*>(stage + lane * EPI_STAGE_LD + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const __nv_bfloat162 hh = pl2_from_floats(u[2 * j], u[2 * j + 1]);
              const float2 ff = pl2_to_float2(hh);
              const __nv_bfloat162 ll = pl2_from_floats(u[2 * j] - ff.x, u[2 * j + 1] - ff.y);
              uh[cc * 16 + j] = *reinterpret_cast<const uint32_t*>(&hh);
              ul[cc * 16 + j] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + tr;
              if (tri->orow[r] >= 0) {
                const float4 x = *reinterpret_cast<const float4*>(stage + r * EPI_STAGE_LD + tc);
                const float px = tri->px[r], pz = tri->pz[r];
                ss_update(st[cc * 4 + 0], x.x, px, py, pz);
                ss_update(st[cc * 4 + 1], x.y, px, py, pz);
                ss_update(st[cc * 4 + 2], x.z, px, py, pz);
                ss_update(st[cc * 4 + 3], x.w, px, py, pz);
              }
            }
            __syncwarp();
          }
          {
            // tap products on the tensor core (three-term fp16 split, A = U from tensor memory): U -> F slot, products -> E slot
            constexpr uint32_t kDescHi128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t idesc_t64 = make_idesc(64), idesc_t32 = make_idesc(32);
            tc_st32(t_f + t_lane, uh);
            tc_st32(t_f + t_lane + 32u, ul);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            const int nsets = p.tail_w2 ? 2 : 1;
            for (int set = 0; set < nsets; ++set) {
              asm volatile("bar.sync 1, 128;" ::: "memory");
              if (warp == 3 && lane == 0) {
                tc_fence_after();
                const uint32_t wb = ((smem_u32(tailw_s + set * CV_TAILW_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  tc_mma_f16_ts(t_e, t_f + 8u * ks, wb + 2u * ks, kDescHi128, idesc_t64, ks != 0);
                  tc_mma_f16_ts(t_e, t_f + 32u + 8u * ks, wb + 2u * ks, kDescHi128, idesc_t32, 1u);
                }
                tc_commit(&tail_done[slot]);
              }
              mbar_wait(&tail_done[slot], (tail_ph >> slot) & 1u);
              tail_ph ^= 1u << slot;
              tc_fence_after();
              uint32_t d0[32], d1[32];
              tc_ld32_nowait(t_e + t_lane, d0);
              tc_ld32_nowait(t_e + t_lane + 32u, d1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              tc_fence_before();
#pragma unroll
              for (int tp = 0; tp < 27; ++tp) {
                const float v = __uint_as_float(d0[tp]) + __uint_as_float(d1[tp]);
                if (set == 0) pt[tp] = v; else pt2[tp] = v;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);
          if (yx >= 0) {
            float* dst = p.ptap + (size_t)b * 27 * V3 + (size_t)zo * V * V + yx;
#pragma unroll
            for (int tp = 0; tp < 27; ++tp) dst[(size_t)tp * V3] = pt[tp];
            if (p.tail_w2) {
              float* dst2 = p.ptap2 + (size_t)b * 27 * V3 + (size_t)zo * V * V + yx;
#pragma unroll
              for (int tp = 0; tp < 27; ++tp) dst2[(size_t)tp * V3] = pt2[tp];
            }
          }
        }
        if (real) {
          const int zc = z0 / p.lz;
          const int chunks = p.zchunks * p.tiles * 16;
          const int chunk = ((zc * p.tiles + t) * 4 + q) * 4 + tr;
          float* o = p.ss_partial + ((size_t)b * chunks + chunk) * 6 * 64;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = (k >> 2) * 32 + tc + (k & 3);
            o[c] = st[k].m; o[64 + c] = st[k].s; o[128 + c] = st[k].sx; o[192 + c] = st[k].sy; o[256 + c] = st[k].sz;
            o[320 + c] = st[k].rm;
          }
        }
      } else {
        {
          const int rr = t * 128 + q * 32 + lane;
          const int yp = rr / Vp, xp = rr - yp * Vp;
          const bool ok = real && rr < Vp2 && yp >= 1 && yp <= V && xp >= 1 && xp <= V;
          ri->orow[lane] = ok ? ((long long)b * V * V * V + (long long)(yp - 1) * V + (xp - 1)) : -1;
          __syncwarp();
        }
        for (int zo = z0; zo < z0 + lz; ++zo) {
          const int slot = (zo - z0) & 3;
          mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
          full_ph ^= 1u << slot;
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v0[32], v1[32];
            tc_ld32(tmem_base + t_lane + (uint32_t)(slot * 64 + c0), v0);
            tc_ld32(tmem_base + t_lane + 256u + (uint32_t)(slot * 64 + c0), v1);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage + lane * EPI_STAGE_LD + j) =
                  make_float4(fmaf(__uint_as_float(v1[j]), escale, __uint_as_float(v0[j])),
                              fmaf(__uint_as_float(v1[j + 1]), escale, __uint_as_float(v0[j + 1])),
                              fmaf(__uint_as_float(v1[j + 2]), escale, __uint_as_float(v0[j + 2])),
                              fmaf(__uint_as_float(v1[j + 3]), escale, __uint_as_float(v0[j + 3])));
            __syncwarp();
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + c0 + tc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + tr;
              const long long orow = ri->orow[r];
              float4 x = *reinterpret_cast<const float4*>(stage + r * EPI_STAGE_LD + tc);
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
              x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);
              x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
              if (orow >= 0)
                *reinterpret_cast<float4*>(p.out + (orow + (long long)zo * V * V) * 64 + c0 + tc) = x;
            }
            __syncwarp();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[slot]);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace umma
}  // namespace vxb

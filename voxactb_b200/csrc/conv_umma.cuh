// Input-stationary 3x3x3 convolution on tcgen05 (split-fp16 x3, fp32 accumulate in TMEM) for sm_100a.
//
//   out[b,z,y,x,:] = act(bias + sum_{src,dz,dy,dx} W[:, tap, src ch] . in_src[b, z+dz, y+dy, x+dx, :])     (Co = 64)
//
// Inputs are 16-bit hi/lo planes (planes16.cuh) of the replicate-PADDED channels-last grids [B, Vp, Vp, Vp, 64] (Vp = V + 2), so
// a tap is a constant shift of the flat padded row index.  The GEMM engine in umma_gemm.cuh re-fetches a
// 128-row operand tile from L2 for every tap (27x) and is L2-bandwidth bound; this kernel instead
//   * stages ONE slab per (z plane, 32-channel block): the 128 output rows of a (y,x)-plane tile plus a
//     Vp+1 row halo on each side (hi and lo planes, SWIZZLE_64B rows of 64 bytes), fetched once by TMA;
//   * expresses the 9 in-plane taps as ROW-SHIFTED shared-memory matrix descriptors into that slab;
//   * marches along z with the slab stationary: input plane zi feeds the three output planes zi-1, zi, zi+1,
//     each with its own TMEM accumulator (4 slots x 128 columns), so every activation byte is staged once;
//   * streams the weights [W_hi ; W_lo] (N = 128 rows) through a deep ring, multicast across a thread-block
//     cluster so the L2 -> SM weight traffic is divided by the cluster size;
//   * issues per 16-wide k step   D[:, 0:128] += A_hi [W_hi ; W_lo]^T   and   D[:, 0:64] += A_lo W_hi^T
//     (the three split terms in two MMAs); the epilogue adds the two column halves.
//
// Fused tail (ConvParams::tail_w): the activation u = act(conv) is never stored.  Per output plane the epilogue
//   * accumulates the SpatialSoftmax3D / max-pool partials of u (column domain, through a per-warp smem transpose);
//   * writes u back as fp16 hi/lo pairs into columns 0..63 of the accumulator slot it has just drained (tcgen05.st)
//     and has one of its threads issue eight small MMAs with the A operand READ FROM TENSOR MEMORY
//     (P[128 rows x 32 taps] = U . Wt^T, same three-term split; B = the 27 x 64 trans_decoder weights, split once
//     per CTA into a SWIZZLE_128B tile) into columns 64..127; the 27 tap products per voxel go to `ptap`
//     (trans_gather_kernel sums them over the 3x3x3 neighbourhood afterwards).
//   The CTA-pair experiment (PAIR = true) keeps the older CUDA-core tail (weights broadcast from shared memory).
//
// Warp roles (224 threads): 0 = slab TMA producer, 1 = weight TMA producer, 2 = MMA issuer (+ TMEM alloc),
// 3-6 = epilogue (TMEM -> registers -> smem transpose -> coalesced fp32 stores, or the fused tail).
#pragma once
#include "umma_gemm.cuh"
#include "stream_ops.cuh"

namespace vxb {
namespace umma {

constexpr int CV_KC = 32;             // channels per slab (64-byte rows, SWIZZLE_64B)
constexpr int CV_THREADS = 224;
constexpr int CV_WSTAGES = 4;         // weight ring depth; a stage holds the 3 dx taps of one (dz, dy): 24 KB
constexpr int CV_SLABS = 2;           // slab double buffering
constexpr int CV_TAPBYTES = 128 * CV_KC * 2;      // [W_hi ; W_lo] of one tap
constexpr int CV_WBYTES = 3 * CV_TAPBYTES;

struct ConvParams {
  int B, V, Vp;
  int ncb;              // 32-channel blocks over all sources
  int cb_src0;          // blocks taken from source 0
  int tiles;            // (y,x)-plane tiles of 128 flat rows
  int zchunks, lz;      // z chunks per column, planes per chunk
  int items;            // B * tiles * zchunks (padded to a multiple of the cluster size by the launcher)
  int items_real;
  int slab_rows;        // 128 + 2 (Vp + 1)
  int box_rows;         // rows per slab TMA box (two boxes per plane)
  int base_off_mode;    // descriptor base-offset policy for row-shifted operands (see make_desc64)
  int debug_skip;       // unused (was: timing experiments)
  const float* bias;
  float act_slope;
  float* out;           // [B, V, V, V, 64] fp32 (null in tail mode)
  // fused tail (trans_decoder taps + ss_final / max-pool partials computed from the accumulator rows; the
  // convolution output itself is never written):
  const float* tail_w;  // [27][64] tap-major weights of the 64 -> 1 convolution, or null
  float* ptap;          // [B][27][V^3]: ptap[b][t][v] = <tail_w[t], u[b, v, :]>
  const float* tail_w2; // optional second tap set (2 robots: trans_decoder_left_arm) and its products
  float* ptap2;
  float* ss_partial;    // [B][chunks][6][64], chunks = zchunks * tiles * 16 (one per item x warp x row group)
  const float* f8s;     // conv_f8c.cuh only: device scalars of the fp8 correction terms ([2] = 2^-s, see umma_ops.cu f8c_*)
};

struct TailRowInfo {
  long long orow[32];   // (y*V + x) of the row's voxel inside an output plane, -1 = halo / padding row
  float px[32];         // SpatialSoftmax3D coordinates of the row (pos_x along H, pos_z along W)
  float pz[32];
};
static_assert(sizeof(TailRowInfo) <= sizeof(RowInfo), "TailRowInfo must fit the RowInfo slot of the staging area");

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// K-major SWIZZLE_64B descriptor: rows of 64 bytes, 8-row groups 512 bytes apart.  `saddr` may be shifted by
// whole rows (64 B) from the 512-byte aligned slab base.  Measured on B200 (tests/test_ops_gpu.py::test_conv3d):
// the hardware applies the swizzle XOR to the absolute shared-memory address, so a row-shifted start address
// needs base offset 0 (mode 0, the default); mode 1 (base offset = address phase) gives wrong results.
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr, int base_off_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  if (base_off_mode == 1) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= (uint64_t)4 << 61;                             // SWIZZLE_64B
  return d;
}

// ---- CTA-pair (cta_group::2) variants: one MMA spans the two CTAs of a cluster (M = 256: 128 rows from each CTA's
// own shared memory), the N rows of the B operand are split between the two CTAs, so each SM reads only half of the
// weights per MMA.  Loads of both CTAs signal the LEADER's (rank 0) mbarrier: clearing bit 24 of a shared-window
// address selects the even CTA of the pair.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on the barrier of BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// low / high words of the SWIZZLE_64B K-major descriptor; the low word advances by (bytes >> 4)
__device__ __forceinline__ uint32_t desc64_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
constexpr uint32_t kDesc64Hi = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
__device__ __forceinline__ void tc_mma_bf16_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDesc64Hi) : "memory");
}

__device__ __forceinline__ void tc_mma_pair_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDesc64Hi) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m256(int n) {
  return (1u << 4) | VXB_IDESC_AB_FORMAT | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
constexpr int CV_TAILW_BYTES = 64 * 128;           // one tail weight set as a tcgen05 B operand (64 rows x 128 B)
constexpr int CV_PAIR_TAPBYTES = 96 * CV_KC * 2;      // per CTA and tap: 64 rows (its half of [W_hi;W_lo]) + 32 rows (its half of W_hi)
constexpr int CV_PAIR_WBYTES = 3 * CV_PAIR_TAPBYTES;

template <int CL, bool PAIR>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3_umma_kernel(const __grid_constant__ CUtensorMap mapA0h, const __grid_constant__ CUtensorMap mapA0l,
                  const __grid_constant__ CUtensorMap mapA1h, const __grid_constant__ CUtensorMap mapA1l,
                  const __grid_constant__ CUtensorMap mapW, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(16) uint8_t epi_smem[4 * EPI_BYTES_PER_WARP];
  __shared__ __align__(8) uint64_t bars[2 * CV_SLABS + 2 * CV_WSTAGES + 8 + 4];
  // tail mode: conv bias (+ the fp32 tail weights of the CUDA-core tail the CTA-pair variant still uses)
  constexpr int TAIL_BIAS_OFF = PAIR ? 2 * 27 * 64 : 0;
  __shared__ __align__(16) float tail_sw[TAIL_BIAS_OFF + 64];
  __shared__ uint32_t tmem_base_smem;

  static_assert(!PAIR || CL == 2, "the CTA-pair variant runs on clusters of two");
  constexpr int WBYTES = PAIR ? CV_PAIR_WBYTES : CV_WBYTES;
  const int plane_bytes = 2 * p.box_rows * 64;                 // one plane (hi or lo) of a slab
  const int slab_bytes = 2 * plane_bytes;
  uint8_t* slab_base = smem;
  uint8_t* w_base = smem + CV_SLABS * slab_bytes;
  uint64_t* slab_full = bars;
  uint64_t* slab_empty = bars + CV_SLABS;
  uint64_t* w_full = bars + 2 * CV_SLABS;
  uint64_t* w_empty = w_full + CV_WSTAGES;
  uint64_t* acc_full = w_empty + CV_WSTAGES;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* tail_done = acc_empty + 4;           // tensor-core tail: the tap-product MMAs of a slot have retired
  // tensor-core tail (PAIR = false): [W_hi (32 tap rows) ; W_lo (32 rows)] x 64 channels fp16, SWIZZLE_128B K-major, per weight set
  uint8_t* tailw_s = (uint8_t*)(((uintptr_t)(w_base + CV_WSTAGES * WBYTES) + 1023) & ~(uintptr_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CL > 1) ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0h); tma_prefetch_desc(&mapA0l);
    tma_prefetch_desc(&mapA1h); tma_prefetch_desc(&mapA1l);
    tma_prefetch_desc(&mapW);
    for (int i = 0; i < CV_SLABS; ++i) { mbar_init(&slab_full[i], 1); mbar_init(&slab_empty[i], 1); }
    for (int i = 0; i < CV_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], PAIR ? 1 : CL); }
    for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], PAIR ? 8 : 4); mbar_init(&tail_done[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.tail_w) {
    if constexpr (PAIR) {
      for (int i = threadIdx.x; i < 27 * 64; i += CV_THREADS) {
        tail_sw[i] = p.tail_w[i];
        tail_sw[27 * 64 + i] = p.tail_w2 ? p.tail_w2[i] : 0.f;
      }
    } else {
      // split the 27 x 64 tail weights into fp16 hi / lo rows of the B operand (tap rows 27..31 are zero)
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
    }
    for (int i = threadIdx.x; i < 64; i += CV_THREADS) tail_sw[TAIL_BIAS_OFF + i] = p.bias[i];
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int groups = p.items / CL;
  const int Vp = p.Vp, Vp2 = Vp * Vp;

  // item -> (b, tile, zchunk): z chunk slowest so that the CTAs of a cluster always share the chunk length
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
      int sb = 0;
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
            const CUtensorMap* ml = s1 ? &mapA1l : &mapA0l;
            mbar_wait(&slab_empty[sb], sph ^ 1);
            uint8_t* s = slab_base + sb * slab_bytes;
            if constexpr (PAIR) {
              // both CTAs' slabs complete on the leader's barrier
              if (rank == 0) mbar_expect_tx(&slab_full[sb], 2 * slab_bytes);
              tma_load_2d_pair(mh, &slab_full[sb], s, col, row0);
              tma_load_2d_pair(mh, &slab_full[sb], s + p.box_rows * 64, col, row0 + p.box_rows);
              tma_load_2d_pair(ml, &slab_full[sb], s + plane_bytes, col, row0);
              tma_load_2d_pair(ml, &slab_full[sb], s + plane_bytes + p.box_rows * 64, col, row0 + p.box_rows);
            } else {
              mbar_expect_tx(&slab_full[sb], slab_bytes);
              tma_load_2d(mh, &slab_full[sb], s, col, row0);
              tma_load_2d(mh, &slab_full[sb], s + p.box_rows * 64, col, row0 + p.box_rows);
              tma_load_2d(ml, &slab_full[sb], s + plane_bytes, col, row0);
              tma_load_2d(ml, &slab_full[sb], s + plane_bytes + p.box_rows * 64, col, row0 + p.box_rows);
            }
            if (++sb == CV_SLABS) { sb = 0; sph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== weight producer (each CTA loads 128/CL rows, multicast)
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      constexpr int WROWS = 128 / CL;
      const uint16_t mask = (uint16_t)((1u << CL) - 1);
      for (int g = cluster_id; g < groups; g += num_clusters) {
        int b, t, z0, lz;
        decode(g * CL + rank, b, t, z0, lz);
        for (int zi = z0 - 1; zi <= z0 + lz; ++zi) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            // a stage = the (up to) three dz taps of one in-plane tap (dy, dx): the MMA warp interleaves the three
            // output planes (different TMEM accumulators) so that back-to-back MMAs never hit the same accumulator
            int nvalid = 0;
            for (int dzc = 0; dzc < 3; ++dzc) {
              const int zo = zi - (dzc - 1);
              nvalid += (zo >= z0 && zo < z0 + lz) ? 1 : 0;
            }
            for (int tap9 = 0; tap9 < 9; ++tap9) {
              mbar_wait(&w_empty[ws], wph ^ 1);
              uint8_t* s = w_base + ws * WBYTES;
              if constexpr (PAIR) {
                if (rank == 0) mbar_expect_tx(&w_full[ws], 2 * nvalid * CV_PAIR_TAPBYTES);
              } else {
                mbar_expect_tx(&w_full[ws], nvalid * CV_TAPBYTES);
              }
              for (int dzc = 0; dzc < 3; ++dzc) {
                const int zo = zi - (dzc - 1);
                if (zo < z0 || zo >= z0 + lz) continue;
                const int r0 = (cb * 27 + dzc * 9 + tap9) * 128;
                if constexpr (PAIR) {
                  // global rows of a tap: [W_hi 0..63][W_lo 64..127].  CTA r keeps its half of [W_hi;W_lo] (64 rows,
                  // the N = 128 operand) followed by its half of W_hi (32 rows, the N = 64 operand); map box = 32 rows
                  uint8_t* d = s + dzc * CV_PAIR_TAPBYTES;
                  tma_load_2d_pair(&mapW, &w_full[ws], d, 0, r0 + (int)rank * 64);
                  tma_load_2d_pair(&mapW, &w_full[ws], d + 32 * 64, 0, r0 + (int)rank * 64 + 32);
                  tma_load_2d_pair(&mapW, &w_full[ws], d + 64 * 64, 0, r0 + (int)rank * 32);
                } else {
                  uint8_t* d = s + dzc * CV_TAPBYTES + rank * WROWS * 64;
                  if (CL > 1) tma_load_2d_mc(&mapW, &w_full[ws], d, 0, r0 + (int)rank * WROWS, mask);
                  else tma_load_2d(&mapW, &w_full[ws], d, 0, r0);
                }
              }
              if (++ws == CV_WSTAGES) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 2) {
    // ===================================================== MMA issuer
    // The whole warp runs the (warp-uniform) loops so that descriptors live in uniform registers; one elected
    // lane issues the tcgen05 instructions.
    {
      uint32_t leader;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc128 = make_idesc(128), idesc64 = make_idesc(64);
      const uint16_t mask = (uint16_t)((1u << CL) - 1);
      int sb = 0, ws = 0;
      uint32_t sph = 0, wph = 0;
      uint32_t acc_ph = 0;                         // bit s: parity of the last completed acc_empty phase of slot s
      uint32_t w_ready = 0;
      constexpr uint32_t idesc128p = make_idesc_m256(128), idesc64p = make_idesc_m256(64);
      // CTA-pair mode: the leader issues every MMA for both CTAs; the peer's MMA warp has nothing to do
      for (int g = cluster_id; g < ((PAIR && rank != 0) ? 0 : groups); g += num_clusters) {
        int b, t, z0, lz;
        decode(g * CL + rank, b, t, z0, lz);
        for (int zi = z0 - 1; zi <= z0 + lz; ++zi) {
          for (int cb = 0; cb < p.ncb; ++cb) {
            mbar_wait(&slab_full[sb], sph);
            tc_fence_after();
            const uint32_t a_hi_lo = desc64_lo(smem_u32(slab_base + sb * slab_bytes));
            const uint32_t a_lo_lo = desc64_lo(smem_u32(slab_base + sb * slab_bytes + plane_bytes));
            // valid output planes of this input plane: zo = zi - (dzc - 1); accumulator slot = (zo - z0) & 3
            uint32_t vmask = 0;
            for (int dzc = 0; dzc < 3; ++dzc) {
              const int zo = zi - (dzc - 1);
              if (zo >= z0 && zo < z0 + lz) vmask |= 1u << dzc;
            }
            // the first contribution to an output plane comes from input plane zo-1 (dzc = 0), channel block 0:
            // the accumulator slot must have been drained by the epilogue
            if (cb == 0 && (vmask & 1u)) {
              const int slot = (zi + 1 - z0) & 3;
              mbar_wait(&acc_empty[slot], ((acc_ph >> slot) & 1u) ^ 1u);
              acc_ph ^= 1u << slot;
              tc_fence_after();
            }
            const uint32_t d0 = tmem_u + (uint32_t)(((zi + 1 - z0) & 3) * 128);   // dzc = 0 -> zo = zi + 1
            const uint32_t d1 = tmem_u + (uint32_t)(((zi - z0) & 3) * 128);       // dzc = 1 -> zo = zi
            const uint32_t d2 = tmem_u + (uint32_t)(((zi - 1 - z0) & 3) * 128);   // dzc = 2 -> zo = zi - 1
            const uint32_t dt[3] = {d0, d1, d2};
#pragma unroll 1
            for (int tap9 = 0; tap9 < 9; ++tap9) {
              // the probe of this stage's barrier was issued right after the previous stage's MMAs (w_ready),
              // so its latency overlaps their execution instead of sitting between two batches of MMAs
              if (!w_ready) mbar_wait(&w_full[ws], wph);
              tc_fence_after();
              const uint32_t wlo = desc64_lo(smem_u32(w_base + ws * WBYTES));
              // row shift of tap (dy, dx): dyc * Vp + dxc rows of 64 bytes = 4 descriptor units per row
              const uint32_t arow = (uint32_t)((tap9 / 3) * Vp + (tap9 % 3)) * 4u;
              const uint32_t first = (cb == 0 && tap9 == 0) ? 1u : 0u;     // dzc = 0 starts its accumulator here
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < CV_KC / 16; ++ks) {
                  const uint32_t aoff = arow + (uint32_t)(ks * 2);
                  // consecutive MMAs go to DIFFERENT accumulators (the three output planes): an MMA that accumulates
                  // into the tile the previous one wrote waits for it (~100 clk for these small shapes)
#pragma unroll
                  for (int dzc = 0; dzc < 3; ++dzc) {
                    if (!(vmask & (1u << dzc))) continue;
                    const uint32_t acc_on = (dzc == 0 && first && ks == 0) ? 0u : 1u;
                    if constexpr (PAIR) {
                      tc_mma_pair_w(dt[dzc], a_hi_lo + aoff, wlo + (uint32_t)(dzc * (CV_PAIR_TAPBYTES >> 4) + ks * 2), idesc128p, acc_on);
                    } else {
                      tc_mma_bf16_w(dt[dzc], a_hi_lo + aoff, wlo + (uint32_t)(dzc * (CV_TAPBYTES >> 4) + ks * 2), idesc128, acc_on);
                    }
                  }
#pragma unroll
                  for (int dzc = 0; dzc < 3; ++dzc) {
                    if (!(vmask & (1u << dzc))) continue;
                    if constexpr (PAIR) {
                      tc_mma_pair_w(dt[dzc], a_lo_lo + aoff, wlo + (uint32_t)(dzc * (CV_PAIR_TAPBYTES >> 4) + ks * 2) + ((64 * 64) >> 4), idesc64p, 1u);
                    } else {
                      tc_mma_bf16_w(dt[dzc], a_lo_lo + aoff, wlo + (uint32_t)(dzc * (CV_TAPBYTES >> 4) + ks * 2), idesc64, 1u);
                    }
                  }
                }
                if constexpr (PAIR) tc_commit_pair(&w_empty[ws]);
                else if (CL > 1) tc_commit_mc(&w_empty[ws], mask);
                else tc_commit(&w_empty[ws]);
              }
              __syncwarp();
              if (++ws == CV_WSTAGES) { ws = 0; wph ^= 1; }
              w_ready = mbar_test(&w_full[ws], wph);
            }
            if (leader) { if constexpr (PAIR) tc_commit_pair(&slab_empty[sb]); else tc_commit(&slab_empty[sb]); }
            __syncwarp();
            if (++sb == CV_SLABS) { sb = 0; sph ^= 1; }
          }
          // input plane zi done: output plane zi-1 has all of its three input planes
          const int zdone = zi - 1;
          if (zdone >= z0 && zdone < z0 + lz && leader) {
            if constexpr (PAIR) tc_commit_pair(&acc_full[(zdone - z0) & 3]); else tc_commit(&acc_full[(zdone - z0) & 3]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================== epilogue (warps 3..6)
    const int q = warp & 3;
    float* stage = reinterpret_cast<float*>(epi_smem + q * EPI_BYTES_PER_WARP);
    RowInfo* ri = reinterpret_cast<RowInfo*>(stage + 32 * EPI_STAGE_LD);
    const int tr = lane >> 3, tc = (lane & 7) * 4;
    const float slope = p.act_slope >= 0.f ? p.act_slope : 1.f;
    uint32_t full_ph = 0;                          // bit s: parity to wait for on acc_full[s]
    [[maybe_unused]] uint32_t tail_ph = 0;         // bit s: parity to wait for on tail_done[s]
    const int V = p.V;
    for (int g = cluster_id; g < groups; g += num_clusters) {
      int b, t, z0, lz;
      const int item = g * CL + rank;
      decode(item, b, t, z0, lz);
      const int per_chunk = p.items / p.zchunks;
      const bool real = (item % per_chunk) < p.B * p.tiles;
      if (p.tail_w) {
        // ---------------- fused tail: u = act(conv) stays in registers / smem
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
          const float py = ss_lin_coord(zo, V);
          mbar_wait(&acc_full[slot], (full_ph >> slot) & 1u);
          full_ph ^= 1u << slot;
          tc_fence_after();
          float pt[27], pt2[27];
#pragma unroll
          for (int tp = 0; tp < 27; ++tp) { pt[tp] = 0.f; pt2[tp] = 0.f; }
          [[maybe_unused]] uint32_t uh[32], ul[32];            // tensor-core tail: this row's 64 channels as packed fp16 pairs (hi, lo)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c0 = cc * 32;
            uint32_t v0[32], v1[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128 + c0), v0);
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128 + 64 + c0), v1);
            float u[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(tail_sw + TAIL_BIAS_OFF + c0 + j);
              u[j] = __uint_as_float(v0[j]) + __uint_as_float(v1[j]) + bv.x;
              u[j + 1] = __uint_as_float(v0[j + 1]) + __uint_as_float(v1[j + 1]) + bv.y;
              u[j + 2] = __uint_as_float(v0[j + 2]) + __uint_as_float(v1[j + 2]) + bv.z;
              u[j + 3] = __uint_as_float(v0[j + 3]) + __uint_as_float(v1[j + 3]) + bv.w;
#pragma unroll
              for (int k = 0; k < 4; ++k) u[j + k] = fmaxf(u[j + k], u[j + k] * slope);
              *reinterpret_cast<float4*>(stage + lane * EPI_STAGE_LD + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
            }
            if constexpr (!PAIR) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const __nv_bfloat162 hh = pl2_from_floats(u[2 * j], u[2 * j + 1]);
                const float2 ff = pl2_to_float2(hh);
                const __nv_bfloat162 ll = pl2_from_floats(u[2 * j] - ff.x, u[2 * j + 1] - ff.y);
                uh[cc * 16 + j] = *reinterpret_cast<const uint32_t*>(&hh);
                ul[cc * 16 + j] = *reinterpret_cast<const uint32_t*>(&ll);
              }
            }
            // CTA-pair variant: the 27 tap dot products of this row on the CUDA cores (weights broadcast from shared memory)
            if constexpr (PAIR) {
#pragma unroll
            for (int tp = 0; tp < 27; ++tp) {
              const float4* w4 = reinterpret_cast<const float4*>(tail_sw + tp * 64 + c0);
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float4 wa = w4[j], wb = w4[j + 1];
                a0 = fmaf(u[4 * j], wa.x, a0); a0 = fmaf(u[4 * j + 1], wa.y, a0);
                a0 = fmaf(u[4 * j + 2], wa.z, a0); a0 = fmaf(u[4 * j + 3], wa.w, a0);
                a1 = fmaf(u[4 * j + 4], wb.x, a1); a1 = fmaf(u[4 * j + 5], wb.y, a1);
                a1 = fmaf(u[4 * j + 6], wb.z, a1); a1 = fmaf(u[4 * j + 7], wb.w, a1);
              }
              pt[tp] += a0 + a1;
            }
            if (p.tail_w2) {
#pragma unroll
              for (int tp = 0; tp < 27; ++tp) {
                const float4* w4 = reinterpret_cast<const float4*>(tail_sw + (27 + tp) * 64 + c0);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                  const float4 wa = w4[j], wb = w4[j + 1];
                  a0 = fmaf(u[4 * j], wa.x, a0); a0 = fmaf(u[4 * j + 1], wa.y, a0);
                  a0 = fmaf(u[4 * j + 2], wa.z, a0); a0 = fmaf(u[4 * j + 3], wa.w, a0);
                  a1 = fmaf(u[4 * j + 4], wb.x, a1); a1 = fmaf(u[4 * j + 5], wb.y, a1);
                  a1 = fmaf(u[4 * j + 6], wb.z, a1); a1 = fmaf(u[4 * j + 7], wb.w, a1);
                }
                pt2[tp] += a0 + a1;
              }
            }
            }   // PAIR
            __syncwarp();
            // column domain: soft-argmax / max partials of ss_final (lane owns columns c0 + tc .. +3)
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
          if constexpr (!PAIR) {
            // ---- tap products on the tensor core: P[128 rows x 32 taps] = U[128 x 64] . Wt^T with the same three-term
            // split.  U goes back into columns 0..63 of this (already drained) accumulator slot as the TMEM A operand,
            // the products land in columns 64..127: [U_hi.W_hi + U_lo.W_hi | U_hi.W_lo].
            const uint32_t t_slot = tmem_base + (uint32_t)(slot * 128);
            const uint32_t t_lane = (uint32_t)(q * 32) << 16;
            constexpr uint32_t kDescHi128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            constexpr uint32_t idesc_t64 = make_idesc(64), idesc_t32 = make_idesc(32);
            tc_st32(t_slot + t_lane, uh);
            tc_st32(t_slot + t_lane + 32u, ul);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            const int nsets = p.tail_w2 ? 2 : 1;
            for (int set = 0; set < nsets; ++set) {
              asm volatile("bar.sync 1, 128;" ::: "memory");     // all four lane quarters of U are in place (set 1: D has been read)
              if (warp == 3 && lane == 0) {
                tc_fence_after();
                const uint32_t wb = ((smem_u32(tailw_s + set * CV_TAILW_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  tc_mma_f16_ts(t_slot + 64u, t_slot + 8u * ks, wb + 2u * ks, kDescHi128, idesc_t64, ks != 0);
                  tc_mma_f16_ts(t_slot + 64u, t_slot + 32u + 8u * ks, wb + 2u * ks, kDescHi128, idesc_t32, 1u);
                }
                tc_commit(&tail_done[slot]);
              }
              mbar_wait(&tail_done[slot], (tail_ph >> slot) & 1u);
              tail_ph ^= 1u << slot;
              tc_fence_after();
              uint32_t d0[32], d1[32];
              tc_ld32_nowait(t_slot + t_lane + 64u, d0);
              tc_ld32_nowait(t_slot + t_lane + 96u, d1);
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
          if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&acc_empty[slot]); else mbar_arrive(&acc_empty[slot]); }
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
          // one partial per (item, warp, row group): merged by ss_merge_kernel
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
      // row -> output voxel (fixed over z): flat plane row rr = t*128 + row -> (yp, xp)
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
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128 + c0), v0);
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * 128 + 64 + c0), v1);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stage + lane * EPI_STAGE_LD + j) =
                make_float4(__uint_as_float(v0[j]) + __uint_as_float(v1[j]), __uint_as_float(v0[j + 1]) + __uint_as_float(v1[j + 1]),
                            __uint_as_float(v0[j + 2]) + __uint_as_float(v1[j + 2]), __uint_as_float(v0[j + 3]) + __uint_as_float(v1[j + 3]));
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
        if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&acc_empty[slot]); else mbar_arrive(&acc_empty[slot]); }
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
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace umma
}  // namespace vxb

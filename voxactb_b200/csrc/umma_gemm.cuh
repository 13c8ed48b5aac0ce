// tcgen05 split-16-bit GEMM / implicit-GEMM convolution engine for sm_100a.
//
//   terms == 3:  D[M,N] (fp32, TMEM) = A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (A = A_hi + A_lo, B = B_hi + B_lo)
//   terms == 2:  D = A_hi*B_hi (kind::f16) + [A_lo8 | A_hi8]*[B_hi8 | B_lo8] (one kind::f8f6f4 MMA for both corrections)
//   terms == 1:  D = A_hi*B_hi
//
// Every fp32 operand is stored as two 16-bit planes (hi = fp16(x), lo = fp16(x - hi), planes16.cuh: 22 significant bits in
// the same 4 bytes per element as fp32; the containers are typed __nv_bfloat16 for historical reasons).  Three kind::f16
// MMAs per logical product give fp32-class accuracy (~2^-22 relative per product) at 1/3 of the 16-bit tensor peak, which
// is what the 1e-3 Q-value gate needs (single-pass TF32 / bf16 miss it).  In the f8c form the "lo"
// plane holds E4M3 bytes (64-column blocks) and the fp16 planes are pre-scaled so that all products share one scale
// (un-scaling by Epilogue::alpha_dev).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer: per k-block loads A_hi/A_lo [128 x 64] and B_hi/B_lo [NT x 64] tiles
//               (SWIZZLE_128B, K-major) into a STAGES-deep smem ring, signalled by mbarriers
//   warp 1      MMA issuer: one elected lane issues 12 tcgen05.mma (128 x NT x 16) per k-block into one
//               of two TMEM accumulator stages; tcgen05.commit releases smem slots / publishes the tile
//   warps 2-5   epilogue: tcgen05.ld the accumulator (lane = row), bias/activation/residual, store
//               fp32 and/or bf16 hi/lo planes; overlaps with the next tile's main loop
//
// The A operand is addressed per k-block as (column a_col[kb], row m0 + a_row[kb]) in one of two
// source tensors: with a_row = the flat offset of a convolution tap in a zero/replicate PADDED
// channels-last tensor this is an implicit-GEMM convolution (rows = flat padded voxel index, taps are
// plain row shifts); with a_row = 0 and a_col = 64*kb it is an ordinary GEMM.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "umma_host.cuh"

namespace vxb {
namespace umma {

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on the barrier phase.  A wait that lasts ~4 s of SM clock means a lost TMA / MMA completion:
// trap (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  long long t0 = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (!done && (spins & 0xFFFu) == 0xFFFu) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000ll) __trap();
    }
  }
}
// one non-blocking probe of the barrier phase (1 = the phase with this parity has completed)
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same MMA with the descriptors given as (low word, shared high word): the low word advances by bytes >> 4
__device__ __forceinline__ void tc_mma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}
// kind::f8f6f4 (E4M3 x E4M3, K = 32 bytes per instruction) with the same descriptor convention: the correction MMA of the
// f16 + fp8-corrected product (terms == 2, see conv_f8c.cuh for the arithmetic)
__device__ __forceinline__ void tc_mma_f8_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}
// A operand from tensor memory (row m = TMEM lane m, two 16-bit K elements per 32-bit column, 8 columns per K = 16
// step; cute SM100_MMA_F16BF16_TS), B from a shared-memory descriptor
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns written from registers (thread t -> TMEM lane lane_base + t)
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
        "r"(taddr)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets columns [col, col+32) of TMEM lane (lane_base + t)
// issue only: the caller batches several loads behind one tcgen05.wait::ld
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of 128 bytes,
// 8-row groups SBO bytes apart, version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset: 1 (16 B), as CUTLASS sets it for swizzled K-major
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                             // version = 1
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = NT
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | VXB_IDESC_AB_FORMAT | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// epilogue staging per warp: a 32 x 32 fp32 transpose tile (row pitch 36 floats: conflict-free 128-bit
// accesses in both domains) + per-row output metadata
constexpr int EPI_STAGE_LD = 36;
struct RowInfo {
  long long orow[32];   // output row of each accumulator row of the warp, -1 = not stored
  long long roff[32];   // residual row offset (elements)
  float rinv[32];       // per-row scale (1 / row_div)
};
constexpr int EPI_BYTES_PER_WARP = 32 * EPI_STAGE_LD * 4 + (int)sizeof(RowInfo);

// specialised epilogues: 8 warps (two per TMEM lane quarter, alternating 32-column chunks), 32 x 32 fp32 staging
// per warp with an XOR swizzle on the 16-byte column index (conflict-free 128-bit accesses in both domains)
constexpr int EPW_FAST = 8;
constexpr int EPI_FAST_BYTES_PER_WARP = 32 * 32 * 4;
__host__ __device__ constexpr int epi_warps(int epi) { return epi == 0 ? 4 : EPW_FAST; }
__host__ __device__ constexpr int gemm_threads(int epi) { return 64 + 32 * epi_warps(epi); }

template <int NT, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;   // 16 KB per plane
  static constexpr int B_BYTES = NT * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int EPI_BYTES = 4 * EPI_BYTES_PER_WARP;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024 /* alignment slack */;   // dynamic part
};

enum { EPIK_GENERIC = 0, EPIK_PLAIN = 1, EPIK_ROWMAX = 2, EPIK_EXP = 3, EPIK_GEGLU = 4, EPIK_PHASE = 5 };

template <int NT, int STAGES, int EPI>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap mapA0h, const __grid_constant__ CUtensorMap mapA0l,
                 const __grid_constant__ CUtensorMap mapA1h, const __grid_constant__ CUtensorMap mapA1l,
                 const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
                 const Params p) {
  using L = SmemLayout<NT, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * L::STAGE_BYTES);
  uint64_t* full_bar = bars;                   // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;         // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES;      // [2] MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + 2; // [2] epilogue -> MMA
  uint32_t* tmem_base_smem = (uint32_t*)(bars + 2 * STAGES + 4);
  // static: keeps the accesses in the shared window (LDS/STS)
  __shared__ __align__(16) uint8_t epi_smem[EPI == EPIK_GENERIC ? L::EPI_BYTES : EPW_FAST * EPI_FAST_BYTES_PER_WARP];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * NT <= 32) ? 32 : (2 * NT <= 64) ? 64 : (2 * NT <= 128) ? 128 : (2 * NT <= 256) ? 256 : 512;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0h); tma_prefetch_desc(&mapA0l);
    tma_prefetch_desc(&mapA1h); tma_prefetch_desc(&mapA1l);
    tma_prefetch_desc(&mapWh); tma_prefetch_desc(&mapWl);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], epi_warps(EPI)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int total_tiles = p.m_tiles * p.n_tiles * p.batches;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const KPlan pl = p.plan;
      const int tiles_per_z = p.m_tiles * p.n_tiles;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / tiles_per_z, tz = tile % tiles_per_z;
        const int zb = z / p.Hz, zh = z % p.Hz;
        const int mt = tz / p.n_tiles, nt = tz % p.n_tiles;
        int m0 = mt * BM + zb * p.a_row_zb + zh * p.a_row_zh + p.a_row_off;
        int n0 = nt * NT + zb * p.w_row_zb + zh * p.w_row_zh;
        int tap_col = 0, tap_rot = 0;
        if (p.tap_m) { m0 -= mt * BM; tap_col = p.tap_acol[mt]; n0 += p.tap_wrow[mt]; tap_rot = p.tap_rot[mt]; }
        const uint32_t km = p.kmask ? __ldg(p.kmask + nt) : 0xffffffffu;
        for (int kb0 = 0; kb0 < pl.num_kb; ++kb0) {
          int kb = kb0 + tap_rot;
          if (kb >= pl.num_kb) kb -= pl.num_kb;
          if (p.kmask && !((km >> kb) & 1u)) continue;             // all-zero weight block of this N tile
          int a_row = 0, a_col = kb * BK, a_src = 0;
          if (pl.taps) {
            const int tap = kb / pl.cpb, cb = kb - tap * pl.cpb;
            const int c = pl.taps >> 1;
            const int dx = tap % pl.taps, dy = (tap / pl.taps) % pl.taps, dz = tap / (pl.taps * pl.taps);
            a_row = ((dz - c) * pl.Vp + (dy - c)) * pl.Vp + (dx - c);
            a_src = cb >= pl.cb_src0;
            a_col = (a_src ? cb - pl.cb_src0 : cb) * BK;
          }
          a_col += zh * p.a_col_zh + (int)p.a_col_off + tap_col;
          int w_col = kb * BK + zh * p.w_col_zh + (int)p.w_col_off;
          int a_rowk = 0, w_rowk = 0;
          if (p.kblk_a) {            // K-blocked storage: the K block index selects a group of rows, columns 0..63 (offsets are multiples of 64)
            a_rowk = (a_col / BK) * p.kblk_a; a_col = 0;
            w_rowk = (w_col / BK) * p.kblk_w; w_col = 0;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + stage * L::STAGE_BYTES;
          const CUtensorMap* ah = a_src ? &mapA1h : &mapA0h;
          const CUtensorMap* al = a_src ? &mapA1l : &mapA0l;
          if (p.terms == 1) {
            mbar_expect_tx(&full_bar[stage], L::A_BYTES + L::B_BYTES);
            tma_load_2d(ah, &full_bar[stage], s, a_col, m0 + a_row + a_rowk);
            tma_load_2d(&mapWh, &full_bar[stage], s + 2 * L::A_BYTES, w_col, n0 + w_rowk);
          } else {
            mbar_expect_tx(&full_bar[stage], L::STAGE_BYTES);
            tma_load_2d(ah, &full_bar[stage], s, a_col, m0 + a_row + a_rowk);
            tma_load_2d(al, &full_bar[stage], s + L::A_BYTES, a_col, m0 + a_row + a_rowk);
            tma_load_2d(&mapWh, &full_bar[stage], s + 2 * L::A_BYTES, w_col, n0 + w_rowk);
            tma_load_2d(&mapWl, &full_bar[stage], s + 2 * L::A_BYTES + L::B_BYTES, w_col, n0 + w_rowk);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp runs the (warp-uniform) loops so that descriptors stay in uniform registers; one elected
    // lane issues the tcgen05 instructions (a single divergent lane makes ptxas emit a broadcast loop per MMA).
    {
      uint32_t leader;
      asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc(NT);
      constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO 1024, version 1, SWIZZLE_128B
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int num_kb = p.plan.num_kb;
      const bool three = p.terms == 3;
      // terms == 2: the "lo" tiles hold E4M3 bytes ([x_lo8 | x_hi8] against [w_hi8 | w_lo8] along K) and ONE fp8 MMA per 32 bytes
      // of K adds both correction terms to the same accumulator (operands pre-scaled so that all products share one scale)
      const bool f8c = p.terms == 2;
      constexpr uint32_t idesc8 = (1u << 4) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_u + (uint32_t)(acc * NT);
        uint32_t km = 0xffffffffu;
        int last_kb = num_kb - 1;
        if (p.kmask) {
          km = __ldg(p.kmask + (tile % (p.m_tiles * p.n_tiles)) % p.n_tiles);
          last_kb = 31 - __clz((int)km);
        }
        uint32_t started = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (p.kmask && !((km >> kb) & 1u)) continue;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint32_t a_hi = ((sa >> 4) & 0x3FFFu) | (1u << 16);
          const uint32_t a_lo = a_hi + (L::A_BYTES >> 4);
          const uint32_t b_hi = a_hi + ((2 * L::A_BYTES) >> 4);
          const uint32_t b_lo = b_hi + (L::B_BYTES >> 4);
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              tc_mma_bf16_lo(d_tmem, a_hi + 2 * k, b_hi + 2 * k, kDescHi, idesc, (started | k) != 0);
              if (three) {
                tc_mma_bf16_lo(d_tmem, a_hi + 2 * k, b_lo + 2 * k, kDescHi, idesc, 1);
                tc_mma_bf16_lo(d_tmem, a_lo + 2 * k, b_hi + 2 * k, kDescHi, idesc, 1);
              }
              if (f8c) tc_mma_f8_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, kDescHi, idesc8, 1);
            }
            tc_commit(&empty_bar[stage]);             // smem slot reusable once these MMAs retire
            if (kb == last_kb) tc_commit(&acc_full[acc]);   // accumulator complete -> epilogue
          }
          started = 1;
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    // Row domain: lane t owns accumulator row q*32 + t (tcgen05.ld 32x32b).  Values are transposed through a
    // per-warp shared-memory tile so that the stores (and the residual loads) are column-contiguous:
    // in the column domain lane l handles rows 4i + l/8 (i = 0..7) and columns 4(l%8)..+3 of a 32-column chunk.
    // EPI selects a compile-time specialisation (the generic body costs ~60 instructions per element):
    //   EPIK_PLAIN   ROWS_PLAIN store (bias / LeakyReLU / residual / row scale; fp32 and/or plane outputs)
    //   EPIK_ROWMAX  attention pass 1, EPIK_EXP attention pass 2, EPIK_GENERIC convolution row modes + V^T planes
    //   EPIK_PHASE   ROWS_PHASE store of the folded up-convolution (bias / LeakyReLU; fp32 rows or hi + lo / c8 planes of the
    //                fine grid).  With the generic body this GEMM was epilogue-bound: 28 us per 128 x 256 tile on four warps
    //                against 18 us of MMAs, so skipping a third of the K blocks (Params::kmask) changed nothing.
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const Epilogue& e = p.ep;
    const int ew = warp - 2;                        // epilogue warp index
    const int half = ew >> 2;                       // fast path: which 32-column chunks of a tile this warp takes
    float* stage = reinterpret_cast<float*>(epi_smem + (EPI == EPIK_GENERIC ? q * EPI_BYTES_PER_WARP : ew * EPI_FAST_BYTES_PER_WARP));
    const int tr = lane >> 3, tc = (lane & 7) * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    if constexpr (EPI != EPIK_GENERIC) {
      const float slope = e.act_slope >= 0.f ? e.act_slope : 1.f;   // max(x, x*slope) == LeakyReLU for slope in [0,1]
      [[maybe_unused]] const float alpha_f = e.alpha_dev ? e.alpha * __ldg(e.alpha_dev) : e.alpha;   // PLAIN / GEGLU: f8c un-scaling
      const bool f32_vec = e.out_f32 && ((e.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0) &&
                           ((p.c_zb & 3) == 0) && ((p.c_zh & 3) == 0);
      const bool res_vec = e.residual && ((e.ldr & 3) == 0) && ((reinterpret_cast<uintptr_t>(e.residual) & 15) == 0);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tiles_per_z = p.m_tiles * p.n_tiles;
        const int z = tile / tiles_per_z, tz = tile % tiles_per_z;
        const int zb = z / p.Hz, zh = z % p.Hz;
        const int mt = tz / p.n_tiles, nt = tz % p.n_tiles;
        const int mw = mt * BM + q * 32;              // first row of this warp
        const int m = mw + lane;                      // row-domain row of this lane
        const int n0 = nt * NT;
        const long long rs_off = zb * p.rs_zb + zh * p.rs_zh;
        const bool row_ok = m < e.M;
        float row_acc = (EPI == EPIK_ROWMAX) ? -INFINITY : 0.f;
        float row_sub = 0.f;
        if constexpr (EPI == EPIK_EXP) { if (row_ok) row_sub = e.row_sub[rs_off + m]; }
        float rinv_r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rinv_r[i] = 1.f;
        if constexpr (EPI == EPIK_PLAIN) {
          if (e.row_div) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = mw + i * 4 + tr;
              if (r < e.M) rinv_r[i] = 1.f / e.row_div[rs_off + r];
            }
          }
        }
        // PHASE: fine-grid row of phase 0 of each of this lane's eight column-domain rows (-1 = halo row, not stored)
        [[maybe_unused]] long long base_r[8];
        if constexpr (EPI == EPIK_PHASE) {
          const int Vp = e.Vp, V = Vp - 2 * e.pad, vp3 = Vp * Vp * Vp, s = e.phase_s, oV = e.out_Vp;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = mw + i * 4 + tr;
            const int mabs = r + zb * p.a_row_zb + p.a_row_off;   // flat row of the padded coarse grid (all batches)
            const int qb = mabs / vp3, rr = mabs - qb * vp3;
            const int qd = rr / (Vp * Vp) - e.pad, qh = (rr / Vp) % Vp - e.pad, qw = rr % Vp - e.pad;
            const bool ok = r < e.M && qd >= 0 && qd < V && qh >= 0 && qh < V && qw >= 0 && qw < V;
            base_r[i] = ok ? (((long long)qb * oV + qd * s + e.out_pad) * oV + qh * s + e.out_pad) * oV + qw * s + e.out_pad : -1;
          }
        }
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        if constexpr (EPI == EPIK_PHASE) {
          const int s = e.phase_s, oV = e.out_Vp;
          float* const o32 = e.out_f32 ? e.out_f32 + zb * p.c_zb + zh * p.c_zh : nullptr;
          __nv_bfloat16* const ohi = e.out_hi ? e.out_hi + zb * p.p_zb + zh * p.p_zh : nullptr;
          __nv_bfloat16* const olo = e.out_lo ? e.out_lo + zb * p.p_zb + zh * p.p_zh : nullptr;
          const float fa = e.f8a ? __ldg(e.f8a) : 0.f;
#pragma unroll 1
          for (int c0 = half * 32; c0 < NT; c0 += 64) {
            const int n = n0 + c0;
            if (n >= e.N) break;                       // warp-uniform (N is a multiple of 64: chunks are whole)
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                  make_float4(__uint_as_float(v[j]) * alpha_f, __uint_as_float(v[j + 1]) * alpha_f,
                              __uint_as_float(v[j + 2]) * alpha_f, __uint_as_float(v[j + 3]) * alpha_f);
            __syncwarp();
            // 64-column block = one polyphase: fine voxel = s * q + r in the (padded) fine grid
            const int ph = e.phase_perm ? (int)__ldg(e.phase_perm + (n >> 6)) : (n >> 6);
            const int rd = ph / (s * s), rh = (ph / s) % s, rw = ph % s;
            const long long poff = ((long long)rd * oV + rh) * oV + rw;
            const int ncol = (n & 63) + tc;
            float4 x4[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + tr;
              x4[i] = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias) bv = __ldg(reinterpret_cast<const float4*>(e.bias + ncol));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4& x = x4[i];
              x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
              x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);
              x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
            }
            if (o32) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (base_r[i] >= 0) *reinterpret_cast<float4*>(o32 + (base_r[i] + poff) * e.ldc + ncol) = x4[i];
            }
            if (ohi) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (base_r[i] < 0) continue;
                const float4 x = x4[i];
                const long long orow = base_r[i] + poff;
                const __nv_bfloat162 h01 = pl2_from_floats(x.x, x.y), h23 = pl2_from_floats(x.z, x.w);
                const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
                uint2 hv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(ohi + orow * e.ldp + ncol) = hv;
                if (e.f8a) {
                  // c8 plane (conv_f8c.cuh): bytes [32 x lo8 | 32 x hi8] per 32-channel block of the 64-channel row
                  uint8_t* rowb = reinterpret_cast<uint8_t*>(olo + orow * e.ldp) + (ncol >> 5) * 64 + (ncol & 31);
                  *reinterpret_cast<uint32_t*>(rowb) = pl_e4m3x4(x.x - f01.x, x.y - f01.y, x.z - f23.x, x.w - f23.y, fa * 2048.f);
                  *reinterpret_cast<uint32_t*>(rowb + 32) = pl_e4m3x4(x.x, x.y, x.z, x.w, fa);
                } else {
                  const __nv_bfloat162 l01 = pl2_from_floats(x.x - f01.x, x.y - f01.y);
                  const __nv_bfloat162 l23 = pl2_from_floats(x.z - f23.x, x.w - f23.y);
                  uint2 lv;
                  lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                  *reinterpret_cast<uint2*>(olo + orow * e.ldp + ncol) = lv;
                }
              }
            }
            __syncwarp();
          }
        } else if constexpr (EPI == EPIK_GEGLU) {
          // pairs of 32-column chunks: [a | gate] of the same 32 output columns; the warps of a lane quarter alternate pairs
#pragma unroll 1
          for (int c0 = half * 64; c0 < NT; c0 += 128) {
            const int n = n0 + c0;
            if (n >= e.N) break;
            float4 xa[8], xg[8];
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              uint32_t v[32];
              tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0 + part * 32), v);
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                    make_float4(__uint_as_float(v[j]) * alpha_f, __uint_as_float(v[j + 1]) * alpha_f,
                                __uint_as_float(v[j + 2]) * alpha_f, __uint_as_float(v[j + 3]) * alpha_f);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + tr;
                const float4 x = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
                if (part == 0) xa[i] = x; else xg[i] = x;
              }
              __syncwarp();
            }
            const int nn = n + tc;
            float ba[4] = {0.f, 0.f, 0.f, 0.f}, bg[4] = {0.f, 0.f, 0.f, 0.f};
            if (e.bias) {
#pragma unroll
              for (int t = 0; t < 4; ++t) { ba[t] = __ldg(e.bias + nn + t); bg[t] = __ldg(e.bias + nn + 32 + t); }
            }
            const long long poff = zb * p.p_zb + zh * p.p_zh + (long long)(mw + tr) * e.ldp + (n >> 1) + tc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (mw + i * 4 + tr < e.M) {
                const float a[4] = {xa[i].x + ba[0], xa[i].y + ba[1], xa[i].z + ba[2], xa[i].w + ba[3]};
                const float g[4] = {xg[i].x + bg[0], xg[i].y + bg[1], xg[i].z + bg[2], xg[i].w + bg[3]};
                float o[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) o[t] = a[t] * (0.5f * g[t] * (1.f + erff(g[t] * 0.70710678118654752f)));
                const __nv_bfloat162 h01 = pl2_from_floats(o[0], o[1]), h23 = pl2_from_floats(o[2], o[3]);
                const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
                const __nv_bfloat162 l01 = pl2_from_floats(o[0] - f01.x, o[1] - f01.y);
                const __nv_bfloat162 l23 = pl2_from_floats(o[2] - f23.x, o[3] - f23.y);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(e.out_hi + poff + (long long)(i * 4) * e.ldp) = hv;
                *reinterpret_cast<uint2*>(e.out_lo + poff + (long long)(i * 4) * e.ldp) = lv;
              }
            }
          }
        } else {
#pragma unroll 1
        for (int c0 = half * 32; c0 < NT; c0 += 64) {
          const int n = n0 + c0;
          if (n >= e.N) break;                         // warp-uniform
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0), v);
          const bool full = n + 31 < e.N;
          if constexpr (EPI == EPIK_ROWMAX) {
            float mx = -INFINITY;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (n + j < e.N) mx = fmaxf(mx, __uint_as_float(v[j]));
            }
            row_acc = fmaxf(row_acc, mx * e.alpha);    // alpha > 0
          } else {
            if constexpr (EPI == EPIK_EXP) {
              if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float t[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[u]) : "f"(fmaf(__uint_as_float(v[j + u]), e.alpha, VXB_P_EXP_BIAS - row_sub)));
                    row_acc += t[u];
                  }
                  *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_float4(t[0], t[1], t[2], t[3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float t[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[u]) : "f"(fmaf(__uint_as_float(v[j + u]), e.alpha, VXB_P_EXP_BIAS - row_sub)));
                    t[u] = (n + j + u < e.N) ? t[u] : 0.f;
                    row_acc += t[u];
                  }
                  *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) = make_float4(t[0], t[1], t[2], t[3]);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(stage + lane * 32 + (((j >> 2) ^ (lane & 7)) << 2)) =
                    make_float4(__uint_as_float(v[j]) * alpha_f, __uint_as_float(v[j + 1]) * alpha_f,
                                __uint_as_float(v[j + 2]) * alpha_f, __uint_as_float(v[j + 3]) * alpha_f);
            }
            __syncwarp();
            // ---- column domain
            const int nn = n + tc;
            const int nv = e.N - nn;
            if (nv > 0) {
              float4 x4[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + tr;
                x4[i] = *reinterpret_cast<const float4*>(stage + r * 32 + ((((lane & 7)) ^ (r & 7)) << 2));
              }
              if constexpr (EPI == EPIK_PLAIN) {
                float bv[4] = {0.f, 0.f, 0.f, 0.f};
                if (e.bias) {
#pragma unroll
                  for (int t = 0; t < 4; ++t) if (t < nv) bv[t] = __ldg(e.bias + nn + t);
                }
                float4 rv[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e.residual) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const int r = mw + i * 4 + tr;
                    if (r < e.M) {
                      const float* rp = e.residual + (long long)(r % e.res_rows) * e.ldr + nn;
                      if (res_vec && nv >= 4) {
                        rv[i] = *reinterpret_cast<const float4*>(rp);
                      } else {
                        rv[i].x = rp[0];
                        if (nv > 1) rv[i].y = rp[1];
                        if (nv > 2) rv[i].z = rp[2];
                        if (nv > 3) rv[i].w = rp[3];
                      }
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float4& x = x4[i];
                  x.x += bv[0]; x.y += bv[1]; x.z += bv[2]; x.w += bv[3];
                  x.x = fmaxf(x.x, x.x * slope); x.y = fmaxf(x.y, x.y * slope);
                  x.z = fmaxf(x.z, x.z * slope); x.w = fmaxf(x.w, x.w * slope);
                  x.x = (x.x + rv[i].x) * rinv_r[i]; x.y = (x.y + rv[i].y) * rinv_r[i];
                  x.z = (x.z + rv[i].z) * rinv_r[i]; x.w = (x.w + rv[i].w) * rinv_r[i];
                }
                if (e.out_f32) {
                  float* dst0 = e.out_f32 + zb * p.c_zb + zh * p.c_zh + (long long)(mw + tr) * e.ldc + nn;
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (mw + i * 4 + tr < e.M) {
                      float* dst = dst0 + (long long)(i * 4) * e.ldc;
                      if (f32_vec && nv >= 4) {
                        *reinterpret_cast<float4*>(dst) = x4[i];
                      } else {
                        dst[0] = x4[i].x;
                        if (nv > 1) dst[1] = x4[i].y;
                        if (nv > 2) dst[2] = x4[i].z;
                        if (nv > 3) dst[3] = x4[i].w;
                      }
                    }
                  }
                }
              }
              if (e.out_hi) {
                const long long poff = zb * p.p_zb + zh * p.p_zh + (long long)(mw + tr) * e.ldp + nn;
                // EXP: values beyond N are exact zeros and rows are padded to ld (a multiple of 8): always 4-wide
                const bool wide = (EPI == EPIK_EXP) || nv >= 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  if (mw + i * 4 + tr < e.M) {
                    const float4 x = x4[i];
                    const __nv_bfloat162 h01 = pl2_from_floats(x.x, x.y), h23 = pl2_from_floats(x.z, x.w);
                    const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
                    const __nv_bfloat162 l01 = pl2_from_floats(x.x - f01.x, x.y - f01.y);
                    const __nv_bfloat162 l23 = pl2_from_floats(x.z - f23.x, x.w - f23.y);
                    __nv_bfloat16* dh = e.out_hi + poff + (long long)(i * 4) * e.ldp;
                    __nv_bfloat16* dl = e.out_lo + poff + (long long)(i * 4) * e.ldp;
                    if (wide) {
                      uint2 hv, lv;
                      hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                      lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                      *reinterpret_cast<uint2*>(dh) = hv;
                      *reinterpret_cast<uint2*>(dl) = lv;
                    } else {
                      dh[0] = h01.x; dl[0] = l01.x;
                      if (nv > 1) { dh[1] = h01.y; dl[1] = l01.y; }
                      if (nv > 2) { dh[2] = h23.x; dl[2] = l23.x; }
                    }
                  }
                }
              }
            }
            __syncwarp();
          }
        }
        }   // !GEGLU
        if constexpr (EPI == EPIK_ROWMAX) {
          if (row_ok) {
            float* a = e.row_stat + rs_off + m;
            if (row_acc >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(row_acc));
            else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(row_acc));
          }
        }
        if constexpr (EPI == EPIK_EXP) { if (row_ok) atomicAdd(e.row_stat + rs_off + m, row_acc); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
    RowInfo* ri = reinterpret_cast<RowInfo*>(stage + 32 * EPI_STAGE_LD);
    const float alpha_g = e.alpha_dev ? e.alpha * __ldg(e.alpha_dev) : e.alpha;   // device-side factor (f8c operand un-scaling)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tiles_per_z = p.m_tiles * p.n_tiles;
      const int z = tile / tiles_per_z, tz = tile % tiles_per_z;
      const int zb = z / p.Hz, zh = z % p.Hz;
      const int mt = tz / p.n_tiles, nt = tz % p.n_tiles;
      const int m = mt * BM + q * 32 + lane;        // this thread's row (within the batch entry)
      const int n0 = nt * NT;
      float* const out_f32 = e.out_f32 ? e.out_f32 + zb * p.c_zb + zh * p.c_zh : nullptr;
      __nv_bfloat16* const out_hi = e.out_hi ? e.out_hi + zb * p.p_zb + zh * p.p_zh : nullptr;
      __nv_bfloat16* const out_lo = e.out_lo ? e.out_lo + zb * p.p_zb + zh * p.p_zh : nullptr;
      // ---- row mapping
      bool row_ok = m < e.M;
      long long orow = m;
      int qd = 0, qh = 0, qw = 0, qb = 0;
      if (e.row_mode != ROWS_PLAIN) {
        const int Vp = e.Vp, V = Vp - 2 * e.pad;
        const int vp3 = Vp * Vp * Vp;
        const int mabs = m + zb * p.a_row_zb + p.a_row_off;   // flat row of the padded grid (all batches)
        qb = mabs / vp3;
        const int r = mabs - qb * vp3;
        qd = r / (Vp * Vp) - e.pad; qh = (r / Vp) % Vp - e.pad; qw = r % Vp - e.pad;
        row_ok = row_ok && qd >= 0 && qd < V && qh >= 0 && qh < V && qw >= 0 && qw < V;
        if (e.row_mode == ROWS_CONV_FLAT && !e.out_padded) orow = (((long long)qb * V + qd) * V + qh) * V + qw;
      }
      const long long rs_off = zb * p.rs_zb + zh * p.rs_zh;
      float row_acc = (e.mode == EPI_ROWMAX) ? -INFINITY : 0.f;
      float row_sub = 0.f, row_inv = 1.f;
      if (row_ok && e.mode == EPI_EXP) row_sub = e.row_sub[rs_off + m];
      if (row_ok && e.mode == EPI_STORE && e.row_div) row_inv = 1.f / e.row_div[rs_off + m];
      ri->orow[lane] = row_ok ? orow : -1;
      ri->roff[lane] = e.residual ? (long long)(m % e.res_rows) * e.ldr : 0;
      ri->rinv[lane] = row_inv;
      __syncwarp();
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < NT; c0 += 32) {
        const int n = n0 + c0;
        if (n >= e.N) break;                         // warp-uniform
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NT + c0), v);
        if (e.mode == EPI_ROWMAX) {
          if (n + 31 < e.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) row_acc = fmaxf(row_acc, __uint_as_float(v[j]) * e.alpha);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < e.N) row_acc = fmaxf(row_acc, __uint_as_float(v[j]) * e.alpha);
          }
          continue;
        }
        float f[32];
        if (e.mode == EPI_EXP) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float t;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(__uint_as_float(v[j]), e.alpha, VXB_P_EXP_BIAS - row_sub)));
            t = (n + j < e.N) ? t : 0.f;
            row_acc += t;
            f[j] = t;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha_g;
        }
        if (e.transpose_planes) {
          // element (row, col) -> plane[col * ldp + row]: lanes are consecutive rows, already contiguous
          if (row_ok) {
#pragma unroll 4
            for (int j = 0; j < 32; ++j) {
              if (n + j < e.N) {
                float t = f[j];
                if (e.bias) t += __ldg(e.bias + n + j);
                if (e.act_slope >= 0.f) t = t > 0.f ? t : t * e.act_slope;
                const __nv_bfloat16 hi = pl_from_float(t);
                out_hi[(long long)(n + j) * e.ldp + orow] = hi;
                out_lo[(long long)(n + j) * e.ldp + orow] = pl_from_float(t - pl_to_float(hi));
              }
            }
          }
          continue;
        }
        int ncol0 = n;                                // column of the chunk inside the output row
        if (e.row_mode == ROWS_PHASE) {
          // 64-column block = one polyphase: fine voxel = s*q + r in the (padded) fine grid
          const int ph = e.phase_perm ? (int)__ldg(e.phase_perm + n / 64) : n / 64;
          const int s = e.phase_s;
          const int rd = ph / (s * s), rh = (ph / s) % s, rw = ph % s;
          const int oV = e.out_Vp;
          orow = (((long long)qb * oV + qd * s + rd + e.out_pad) * oV + qh * s + rh + e.out_pad) * oV + qw * s + rw + e.out_pad;
          ri->orow[lane] = row_ok ? orow : -1;
          ncol0 = n % 64;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * EPI_STAGE_LD + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        __syncwarp();
        // ---- column domain: all shared-memory reads first (independent), then math, then predicated stores
        const int nn = n + tc;                        // absolute GEMM column of this lane's first element
        const int nv = e.N - nn;                      // valid columns from nn (>= 4: all four)
        if (nv > 0) {
          const int ncol = ncol0 + tc;
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (e.bias) {
            const float* bp = e.bias + ((e.row_mode == ROWS_PHASE) ? ncol : nn);
#pragma unroll
            for (int t = 0; t < 4; ++t) if (t < nv) bv[t] = __ldg(bp + t);
          }
          long long orow_r[8];
          float rinv_r[8];
          float4 x4[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + tr;
            orow_r[i] = ri->orow[r];
            rinv_r[i] = ri->rinv[r];
            x4[i] = *reinterpret_cast<const float4*>(stage + r * EPI_STAGE_LD + tc);
          }
          float4 rv[8];
          if (e.residual) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (orow_r[i] >= 0) {
                const float* rp = e.residual + ri->roff[i * 4 + tr] + nn;
                if (nv >= 4 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
                  rv[i] = *reinterpret_cast<const float4*>(rp);
                } else {
                  rv[i].x = rp[0];
                  if (nv > 1) rv[i].y = rp[1];
                  if (nv > 2) rv[i].z = rp[2];
                  if (nv > 3) rv[i].w = rp[3];
                }
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x[4] = {x4[i].x + bv[0], x4[i].y + bv[1], x4[i].z + bv[2], x4[i].w + bv[3]};
            if (e.act_slope >= 0.f) {
#pragma unroll
              for (int t = 0; t < 4; ++t) x[t] = x[t] > 0.f ? x[t] : x[t] * e.act_slope;
            }
            if (e.residual) { x[0] += rv[i].x; x[1] += rv[i].y; x[2] += rv[i].z; x[3] += rv[i].w; }
#pragma unroll
            for (int t = 0; t < 4; ++t) x[t] *= rinv_r[i];
            const bool ok = orow_r[i] >= 0;
            if (out_f32 && ok) {
              float* dst = out_f32 + orow_r[i] * e.ldc + ncol;
              if (nv >= 4 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
              } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) if (t < nv) dst[t] = x[t];
              }
            }
            if (out_hi && ok) {
              __nv_bfloat16* dh = out_hi + orow_r[i] * e.ldp + ncol;
              __nv_bfloat16* dl = out_lo + orow_r[i] * e.ldp + ncol;
              const __nv_bfloat162 h01 = pl2_from_floats(x[0], x[1]), h23 = pl2_from_floats(x[2], x[3]);
              const float2 f01 = pl2_to_float2(h01), f23 = pl2_to_float2(h23);
              const __nv_bfloat162 l01 = pl2_from_floats(x[0] - f01.x, x[1] - f01.y);
              const __nv_bfloat162 l23 = pl2_from_floats(x[2] - f23.x, x[3] - f23.y);
              if (e.f8a) {
                // c8 plane (conv_f8c.cuh): bytes [32 x lo8 | 32 x hi8] per 32-channel block of the 64-channel row
                const float fa = __ldg(e.f8a);
                uint8_t* rowb = reinterpret_cast<uint8_t*>(out_lo + orow_r[i] * e.ldp) + (ncol >> 5) * 64 + (ncol & 31);
                uint2 hv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                *reinterpret_cast<uint2*>(dh) = hv;
                *reinterpret_cast<uint32_t*>(rowb) = pl_e4m3x4(x[0] - f01.x, x[1] - f01.y, x[2] - f23.x, x[3] - f23.y, fa * 2048.f);
                *reinterpret_cast<uint32_t*>(rowb + 32) = pl_e4m3x4(x[0], x[1], x[2], x[3], fa);
              } else
              // EXP mode: values beyond N are exact zeros and the row is padded to ld (a multiple of 8)
              if ((nv >= 4 || e.mode == EPI_EXP) && ((reinterpret_cast<uintptr_t>(dh) & 7) == 0)) {
                uint2 hv, lv;
                hv.x = *reinterpret_cast<const uint32_t*>(&h01); hv.y = *reinterpret_cast<const uint32_t*>(&h23);
                lv.x = *reinterpret_cast<const uint32_t*>(&l01); lv.y = *reinterpret_cast<const uint32_t*>(&l23);
                *reinterpret_cast<uint2*>(dh) = hv;
                *reinterpret_cast<uint2*>(dl) = lv;
              } else {
                const __nv_bfloat16 hh[4] = {h01.x, h01.y, h23.x, h23.y}, ll[4] = {l01.x, l01.y, l23.x, l23.y};
#pragma unroll
                for (int t = 0; t < 4; ++t) if (t < nv) { dh[t] = hh[t]; dl[t] = ll[t]; }
              }
            }
          }
        }
        __syncwarp();
      }
      if (row_ok && e.mode == EPI_ROWMAX) {
        float* a = e.row_stat + rs_off + m;
        if (row_acc >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(row_acc));
        else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(row_acc));
      }
      if (row_ok && e.mode == EPI_EXP) atomicAdd(e.row_stat + rs_off + m, row_acc);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace umma
}  // namespace vxb

// Fused attention on tcgen05: softmax(scale q k^T) v per (batch, head) without materialising the probabilities.
//
// Pass 1 (the GEMM engine with the EPIK_ROWMAX epilogue) leaves the row maxima of the log2-domain scores;
// this kernel is pass 2.  A CTA owns 128 queries of one (batch, head) and walks the keys in tiles of 64:
//   S  = Q K^T           tcgen05, split-16 x3 as  S[:,0:128] = Q_hi [K_hi;K_lo]^T,  S[:,0:64] += Q_lo K_hi^T   (TMEM, 2 buffers)
//   P  = 2^(a S - max)   softmax warps: tcgen05.ld (both column halves summed on the way in, the S buffer is handed
//                        back at once), ex2 / row sums / hi-lo split on register pairs (FADD2 / FFMA2), then
//                        tcgen05.st into tensor memory: lane = query row, two keys per 32-bit column (TMEM, 2 buffers)
//   O += P V             tcgen05 with the A operand read from TMEM:  O[:,0:128] += P_hi [V_hi;V_lo],  O[:,0:64] += P_lo V_hi
// and finally O (both halves added) / row sum is written as hi/lo planes for the output projection.
// Warps: 0 = TMA producer (Q once per item, K / V^T tiles through a 4-stage ring), 1 = QK^T issuer, 2 = PV issuer
// (two issuing threads: S runs ahead as far as its two buffers allow, PV follows the softmax), 3-10 = softmax +
// epilogue in two groups that take alternate key tiles (group = S / P buffer = 32-channel half of O it writes out).
#pragma once
#include <type_traits>
#include "umma_gemm.cuh"

namespace vxb {
namespace umma {

constexpr int FA_THREADS = 352;               // TMA warp, QK^T issuer, PV issuer, 8 softmax warps (two per TMEM lane quarter)
constexpr int FA_KT = 64;                        // keys per tile
constexpr int FA_KVSTAGES = 4;                   // K/V tiles between the PV issuer (behind) and the TMA prefetch (ahead)
constexpr int FA_QBYTES = 2 * 128 * 128;         // Q hi + lo
constexpr int FA_KVBYTES = 4 * FA_KT * 128;      // K hi, K lo, Vt hi, Vt lo (64 rows x 128 B each)
constexpr int FA_SMEM = FA_QBYTES + FA_KVSTAGES * FA_KVBYTES + 1024;   // P lives in tensor memory
// TMEM columns: S0 @0, S1 @128 (hi*hi+lo*hi | hi*lo), O @256, P0 @384 (hi 32 cols, lo 32 cols), P1 @448
constexpr uint32_t FA_TMEM_O = 256, FA_TMEM_P = 384;

struct FlashParams {
  int B, H, Nq, Nk, dh;
  int q_batched;            // 0: one Q shared by all batches (first cross-attention iteration)
  int q_tiles, k_tiles, items;
  float alpha;              // scale * log2(e)
  const float* rowmax;      // [B*H*Nq] log2-domain row maxima (stabiliser; cancels in P / sum P)
  __nv_bfloat16* out_hi;    // O planes [B*Nq, ldo]
  __nv_bfloat16* out_lo;
  long long ldo;
  // train-mode dropout on the probabilities (0 = off): P_ij is multiplied by keep_ij / (1 - p) AFTER the row sum, exactly
  // softmax -> dropout -> @ v of the reference (perceiver_lang_io.py:124-130); mask = dropout_keep(seed, row * ld + j)
  unsigned int drop_thresh;
  float drop_inv_keep;
  unsigned long long drop_seed;
  int drop_ld;
};

__global__ void __launch_bounds__(FA_THREADS, 1)
flash_attn_kernel(const __grid_constant__ CUtensorMap mapQh, const __grid_constant__ CUtensorMap mapQl,
                  const __grid_constant__ CUtensorMap mapKh, const __grid_constant__ CUtensorMap mapKl,
                  const __grid_constant__ CUtensorMap mapVh, const __grid_constant__ CUtensorMap mapVl,
                  const FlashParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 + 2 * FA_KVSTAGES + 8 + 2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float rs_part[2][128];              // row-sum partials of the two key halves
  uint8_t* q_s = smem;
  uint8_t* kv_s = smem + FA_QBYTES;
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + FA_KVSTAGES;
  uint64_t* s_full = kv_empty + FA_KVSTAGES;
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* p_empty = p_full + 2;
  uint64_t* o_full = p_empty + 2;
  uint64_t* o_empty = o_full + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQh); tma_prefetch_desc(&mapQl);
    tma_prefetch_desc(&mapKh); tma_prefetch_desc(&mapKl);
    tma_prefetch_desc(&mapVh); tma_prefetch_desc(&mapVl);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < FA_KVSTAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
      mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
    }
    mbar_init(o_full, 1); mbar_init(o_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int inner = p.H * p.dh;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0, n = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
        const int qt = item % p.q_tiles, bh = item / p.q_tiles;
        const int h = bh % p.H, b = bh / p.H;
        mbar_wait(q_empty, (n & 1) ^ 1);
        mbar_expect_tx(q_full, FA_QBYTES);
        const int qrow = (p.q_batched ? b * p.Nq : 0) + qt * 128;
        tma_load_2d(&mapQh, q_full, q_s, h * p.dh, qrow);
        tma_load_2d(&mapQl, q_full, q_s + 128 * 128, h * p.dh, qrow);
        for (int j = 0; j < p.k_tiles; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          uint8_t* s = kv_s + st * FA_KVBYTES;
          mbar_expect_tx(&kv_full[st], FA_KVBYTES);
          tma_load_2d(&mapKh, &kv_full[st], s, h * p.dh, b * p.Nk + j * FA_KT);
          tma_load_2d(&mapKl, &kv_full[st], s + FA_KT * 128, h * p.dh, b * p.Nk + j * FA_KT);
          tma_load_2d(&mapVh, &kv_full[st], s + 2 * FA_KT * 128, j * FA_KT, bh * p.dh);
          tma_load_2d(&mapVl, &kv_full[st], s + 3 * FA_KT * 128, j * FA_KT, bh * p.dh);
          if (++st == FA_KVSTAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== QK^T issuer (uniform loops, one elected lane issues)
    // Two issuing warps: S = Q K^T runs ahead as far as the two S buffers allow, independently of the PV warp's waits
    // for the softmax (a single in-order issuer coupled QK^T(j) to the softmax of tile j-3).
    uint32_t leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc128 = make_idesc(128), idesc64 = make_idesc(64);
    constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    auto dlo = [](uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); };
    const uint32_t q_hi = dlo(smem_u32(q_s)), q_lo = q_hi + ((128 * 128) >> 4);
    int st = 0;             // kv stage of the tile whose QK^T is issued next
    uint32_t ph = 0;
    uint32_t it = 0;        // global key-tile counter (S double buffer)
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      mbar_wait(q_full, n & 1);
      tc_fence_after();
      for (int j = 0; j < p.k_tiles; ++j, ++it) {
        // ---- S(j) = Q K_j^T into S buffer (it & 1)
        const uint32_t sb = it & 1u;
        mbar_wait(&kv_full[st], ph);
        mbar_wait(&s_empty[sb], ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t kb = dlo(smem_u32(kv_s + st * FA_KVBYTES));
        const uint32_t d_s = tmem_u + sb * 128u;
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            tc_mma_bf16_lo(d_s, q_hi + 2 * ks, kb + 2 * ks, kDescHi, idesc128, ks != 0);   // hi*hi | hi*lo
            tc_mma_bf16_lo(d_s, q_lo + 2 * ks, kb + 2 * ks, kDescHi, idesc64, 1);          // lo*hi
          }
          tc_commit(&s_full[sb]);
          if (j == p.k_tiles - 1) tc_commit(q_empty);          // Q tile free once the last QK^T retires
        }
        __syncwarp();
        if (++st == FA_KVSTAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================================================== PV issuer: O += P(j) V_j as soon as the softmax hands P(j) over
    uint32_t leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc128 = make_idesc(128), idesc64 = make_idesc(64);
    constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    auto dlo = [](uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); };
    int st_pv = 0;          // kv stage of the tile whose PV is issued next
    uint32_t itp = 0;       // global key-tile counter (P double buffer)
    uint32_t n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      for (int jt = 0; jt < p.k_tiles; ++jt, ++itp) {
        const uint32_t pb = itp & 1u;
        if (jt == 0) {
          mbar_wait(o_empty, (n & 1) ^ 1);                     // previous item's O has been read out
          tc_fence_after();
        }
        // p_full(j) implies s_full(j) (the softmax read S(j)), i.e. K_j / V_j have landed and QK^T(j) has retired:
        // committing kv_empty after these MMAs releases the stage for both uses
        mbar_wait(&p_full[pb], (itp >> 1) & 1u);
        tc_fence_after();
        const uint32_t p_hi = tmem_u + FA_TMEM_P + pb * 64u, p_lo = p_hi + 32u;   // A operand in TMEM: 8 columns per K = 16
        const uint32_t vb = dlo(smem_u32(kv_s + st_pv * FA_KVBYTES + 2 * FA_KT * 128));
        const uint32_t d_o = tmem_u + FA_TMEM_O;
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < FA_KT / 16; ++ks) {
            tc_mma_f16_ts(d_o, p_hi + 8 * ks, vb + 2 * ks, kDescHi, idesc128, (jt > 0 || ks != 0));   // hi*hi | hi*lo
            tc_mma_f16_ts(d_o, p_lo + 8 * ks, vb + 2 * ks, kDescHi, idesc64, 1);                      // lo*hi
          }
          tc_commit(&p_empty[pb]);
          tc_commit(&kv_empty[st_pv]);
          if (jt == p.k_tiles - 1) tc_commit(o_full);
        }
        __syncwarp();
        if (++st_pv == FA_KVSTAGES) st_pv = 0;
      }
    }
  } else {
    // ===================================================== softmax + epilogue (thread = query row)
    const int q = warp & 3;
    const int c = (warp - 3) >> 2;                             // softmax group = S/P buffer it owns = 32-channel half of O it writes
    const int r = q * 32 + lane;                               // row inside the 128-query tile
    uint32_t it = 0, n = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++n) {
      const int qt = item % p.q_tiles, bh = item / p.q_tiles;
      const int h = bh % p.H, b = bh / p.H;
      const int qi = qt * 128 + r;
      const bool row_ok = qi < p.Nq;
      const float m = row_ok ? p.rowmax[(size_t)bh * p.Nq + qi] : 0.f;
      const float bias = VXB_P_EXP_BIAS - m;
      float2 rs2 = make_float2(0.f, 0.f);                      // row sum as two interleaved partials (packed adds)
      for (int j = 0; j < p.k_tiles; ++j, ++it) {
        const uint32_t sb = it & 1u;
        if (sb != (uint32_t)c) continue;                       // the other group's tile
        mbar_wait(&s_full[sb], (it >> 1) & 1u);
        tc_fence_after();
        uint32_t phv[32], plv[32];                             // this row's P as packed fp16 pairs: hi and lo planes, 64 keys
        const int key0 = j * FA_KT;
        // S(tile) -> registers: both 32-key halves, hi*hi|lo*hi columns + hi*lo columns summed on the way in, then the
        // S buffer goes straight back to the MMA warp (QK^T of this group's next tile does not wait for the exps)
        float2 sv2[2][16];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v0[32], v1[32];
          tc_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + sb * 128u + (uint32_t)(cc * 32), v0);
          tc_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + sb * 128u + 64u + (uint32_t)(cc * 32), v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int e = 0; e < 16; ++e)
            sv2[cc][e] = __fadd2_rn(make_float2(__uint_as_float(v0[2 * e]), __uint_as_float(v0[2 * e + 1])),
                                    make_float2(__uint_as_float(v1[2 * e]), __uint_as_float(v1[2 * e + 1])));
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[sb]);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          // The softmax warps are issue-bound (two warps per scheduler): everything runs on register PAIRS with the
          // packed fp32 instructions (FADD2 / FFMA2), and the key mask only exists in the ragged last tile.
          const float2 alpha2 = make_float2(p.alpha, p.alpha), bias2 = make_float2(bias, bias), neg1 = make_float2(-1.f, -1.f);
          const unsigned long long drop_row = ((unsigned long long)bh * p.Nq + (unsigned long long)qi) * (unsigned long long)p.drop_ld;
          auto chunk = [&](auto masked_tag, auto drop_tag) {
            constexpr bool kMasked = decltype(masked_tag)::value;
            constexpr bool kDrop = decltype(drop_tag)::value;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t* hp = phv + cc * 16 + g * 4;
              uint32_t* lp = plv + cc * 16 + g * 4;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int jj = g * 8 + 2 * e;
                const float2 ar = __ffma2_rn(sv2[cc][g * 4 + e], alpha2, bias2);
                float2 t;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(ar.x));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(ar.y));
                if constexpr (kMasked) {
                  t.x = (key0 + cc * 32 + jj < p.Nk) ? t.x : 0.f;
                  t.y = (key0 + cc * 32 + jj + 1 < p.Nk) ? t.y : 0.f;
                }
                rs2 = __fadd2_rn(rs2, t);
                if constexpr (kDrop) {
                  const unsigned long long kc = drop_row + (unsigned long long)(key0 + cc * 32 + jj);
                  t.x = dropout_keep(p.drop_seed, kc, p.drop_thresh) ? t.x * p.drop_inv_keep : 0.f;
                  t.y = dropout_keep(p.drop_seed, kc + 1ull, p.drop_thresh) ? t.y * p.drop_inv_keep : 0.f;
                }
                const __nv_bfloat162 hh = pl2_from_floats(t.x, t.y);
                const float2 lo2 = __ffma2_rn(pl2_to_float2(hh), neg1, t);      // t - float(hi), exact
                const __nv_bfloat162 ll = pl2_from_floats(lo2.x, lo2.y);
                hp[e] = *reinterpret_cast<const uint32_t*>(&hh);
                lp[e] = *reinterpret_cast<const uint32_t*>(&ll);
              }
            }
          };
          if (p.drop_thresh) {                                   // training forward only (uniform branch)
            if (key0 + cc * 32 + 31 < p.Nk) chunk(std::false_type{}, std::true_type{});
            else chunk(std::true_type{}, std::true_type{});
          } else {
            if (key0 + cc * 32 + 31 < p.Nk) chunk(std::false_type{}, std::false_type{});
            else chunk(std::true_type{}, std::false_type{});
          }
        }
        // P -> tensor memory (the A operand of the PV MMAs): lane = query row, two keys per 32-bit column
        mbar_wait(&p_empty[sb], ((it >> 1) & 1u) ^ 1u);        // PV of this group's previous tile has consumed the P buffer
        tc_fence_after();
        tc_st32(tmem_base + ((uint32_t)(q * 32) << 16) + FA_TMEM_P + sb * 64u, phv);
        tc_st32(tmem_base + ((uint32_t)(q * 32) << 16) + FA_TMEM_P + sb * 64u + 32u, plv);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[sb]);
      }
      // ---- O / row_sum -> planes
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      // the two key halves of a row live in different warps: exchange the partial sums through shared memory
      rs_part[c][r] = rs2.x + rs2.y;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float inv = 1.f / (rs_part[0][r] + rs_part[1][r]);
      __nv_bfloat16* oh = p.out_hi + ((size_t)b * p.Nq + qi) * p.ldo + (size_t)h * p.dh;
      __nv_bfloat16* ol = p.out_lo + ((size_t)b * p.Nq + qi) * p.ldo + (size_t)h * p.dh;
      {
        uint32_t v0[32], v1[32];
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + FA_TMEM_O + (uint32_t)(c * 32), v0);
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + FA_TMEM_O + 64u + (uint32_t)(c * 32), v1);
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 hv, lv;
            uint32_t* hp = reinterpret_cast<uint32_t*>(&hv);
            uint32_t* lp = reinterpret_cast<uint32_t*>(&lv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int jj = g * 8 + 2 * e;
              const float a0 = (__uint_as_float(v0[jj]) + __uint_as_float(v1[jj])) * inv;
              const float a1 = (__uint_as_float(v0[jj + 1]) + __uint_as_float(v1[jj + 1])) * inv;
              const __nv_bfloat162 hh = pl2_from_floats(a0, a1);
              const float2 ff = pl2_to_float2(hh);
              const __nv_bfloat162 ll = pl2_from_floats(a0 - ff.x, a1 - ff.y);
              hp[e] = *reinterpret_cast<const uint32_t*>(&hh);
              lp[e] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            *reinterpret_cast<uint4*>(oh + c * 32 + g * 8) = hv;
            *reinterpret_cast<uint4*>(ol + c * 32 + g * 8) = lv;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);
      asm volatile("bar.sync 1, 256;" ::: "memory");           // rs_part is rewritten by the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace umma
}  // namespace vxb

// patchify: Conv3d(64 -> 64, kernel k, stride s, replicate padding k/2) + LeakyReLU on the channels-last fp32
// grid d0 [B, V, V, V, 64] -> voxel tokens [B, S^3, 64], S = V/s  (perceiver_lang_io.py:223-226,363).
//
// With stride >= kernel - (a few) every input voxel is read about once, so the op is a streaming read of d0
// (256 MB/sample) feeding a [128 tokens] x [64] x [K = k^3 * 64] GEMM per tile.  The operand rows of a token tile
// are not an affine box (stride-s windows, clamped at the borders), so instead of TMA the four loader warps gather
// them.  fp32 input (PLANES = false): thread r owns token row r, loads the 64 fp32 channels of its (clamped) tap voxel
// with 128-bit loads, splits them into 16-bit hi/lo and writes the SWIZZLE_128B K-major rows of the A stage itself.
// Plane input (PLANES = true, the forward's path: d0 already exists as padded hi/lo planes): the rows are copied with
// cp.async (16 B) straight into the swizzled stage, eight consecutive lanes per 128-byte row so that every request
// is a full line, PF_DIST (tile, tap) steps ahead of the step handed to the MMA warp and across tile boundaries.
// Weights [W_hi ; W_lo] (N = 128) arrive by TMA; one elected lane issues  D[:,0:128] += A_hi [W_hi;W_lo]^T  and
// D[:,0:64] += A_lo W_hi^T  per 16-wide k step (split x3, fp32 accumulation in TMEM).
// Warps: 0-3 = loaders (+ epilogue of their TMEM lane quarter), 4 = weight TMA producer, 5 = MMA issuer.
#pragma once
#include "umma_gemm.cuh"

namespace vxb {
namespace umma {

constexpr int PF_THREADS = 192;
constexpr int PF_ASTAGES = 5;                 // 32 KB each (hi + lo planes of 128 rows x 128 B)
constexpr int PF_DIST = 4;                    // cp.async prefetch distance in stages (< PF_ASTAGES)
constexpr int PF_WSTAGES = 3;                 // 16 KB each ([W_hi ; W_lo] x 64 channels)
constexpr int PF_ABYTES = 2 * 128 * 128;
constexpr int PF_WBYTES = 128 * 128;
constexpr int PF_SMEM = PF_ASTAGES * PF_ABYTES + PF_WSTAGES * PF_WBYTES + 1024;

struct PatchifyParams {
  const float* x;        // [B, V, V, V, 64] fp32, or null when the input comes as planes
  const __nv_bfloat16* xhi;   // hi/lo planes of the replicate-padded grid [B, V+2, V+2, V+2, 64]
  const __nv_bfloat16* xlo;
  const float* bias;     // [64]
  float* out;            // [B, S^3, 64]
  int B, V, S, k, s, pad;
  int tokens;            // B * S^3
  int tiles;             // ceil(tokens / 128)
  float act_slope;
};

__device__ __forceinline__ void split_store_row(uint8_t* hi_row, uint8_t* lo_row, int r, const float4 (&v)[16]) {
  // 64 floats -> 8 chunks of 8 bf16 (16 bytes) per plane; chunk c of row r lives at ((c ^ (r & 7)) * 16)
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 a = v[2 * c], b = v[2 * c + 1];
    const __nv_bfloat162 h0 = pl2_from_floats(a.x, a.y), h1 = pl2_from_floats(a.z, a.w);
    const __nv_bfloat162 h2 = pl2_from_floats(b.x, b.y), h3 = pl2_from_floats(b.z, b.w);
    const float2 f0 = pl2_to_float2(h0), f1 = pl2_to_float2(h1), f2 = pl2_to_float2(h2), f3 = pl2_to_float2(h3);
    const __nv_bfloat162 l0 = pl2_from_floats(a.x - f0.x, a.y - f0.y), l1 = pl2_from_floats(a.z - f1.x, a.w - f1.y);
    const __nv_bfloat162 l2 = pl2_from_floats(b.x - f2.x, b.y - f2.y), l3 = pl2_from_floats(b.z - f3.x, b.w - f3.y);
    uint4 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
    hv.z = *reinterpret_cast<const uint32_t*>(&h2); hv.w = *reinterpret_cast<const uint32_t*>(&h3);
    lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
    lv.z = *reinterpret_cast<const uint32_t*>(&l2); lv.w = *reinterpret_cast<const uint32_t*>(&l3);
    const int off = (c ^ (r & 7)) * 16;
    *reinterpret_cast<uint4*>(hi_row + off) = hv;
    *reinterpret_cast<uint4*>(lo_row + off) = lv;
  }
}

// PLANES = true: the input is already split (written by input_preprocess): the loaders only copy 2 x 128 bytes per row
template <bool PLANES>
__global__ void __launch_bounds__(PF_THREADS, 1)
patchify_umma_kernel(const __grid_constant__ CUtensorMap mapW, const PatchifyParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bars[2 * PF_ASTAGES + 2 * PF_WSTAGES + 2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t* a_base = smem;
  uint8_t* w_base = smem + PF_ASTAGES * PF_ABYTES;
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + PF_ASTAGES;
  uint64_t* w_full = bars + 2 * PF_ASTAGES;
  uint64_t* w_empty = w_full + PF_WSTAGES;
  uint64_t* acc_full = w_empty + PF_WSTAGES;
  uint64_t* acc_empty = acc_full + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapW);
    for (int i = 0; i < PF_ASTAGES; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < PF_WSTAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int k3 = p.k * p.k * p.k;

  if (warp < 4) {
    // ===================================================== loaders (token row = threadIdx.x) + epilogue
    const int r = threadIdx.x;
    int st = 0;
    uint32_t ph = 0, accph = 0;
    const int S = p.S, V = p.V;
    // PLANES: the gather runs as cp.async (16 B, straight into the swizzled stage) PF_DIST (tile, tap) steps ahead of
    // the step being handed to the MMA warp, across tile boundaries -- PF_DIST x 32 KB in flight per SM.
    // Eight consecutive lanes fetch the eight 16-byte chunks of one token row (a full 128-byte line per plane), so a
    // warp-wide cp.async touches 4 rows x 128 B; a thread serves rows (warp*32 + i*4 + lane/8), i = 0..7.
    [[maybe_unused]] int is_tile = blockIdx.x, is_tap = 0, is_st = 0;
    [[maybe_unused]] uint32_t is_ph = 0;
    [[maybe_unused]] uint32_t rs_pos[8];       // per served row: od*s | oh*s << 10 | ow*s << 20 | valid << 31
    [[maybe_unused]] int rs_b[8];              // batch index
    [[maybe_unused]] auto issue = [&]() {
      if constexpr (PLANES) {
        if (is_tile < p.tiles) {
          const int chunk = lane & 7, sub = lane >> 3;
          if (is_tap == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int tk = is_tile * 128 + warp * 32 + i * 4 + sub;
              const bool ok = tk < p.tokens;
              const int tkc = ok ? tk : 0;
              const int ow_ = tkc % S, oh_ = (tkc / S) % S, od_ = (tkc / (S * S)) % S;
              rs_b[i] = tkc / (S * S * S);
              rs_pos[i] = (uint32_t)(od_ * p.s) | ((uint32_t)(oh_ * p.s) << 10) | ((uint32_t)(ow_ * p.s) << 20) | (ok ? 0x80000000u : 0u);
            }
          }
          const int tw = is_tap % p.k - p.pad, th = (is_tap / p.k) % p.k - p.pad, td = is_tap / (p.k * p.k) - p.pad;
          const int Vp = V + 2;
          mbar_wait(&a_empty[is_st], is_ph ^ 1);
          const uint32_t sa = smem_u32(a_base + is_st * PF_ABYTES);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = warp * 32 + i * 4 + sub;
            const uint32_t ps = rs_pos[i];
            const int vd = min(max((int)(ps & 1023u) + td, 0), V - 1), vh = min(max((int)((ps >> 10) & 1023u) + th, 0), V - 1),
                      vw = min(max((int)((ps >> 20) & 1023u) + tw, 0), V - 1);
            const size_t o = ((((size_t)rs_b[i] * Vp + vd + 1) * Vp + vh + 1) * Vp + vw + 1) * 64 + chunk * 8;
            const uint32_t nbytes = (ps >> 31) ? 16u : 0u;         // src-size 0 = zero fill (rows past the last token)
            const uint32_t dst = sa + row * 128 + (uint32_t)((chunk ^ (row & 7)) * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(p.xhi + o), "r"(nbytes) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 128 * 128), "l"(p.xlo + o), "r"(nbytes) : "memory");
          }
          if (++is_st == PF_ASTAGES) { is_st = 0; is_ph ^= 1; }
          if (++is_tap == k3) { is_tap = 0; is_tile += gridDim.x; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    };
    if constexpr (PLANES) {
#pragma unroll 1
      for (int i = 0; i < PF_DIST; ++i) issue();
    }
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      const int tok = tile * 128 + r;
      const bool valid = tok < p.tokens;
      int b = 0, od = 0, oh = 0, ow = 0;
      if (valid) {
        ow = tok % S; oh = (tok / S) % S; od = (tok / (S * S)) % S; b = tok / (S * S * S);
      }
      const int d0 = od * p.s - p.pad, h0 = oh * p.s - p.pad, w0 = ow * p.s - p.pad;
      if constexpr (PLANES) {
#pragma unroll 1
        for (int tap = 0; tap < k3; ++tap) {
          asm volatile("cp.async.wait_group %0;" ::"n"(PF_DIST - 1) : "memory");   // this step's rows have landed
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");             // generic-proxy writes -> visible to tcgen05
          mbar_arrive(&a_full[st]);
          if (++st == PF_ASTAGES) { st = 0; ph ^= 1; }
          issue();                                                                 // step + PF_DIST, into the stage the MMA of step - 1 frees
        }
      } else {
      const float* xb = p.x + (size_t)b * V * V * V * 64;
      auto src = [&](int tap) -> const float4* {
        const int tw = tap % p.k, th = (tap / p.k) % p.k, td = tap / (p.k * p.k);
        const int vd = min(max(d0 + td, 0), V - 1), vh = min(max(h0 + th, 0), V - 1), vw = min(max(w0 + tw, 0), V - 1);
        return reinterpret_cast<const float4*>(xb + (((size_t)vd * V + vh) * V + vw) * 64);
      };
      float4 cur[16];
      {
        const float4* sp = src(0);
#pragma unroll
        for (int j = 0; j < 16; ++j) cur[j] = valid ? __ldg(sp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int tap = 0; tap < k3; ++tap) {
        float4 nxt[16];
        if (tap + 1 < k3) {
          const float4* sp = src(tap + 1);
#pragma unroll
          for (int j = 0; j < 16; ++j) nxt[j] = valid ? __ldg(sp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(&a_empty[st], ph ^ 1);
        uint8_t* sa = a_base + st * PF_ABYTES;
        split_store_row(sa + r * 128, sa + 128 * 128 + r * 128, r, cur);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05
        mbar_arrive(&a_full[st]);
        if (++st == PF_ASTAGES) { st = 0; ph ^= 1; }
        if (tap + 1 < k3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
        }
      }
      }
      // ---- epilogue of this tile: out = act(D[:, 0:64] + D[:, 64:128] + bias)
      mbar_wait(acc_full, accph);
      accph ^= 1;
      tc_fence_after();
      float* orow = p.out + (size_t)tok * 64;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v0[32], v1[32];
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v0);
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(64 + c0), v1);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(p.bias + c0 + j);
            float4 o;
            o.x = __uint_as_float(v0[j]) + __uint_as_float(v1[j]) + bv.x;
            o.y = __uint_as_float(v0[j + 1]) + __uint_as_float(v1[j + 1]) + bv.y;
            o.z = __uint_as_float(v0[j + 2]) + __uint_as_float(v1[j + 2]) + bv.z;
            o.w = __uint_as_float(v0[j + 3]) + __uint_as_float(v1[j + 3]) + bv.w;
            if (p.act_slope >= 0.f) {
              o.x = o.x > 0.f ? o.x : o.x * p.act_slope; o.y = o.y > 0.f ? o.y : o.y * p.act_slope;
              o.z = o.z > 0.f ? o.z : o.z * p.act_slope; o.w = o.w > 0.f ? o.w : o.w * p.act_slope;
            }
            *reinterpret_cast<float4*>(orow + c0 + j) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  } else if (warp == 4) {
    // ===================================================== weight producer
    if (lane == 0) {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        for (int tap = 0; tap < k3; ++tap) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_expect_tx(&w_full[ws], PF_WBYTES);
          tma_load_2d(&mapW, &w_full[ws], w_base + ws * PF_WBYTES, 0, tap * 128);
          if (++ws == PF_WSTAGES) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== MMA issuer (warp-uniform loops, one elected lane issues)
    uint32_t leader;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    constexpr uint32_t idesc128 = make_idesc(128), idesc64 = make_idesc(64);
    int st = 0, ws = 0;
    uint32_t ph = 0, wph = 0, accph = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
      mbar_wait(acc_empty, accph ^ 1);
      accph ^= 1;
      tc_fence_after();
      for (int tap = 0; tap < k3; ++tap) {
        mbar_wait(&a_full[st], ph);
        mbar_wait(&w_full[ws], wph);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(a_base + st * PF_ABYTES), sa_lo = sa_hi + 128 * 128;
        const uint32_t sw = smem_u32(w_base + ws * PF_WBYTES);
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da_hi = make_desc(sa_hi + ks * 32, 1024), da_lo = make_desc(sa_lo + ks * 32, 1024);
            const uint64_t db = make_desc(sw + ks * 32, 1024);
            tc_mma_bf16(tmem_u, da_hi, db, idesc128, (tap | ks) != 0);   // hi*hi -> cols [0,64), hi*lo -> [64,128)
            tc_mma_bf16(tmem_u, da_lo, db, idesc64, 1u);                  // lo*hi -> cols [0,64)
          }
          tc_commit(&a_empty[st]);
          tc_commit(&w_empty[ws]);
          if (tap == k3 - 1) tc_commit(acc_full);
        }
        __syncwarp();
        if (++st == PF_ASTAGES) { st = 0; ph ^= 1; }
        if (++ws == PF_WSTAGES) { ws = 0; wph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
  }
}

}  // namespace umma
}  // namespace vxb

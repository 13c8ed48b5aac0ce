// Host-visible types of the tcgen05 split-bf16 engine (kernel in umma_gemm.cuh, host code in umma_ops.cu).
#pragma once
#include "planes16.cuh"
#include "common.cuh"

namespace vxb {
namespace umma {

constexpr int BM = 128;
constexpr int BK = 64;  // bf16 elements per k-block row = 128 bytes = one SWIZZLE_128B span
constexpr int THREADS = 192;

// Per-k-block operand addressing, derived in the producer from a few integers (no table):
//   plain GEMM (taps == 0): A block kb = columns [64 kb, 64 kb + 64) of source 0, rows unshifted
//   convolution (taps = k per axis): kb -> (tap, channel block); the tap is a row shift of
//   ((dz-c)*Vp + (dy-c))*Vp + (dx-c) in the flat padded grid; channel blocks [0, cb_src0) come from
//   source 0, the rest from source 1 (fused channel concat); W block kb = columns [64 kb, 64 kb + 64)
struct KPlan {
  int num_kb;
  int taps;        // 0 = plain GEMM, else kernel size per axis (odd)
  int Vp;          // padded grid extent (conv)
  int cpb;         // channel blocks (of 64) per tap
  int cb_src0;     // channel blocks taken from source 0
};

enum { ROWS_PLAIN = 0, ROWS_CONV_FLAT = 1, ROWS_PHASE = 2 };
// epilogue value modes (attention runs as three GEMMs without ever storing fp32 scores):
//   EPI_STORE   out = act(alpha*acc + bias) + residual, optionally divided by row_div[row]
//   EPI_ROWMAX  row_stat[row] = max(row_stat[row], max_n alpha*acc)            (atomic, nothing stored)
//   EPI_EXP     p = 2^(alpha*acc - row_sub[row]); row_stat[row] += sum_n p (atomic); p stored as planes
//   EPI_GEGLU   out[:, j] = (acc_a + bias_a) * gelu_erf(acc_g + bias_g) as planes; the GEMM's N columns come in
//               interleaved 32-column chunks [a | gate] (weights / bias permuted at prepare time), N_out = N / 2
enum { EPI_STORE = 0, EPI_ROWMAX = 1, EPI_EXP = 2, EPI_GEGLU = 3 };

struct Epilogue {
  int M, N;                   // logical extents (rows beyond M / cols beyond N are not stored)
  int row_mode;
  // ROWS_CONV_FLAT / ROWS_PHASE: rows are flat indices into a padded [B, Vp, Vp, Vp] grid with `pad`
  // halo voxels per side; only interior rows are stored.
  int Vp, pad;
  int out_padded;             // CONV_FLAT: 1 = output rows keep the padded geometry (row = m), 0 = compact [B,V^3]
  int phase_s;                // ROWS_PHASE: column block j (64 cols) = phase p = n/64 -> fine voxel s*q + r
  const uint8_t* phase_perm;  // optional device table: column block j holds phase phase_perm[j] (blocks sorted so that the
                              // four phases of an N tile share their zero taps, see Params::kmask)
  int out_Vp, out_pad;        // ROWS_PHASE: geometry of the fine output grid (padded)
  const float* bias;          // [N] (ROWS_PHASE: [64], shared by all phases)
  float alpha;
  const float* alpha_dev;     // generic epilogue: optional device scalar multiplied into alpha
  float act_slope;            // < 0: none
  const float* residual;      // fp32 [(row % res_rows), ldr] added after the activation, or null
  int res_rows, ldr;
  float* out_f32;             // fp32 output or null
  long long ldc;
  __nv_bfloat16* out_hi;      // bf16 plane outputs or null
  __nv_bfloat16* out_lo;
  long long ldp;
  const float* f8a;           // non-null (generic epilogue, ldp = 64 only): out_lo receives the c8 plane of conv_f8c.cuh instead
                              // of the fp16 lo plane -- per 32-channel block 32 B of e4m3(2^11 a x_lo) then 32 B of e4m3(a x), a = *f8a
  int transpose_planes;       // planes written transposed: element (row, col) -> out[(col) * ldp + row] (V^T for attention)
  int mode;                   // EPI_*
  float* row_stat;            // EPI_ROWMAX / EPI_EXP: per-row statistic, indexed rs offset(z) + row
  const float* row_sub;       // EPI_EXP: per-row value subtracted before exp2
  const float* row_div;       // EPI_STORE: per-row divisor (softmax normalisation folded into P V)
};

struct Params {
  int m_tiles, n_tiles;
  KPlan plan;
  // batching: tile z = zb * Hz + zh; operand bases shift per z (attention heads / batches)
  int batches, Hz;
  int a_row_zb, a_col_zh;        // A: rows per zb, columns per zh
  int a_row_zh;                  // A: rows per zh
  int a_row_off;                 // A: first row of every batch entry (conv modes skip the all-halo leading z planes)
  long long a_col_off, w_col_off; // plain GEMM: constant K offsets of the two operands (weight gradients: a tap = a K shift)
  int terms;                     // 3 = hi*hi + hi*lo + lo*hi (default), 1 = hi*hi only (lo planes are not loaded),
                                 // 2 = fp16 hi*hi + one E4M3 MMA on the "lo" planes (c8 layout, 64-channel blocks)
  long long rs_zb, rs_zh;        // row statistic element offsets per zb / zh
  int w_row_zb, w_row_zh, w_col_zh;
  // tap_m != 0 (convolution weight gradients): every M tile multiplies the SAME A rows at its own K offset against its own W
  // rows -- M tile t reads A columns shifted by tap_acol[t] and W rows tap_wrow[t]..; its output rows are t * BM.. as usual.
  // All taps of one K chunk are adjacent in the tile order, so the chunk is fetched from HBM once and re-read from L2.
  // tap_rot[t]: M tile t walks its K blocks rotated by that many blocks (sum order is free), chosen so that ALL taps read the
  // same A columns at the same time -- the A chunk then comes from HBM once and from L2 26 times.
  // kblk_a / kblk_w != 0: K-blocked operand storage [K / 64][rows][64] (a K block of all rows is one contiguous 16 KB box; the
  // plain [rows][K] layout of a 17M-column operand puts every row of a TMA box into its own 2 MB page: the weight-gradient GEMM
  // ran at 15 % tensor-pipe activity with DRAM at 14 % -- translation-bound).  Values = rows per K block of A / W.
  int kblk_a, kblk_w;
  // kmask != null (num_kb <= 32): device array, one word per N tile; K block kb of that tile is loaded and multiplied only
  // when bit kb is set.  The folded up-convolution's weight matrix is block-sparse (phase 0 of an axis never reads the +1
  // neighbour, phase s-1 never the -1 one: 2197 of 3375 (phase, tap) blocks are non-zero); the masks come from the weights
  // themselves (upconv_kmask_build), so skipping is exact.
  const uint32_t* kmask;
  int tap_m;
  int tap_acol[27];
  int tap_wrow[27];
  int tap_rot[27];
  long long c_zb, c_zh;          // fp32 output element offsets per zb / zh
  long long p_zb, p_zh;          // plane output element offsets per zb / zh
  Epilogue ep;
};


// bf16 hi/lo planes of an fp32 matrix: x ~= hi + lo, row-major with ld elements between rows
struct Planes {
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  long long ld;
};
struct Operand {
  Planes p;
  long long rows, cols;   // logical extent seen by TMA (out-of-bounds reads are zero)
};

inline size_t plane_elems(long long rows, long long ld) { return (size_t)rows * (size_t)ld; }
inline long long pad8(long long v) { return (v + 7) / 8 * 8; }

int split_rows(const float* x, long long ldx, long long rows, int cols, Planes out, cudaStream_t st);
int pad_split(const float* x, int B, int V, int pad, int C, Planes out, cudaStream_t st);
int halo_fill(Planes p, int B, int V, int pad, int C, cudaStream_t st, __nv_bfloat16* third = nullptr);   // lo / third may be null
// D = A W^T on tcgen05; A1 is the optional second A source (fused channel concat in convolutions)
int gemm(const Operand& A0, const Operand* A1, const Operand& W, int n_tile, Params p, cudaStream_t st);
long long launches();   // tcgen05 GEMM launches so far in this process

// fp32-in / fp32-out wrappers (operands split into `scratch`)
size_t linear_scratch_bytes(long long M, long long N, long long K, bool split_w);
int linear_f32(const float* A, int lda, const float* W, int ldw, const Planes* Wpre, const float* bias,
               const float* residual, int res_rows, int ldr, float* C, int ldc, int M, int N, int K, float alpha,
               float act_slope, Arena& scratch, cudaStream_t st);
size_t conv3d_scratch_bytes(int B, int V, int C0, int C1, int k);
int conv3d_f32(const float* x0, const float* x1, int C0, int C1, const Planes& Wp, const float* bias, float* out,
               int B, int V, int Co, int k, float act_slope, Arena& scratch, cudaStream_t st);
size_t upconv_scratch_bytes(int B, int S, int Ci);
// f8c form of the folded up-convolution GEMM: Wp = upconv_f8c_prepare planes, sc = device scalars {alpha of `low`, 1/(alpha beta)}
struct F8cGemm { const float* alpha; const float* unscale; };
// block sparsity of the folded weights (upconv_kmask_build): K-block mask per N tile and, optionally, the phase held by every
// 64-column block (the weight planes must then be stored in that order: upconv_f8c_prepare(..., perm))
constexpr int UPCONV_NT = 256, UPCONV_MAX_TILES = 64;
struct UpconvSparsity { const uint32_t* kmask; const uint8_t* phase_perm; };
int upconv_kmask_build(const float* wfold, int s, uint32_t* nz /*[s^3]*/, uint8_t* perm /*[s^3]*/, uint32_t* kmask_natural /*[64]*/,
                       uint32_t* kmask_perm /*[64]*/, cudaStream_t st);
int upconv_f32(const float* low, const Planes& Wp, const float* bias, float* out, int B, int S, int Ci, int Co, int s,
               float act_slope, Arena& scratch, cudaStream_t st, const Planes* out_planes = nullptr, const float* f8a = nullptr,
               const F8cGemm* f8g = nullptr, const UpconvSparsity* sp = nullptr);
// static operand of that GEMM: hi = fp16(2^-5 beta W), lo = per 64 columns [64 x e4m3(2^-11 beta W) | 64 x e4m3(beta W_lo)];
// beta (largest power of two with 2^-11 beta max|W| <= 240) is left in *beta_out
// (perm: output row block j of 64 rows is taken from source row block perm[j])
int upconv_f8c_prepare(const float* wfold, long long rows, long long cols, Planes out, float* beta_out, unsigned int* tmp,
                       cudaStream_t st, const uint8_t* perm = nullptr);

// ---- backward (training) contractions on the tensor cores ------------------------------------------------------------
// Gradient tensors span many decades (softmax-over-10^6 logit gradients are ~1e-8), below the fp16 planes' range, so every
// operand of these wrappers is scaled by a power of two derived on the device from its max-abs (no host sync) and the
// result is unscaled by a small post pass.
// C[M,N] (+)= opA(A) opW(W)^T with K-major products: A is [M,K] (a_trans = false) or stored [K,M] (a_trans = true),
// W is [N,K] (w_trans = false) or stored [K,N] (w_trans = true).  Returns VXB_E_WORKSPACE_TOO_SMALL if scratch is short.
// A (the gradient operand) is always scaled dynamically; W (weights / forward activations) only with w_dynamic_scale.
int gemm_any_f32(const float* A, long long lda, bool a_trans, const float* W, long long ldw, bool w_trans, float* C, int ldc,
                 int M, int N, int K, bool accumulate, Arena& scratch, cudaStream_t st, bool w_dynamic_scale = false);
size_t gemm_any_scratch_bytes(int M, int N, int K, bool accumulate);
// Batched K = dim_head products of the attention backward: out[b, h] = alpha * A_bh W_bh^T (fp32 [B, H, Nq, ldc]); A scaled
// per tensor when a_dynamic (a gradient).  VXB_E_UNSUPPORTED_SHAPE / VXB_E_WORKSPACE_TOO_SMALL: the caller runs its FFMA form.
size_t attn_scores_scratch_bytes(int B, int H, int Nq, int Nk, int dh);
int attn_scores_f32(const float* A, int lda, long long abs_, const float* W, int ldw, long long wbs, float* out, long long ldc,
                    int B, int H, int Nq, int Nk, int dh, float alpha, bool a_dynamic, Arena& scratch, cudaStream_t st);
// Weight gradient of the 3x3x3 convolution on cat[x0, x1] (64 channels each, fp32 compact [B, V^3, 64]) given the
// pre-activation gradient gz [B, V^3, 64]:  dwt[(tap, ci)][co] = sum_rows xpad[row + shift(tap)][ci] gz[row][co]
// (tap-major, the layout bwd::wgrad_to_torch_kernel converts).  Runs as 27 split-K GEMMs on channel-major (transposed)
// planes of the replicate-padded inputs and the zero-padded gradient: a tap is a constant K offset of the A operand.
int conv3_wgrad_f32(const float* x0, const float* x1, const float* gz, float* dwt, int B, int V, Arena& scratch, cudaStream_t st);
size_t conv3_wgrad_scratch_bytes(int B, int V);
// Padded-gradient grid of a stride-1 convolution: gxp[j][b, p, 0:64] = sum_t w_t^T gz[b, p - t] for the j-th block of 64
// input channels, p in (V + 2 (k/2))^3, gz [B, V^3, Cz] zero outside the grid.  wd: fp32 tap-major dgrad weights
// [Cx][k^3][Cz] (bwd::conv_dgrad_weight_kernel).  gxp blocks are consecutive [B, (V+2pad)^3, 64] tensors, multiplied by
// the device scalar *scale_out (the caller's fold divides it out).
int conv_dgrad_f32(const float* gz, int Cz, const float* wd, int Cx, float* gxp, int B, int V, int k, float* scale_out,
                   Arena& scratch, cudaStream_t st);
size_t conv_dgrad_scratch_bytes(int B, int V, int Cz, int Cx, int k);

// ---- plane-domain building blocks of the transformer (no fp32 round trips between GEMMs)
// y = LayerNorm(x) written as planes [rows, n]; input rows may be a strided slice per batch
// f8alpha != null (n % 64 == 0): f8c A-operand planes -- hi = fp16(32 a y), out.lo = c8 blocks [64 x e4m3(2^11 a y_lo) | 64 x e4m3(a y)]
int layernorm_planes(const float* x, size_t x_batch_stride, int rows_per_batch, const float* w, const float* b,
                     Planes out, long long rows, int n, cudaStream_t st, const float* f8alpha = nullptr);
// device scalar a = largest power of two with a * (sqrt(n - 1) max|w| + max|b|) <= 240: the bound of |LayerNorm(x)|
int layernorm_f8c_alpha(const float* w, const float* b, int n, float* alpha_out, cudaStream_t st);
// out[0] = 1 / (alpha[0] * beta[0])
int f8c_unscale(const float* alpha, const float* beta, float* out, cudaStream_t st);
// out = h[:, :n] * gelu_erf(h[:, n:]) written as planes [rows, n]
int geglu_planes(const float* h, Planes out, long long rows, int n, cudaStream_t st);
struct LinOut {
  const float* bias = nullptr;
  float alpha = 1.f, act_slope = -1.f;
  const float* residual = nullptr;
  int res_rows = 1, ldr = 0;
  float* out_f32 = nullptr;
  long long ldc = 0;
  const Planes* out_planes = nullptr;   // planes [M, ld] or, with transposed, [batches][N][ld] (ld >= rows per batch)
  int transposed = 0;
  int batches = 1;                      // transposed only: M = batches * rows_per_batch
  int geglu = 0;                        // fused GEGLU epilogue (out_planes has N / 2 columns; W and bias pre-permuted)
  // terms == 2 (fp16 hi*hi + one E4M3 MMA, see umma_gemm.cuh): A and W are f8c planes (layernorm_planes with f8alpha /
  // upconv_f8c_prepare), alpha_dev = device scalar 1 / (alpha_A beta_W) multiplied into the accumulator
  int terms = 3;
  const float* alpha_dev = nullptr;
};
// C[M,N] = A[M,K] W[N,K]^T with plane operands
int linear_planes(const Planes& A, long long M, int K, const Planes& W, int N, const LinOut& o, cudaStream_t st);
// V^T planes [B][inner][ld >= Nk] = Wv[inner, K] ctx_b[Nk, K]^T per batch: the projection written directly in the
// layout the attention kernel consumes (the weight matrix is the M operand, so the stores stay row-contiguous)
int project_vt(const Planes& ctx, int B, int Nk, int K, const Planes& Wv, int inner, const Planes& vt, cudaStream_t st,
               int terms = 3, const float* alpha_dev = nullptr);
// softmax(scale q k^T) v per (batch, head) as three tcgen05 GEMMs (row max, exp + row sum -> P planes, P V / sum).
// Q [(B or 1)*Nq, H*dh] (q_batched = 0: one Q shared by all batches), K [B*Nk, H*dh], Vt [B*H*dh, pad8(Nk)],
// P scratch planes [B*H*Nq, pad8(Nk)], O planes [B*Nq, H*dh]; rowmax / rowsum: B*H*Nq floats each.
// drop (training forward): softmax -> dropout -> @ v with the counter-based mask of common.cuh dropout_keep
struct AttnDrop { unsigned long long seed; unsigned int thresh; float inv_keep; int ld; };
int attention_planes(const Planes& Q, int q_batched, const Planes& K, const Planes& Vt, int B, int H, int Nq, int Nk,
                     int dh, float scale, float* rowmax, float* rowsum, const Planes& P, const Planes& O,
                     cudaStream_t st, const AttnDrop* drop = nullptr);

// ---- input-stationary 3x3x3 convolution (conv_umma.cuh), Co = 64, channels a multiple of 32
// weights: tap-major fp32 [64][27][C0+C1] -> bf16 [ncb][27][{hi,lo}][64][32]
size_t conv3_weight_elems(int Cin);
int conv3_prepare_weights(const float* w_tapmajor, int Cin, __nv_bfloat16* wc, cudaStream_t st);
// x0 / x1: hi-lo planes of the replicate-padded grids [B, V+2, V+2, V+2, 64]; out fp32 [B, V^3, 64]
// Fused tail of the final convolution (trans_decoder taps + ss_final / max-pool): u = act(conv) is never stored.
struct ConvTail {
  const float* tail_w = nullptr;   // [27][64] tap-major weights of the 64 -> 1 convolution
  const float* tail_b = nullptr;   // [1]
  float* ptap = nullptr;           // scratch [B][27][V^3] fp32
  float* ss_partial = nullptr;     // scratch, conv3_tail_partial_floats(B, V) floats
  float* q_trans;        // out [B, V^3]
  // optional second 64 -> 1 head on the same u (2 robots: trans_decoder_left_arm)
  const float* tail_w2 = nullptr;
  const float* tail_b2 = nullptr;
  float* ptap2 = nullptr;
  float* q_trans2 = nullptr;
  float* ss;             // out: soft-argmax [B, ss_stride] (3 per channel) and max [B, mx_stride]
  int ss_stride;
  float* mx;
  int mx_stride;
};
size_t conv3_tail_partial_floats(int B, int V);
int conv3_tail_finish(const ConvTail& tail, int B, int V, cudaStream_t st);   // gather + merge after conv3_planes(tail)
int conv3_planes(const Planes& x0, const Planes* x1, int C0, int C1, const __nv_bfloat16* wc, const float* bias,
                 float act_slope, float* out, int B, int V, cudaStream_t st, const ConvTail* tail = nullptr);

// ---- f16 + fp8-corrected variant of the input-stationary convolution (conv_f8c.cuh; scalars and layouts in umma_ops.cu f8c_*)
size_t conv3_f8c_w16_elems(int Cin);
size_t conv3_f8c_w8_bytes(int Cin);
int conv3_f8c_prepare(const float* w_tapmajor, int Cin, int C0, __nv_bfloat16* w16, unsigned int* wmax, cudaStream_t st);
int conv3_f8c_fold_abs(const float* wfold, long long rows, int Ci, float* S, cudaStream_t st);
int conv3_f8c_bound_ipp(const float* grid, long long rows, int CIN, const float* w, const float* bias, int C, unsigned int* gmax,
                        float* f8s, cudaStream_t st);
int conv3_f8c_bound_up(const float* low, long long rows, int Ci, const float* S, long long srows, const float* bias,
                       const unsigned int* wmax, unsigned int* lmax, float* f8s, cudaStream_t st, const float* up_beta = nullptr);
int conv3_f8c_quantize_weights(const float* w_tapmajor, int Cin, int C0, const float* f8s, uint8_t* w8, cudaStream_t st);
int conv3_f8c_planes(const Planes& x0, const Planes* x1, int C0, int C1, const __nv_bfloat16* w16, const uint8_t* w8, const float* f8s,
                     const float* bias, float act_slope, float* out, int B, int V, cudaStream_t st, const ConvTail* tail = nullptr);
size_t conv3_f8c_scratch_bytes(int B, int V);
int conv3_f8c_f32(const float* x, const float* w_tapmajor, const float* bias, float act_slope, float* out, int B, int V,
                  Arena& scratch, cudaStream_t st);
int absmax_cols(const float* x, long long rows, int C, unsigned int* out, cudaStream_t st);

// ---- patchify (patchify_umma.cuh): Conv3d(64 -> 64, k, stride s, replicate pad k/2) + act on fp32 channels-last x
// weights: tap-major fp32 [64][k^3][64] -> bf16 [k^3][{hi,lo}][64][64]
size_t patchify_weight_elems(int k);
int patchify_prepare_weights(const float* w_tapmajor, int k, __nv_bfloat16* wc, cudaStream_t st);
// x fp32 [B,V^3,64], or (xplanes != null) the hi/lo planes of the replicate-padded grid [B,(V+2)^3,64]
int patchify_f32(const float* x, const __nv_bfloat16* wc, const float* bias, float act_slope, float* out, int B, int V,
                 int k, int s, cudaStream_t st, const Planes* xplanes = nullptr);

// unit-test entry: fp32 q/k/v in, fp32 out, through the plane-domain attention (scratch sized by the _bytes query)
size_t attention_f32_scratch_bytes(int B, int H, int Nq, int Nk, int dh);
int attention_f32(const float* q, int ldq, long long qbs, const float* k, const float* v, int ldkv, long long kvbs,
                  float* out, int ldo, long long obs, int B, int H, int Nq, int Nk, int dh, float scale, Arena& scratch,
                  cudaStream_t st, const AttnDrop* drop = nullptr);

inline void params_init(Params& p) {
  memset(&p, 0, sizeof(p));
  p.batches = 1;
  p.Hz = 1;
  p.ep.alpha = 1.f;
  p.ep.act_slope = -1.f;
  p.ep.res_rows = 1;
  p.terms = 3;
}

}  // namespace umma
}  // namespace vxb

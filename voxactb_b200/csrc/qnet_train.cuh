// Training step of the Q-network behind the C ABI: a forward that keeps what the backward
// needs (vxb_qnet_forward_train_f32) and the backward itself (vxb_qnet_backward_f32), i.e. the device side of
// `total_loss.backward()` in QAttentionPerActBCAgent.update (reference qattention_peract_bc_agent.py:484-582) for
// PerceiverVoxelLangEncoder.forward (perceiver_lang_io.py:345-485).  Included by qnet.cu (shares Dims / Prepared / Work).
//
// Everything is fp32 and channels-last; activations are saved, attention probabilities are recomputed.  The voxel grid is
// detached (agent:107-108), so there is no gradient past input_preprocess.  Train-mode dropout acts on the attention
// probabilities (perceiver_lang_io.py:127-128) with a counter-based mask that the backward regenerates.
#pragma once
#include "bwd_ops.cuh"

namespace vxb {

// one transformer block (cross block = index 0, latent layers = 1..depth): activations kept for the backward
struct BlockSaved {
  float *x_in;    // [B,L,D] block input (index 0: unused, the input is the broadcast latents)
  float *xn_a;    // [B,L,D] LN(x_in) (index 0: unused, LN(latents) lives in the prepared arena)
  float *q;       // [B,L,inner] (index 0: unused, prepared q_cross)
  float *kv;      // [B,L,2 inner] (index 0: unused, Work::kv_c)
  float *att;     // [B,L,inner] attention output before to_out
  float *x_mid;   // [B,L,D] after the attention residual
  float *xn_f;    // [B,L,D] LN(x_mid)
  float *ffh;     // [B,L,8D] net.0 output
  float *ffg;     // [B,L,4D] GEGLU output
};

struct TrainBufs {
  std::vector<BlockSaved> blk;     // depth + 1
  float *x_out;                    // [B,L,D] final latents
  float *qn, *qd, *xn_d, *kv_d, *att_d;       // decoder cross attention
  float *stats0, *stats1, *statsF;            // softmax (m, s) per (b, c)
  int *arg0, *arg1, *argF;                    // arg-max position per (b, c)
  float *simA, *simB, *simC;                  // attention probability / gradient buffers
  // dgrad weights (re-laid from the current parameters at the start of every backward)
  float *wd_final, *wd_up0, *wd_fold, *wd_patch;
  // gradients of activations
  float *g_feats, *g_h0, *g_h1, *g_h2, *g_rgc;
  float *g_u, *g_d0, *g_u0, *g_ph, *gxp_big;  // full-resolution tensors (gxp_big: padded-gradient grid of the final conv)
  float *g_low, *g_lowp, *gxp_low, *g_dec;
  float *g_x, *g_xn, *g_ffg, *g_ffh, *g_att, *g_q, *g_kv;
  float *g_ins, *g_ctx, *g_kv_c, *g_qcb, *g_qc, *g_latn, *g_qd, *g_qn, *g_att_d, *g_kv_d;
  float *g_lang, *g_patch, *g_pfeat;
  float *dwt;                                 // wgrad GEMM output ([taps][Ci][Co]) before the layout change
  float *im2col;                              // [B*S^3][27*64] low-resolution im2col (folded up-conv weight gradient)
  float *tmp_small;
  char *tc_scratch;                           // operand planes of the tensor-core backward contractions
  size_t tc_scratch_bytes;
};

static size_t train_tc_scratch_bytes(const Dims& m, int B) {
  const int R = B * m.L, Bn = B * m.n, BT = B * m.T, cq = m.ch * m.cdh, lq = m.lh * m.ldh;
  size_t sb = 0;
  auto g = [&](int M, int N, int K, bool acc) { sb = std::max(sb, umma::gemm_any_scratch_bytes(M, N, K, acc)); };
  g(8 * m.D, m.D, R, false); g(m.D, 4 * m.D, R, false); g(R, 4 * m.D, m.D, false); g(R, m.D, 8 * m.D, false);   // FF
  g(m.D, lq, R, false); g(R, lq, m.D, false); g(2 * lq, m.D, R, false); g(R, m.D, 2 * lq, true); g(lq, m.D, R, false);
  g(2 * cq, m.C, Bn, false); g(Bn, m.C, 2 * cq, false); g(m.C, cq, BT, false); g(BT, cq, m.C, false); g(cq, m.C, BT, false);
  g(BT, m.C, cq, false); g(2 * cq, m.D, R, false); g(R, m.D, 2 * cq, false); g(m.D, cq, R, false); g(R, cq, m.D, false);
  g(m.C, 512, B * m.nl, false);
  g(27 * 64, m.s * m.s * m.s * 64, BT, false); g(BT, 27 * 64, m.s * m.s * m.s * 64, false);                       // folded wgrad / dgrad
  sb = std::max(sb, umma::conv3_wgrad_scratch_bytes(B, m.V));
  sb = std::max(sb, umma::conv_dgrad_scratch_bytes(B, m.V, 64, 128, 3));
  sb = std::max(sb, umma::conv_dgrad_scratch_bytes(B, m.S, m.s * m.s * m.s * 64, 64, 3));
  sb = std::max(sb, umma::conv_dgrad_scratch_bytes(B, m.S, 64, m.C, m.k));
  // forward attention through the fused tensor-core kernel (operand planes of q / k / v^T, O planes)
  sb = std::max(sb, umma::attention_f32_scratch_bytes(B, m.lh, m.L, m.L, m.ldh));
  sb = std::max(sb, umma::attention_f32_scratch_bytes(B, m.ch, m.L, m.n, m.cdh));
  sb = std::max(sb, umma::attention_f32_scratch_bytes(B, m.ch, m.T, m.L, m.cdh));
  // backward: operand planes of the two K = dim_head products of every attention (P logits, gA)
  sb = std::max(sb, umma::attn_scores_scratch_bytes(B, m.lh, m.L, m.L, m.ldh));
  sb = std::max(sb, umma::attn_scores_scratch_bytes(B, m.ch, m.L, m.n, m.cdh));
  sb = std::max(sb, umma::attn_scores_scratch_bytes(B, m.ch, m.T, m.L, m.cdh));
  return sb + 8192;
}

static size_t sim_train_floats(const Dims& m, int B) {
  auto pad4 = [](size_t v) { return (v + 3) / 4 * 4; };
  return std::max((size_t)B * m.ch * m.L * pad4(m.n),
                  std::max((size_t)B * m.lh * m.L * pad4(m.L), (size_t)B * m.ch * m.T * pad4(m.L)));
}

static void carve_train(const Dims& m, int B, Arena& a, TrainBufs& t) {
  const size_t Bz = B, cq = (size_t)m.ch * m.cdh, lq = (size_t)m.lh * m.ldh, inner = std::max(cq, lq);
  const size_t rowsL = Bz * m.L, k3 = (size_t)m.k * m.k * m.k, s3 = (size_t)m.s * m.s * m.s;
  t.blk.resize(m.depth + 1);
  for (int i = 0; i <= m.depth; ++i) {
    BlockSaved& s = t.blk[i];
    s.x_in = a.get<float>(rowsL * m.D);
    s.xn_a = a.get<float>(rowsL * m.D);
    s.q = a.get<float>(rowsL * inner);
    s.kv = a.get<float>(rowsL * 2 * inner);
    s.att = a.get<float>(rowsL * inner);
    s.x_mid = a.get<float>(rowsL * m.D);
    s.xn_f = a.get<float>(rowsL * m.D);
    s.ffh = a.get<float>(rowsL * 8 * m.D);
    s.ffg = a.get<float>(rowsL * 4 * m.D);
  }
  t.x_out = a.get<float>(rowsL * m.D);
  t.qn = a.get<float>(Bz * m.T * m.C);
  t.qd = a.get<float>(Bz * m.T * cq);
  t.xn_d = a.get<float>(rowsL * m.D);
  t.kv_d = a.get<float>(rowsL * 2 * cq);
  t.att_d = a.get<float>(Bz * m.T * cq);
  t.stats0 = a.get<float>(Bz * 2 * 64);
  t.stats1 = a.get<float>(Bz * 2 * m.C);
  t.statsF = a.get<float>(Bz * 2 * 64);
  t.arg0 = a.get<int>(Bz * 64);
  t.arg1 = a.get<int>(Bz * m.C);
  t.argF = a.get<int>(Bz * 64);
  const size_t sf = sim_train_floats(m, B);
  t.simA = a.get<float>(sf);
  t.simB = a.get<float>(sf);
  t.simC = a.get<float>(sf);
  t.wd_final = a.get<float>((size_t)128 * 27 * 64);
  t.wd_up0 = a.get<float>((size_t)m.C * k3 * 64);
  t.wd_fold = a.get<float>((size_t)64 * 27 * s3 * 64);
  t.wd_patch = a.get<float>(k3 * 64 * 64);
  t.g_feats = a.get<float>(Bz * m.flat);
  t.g_h0 = a.get<float>(Bz * 256);
  t.g_h1 = a.get<float>(Bz * 64);
  t.g_h2 = a.get<float>(Bz * 64);
  t.g_rgc = a.get<float>(Bz * (3 * m.R + m.G + m.Cc));
  const size_t full = Bz * m.V3 * 64;
  t.g_u = a.get<float>(full);
  t.g_d0 = a.get<float>(full);
  t.g_u0 = a.get<float>(full);
  t.g_ph = a.get<float>(full);
  const size_t Vp = m.V + 2;
  t.gxp_big = a.get<float>(Bz * Vp * Vp * Vp * 128);
  t.g_low = a.get<float>(Bz * m.T * 64);
  const size_t Sp = m.S + 2, Sq = m.S + 2 * (m.k / 2);
  t.g_lowp = a.get<float>(Bz * Sp * Sp * Sp * 64);
  t.gxp_low = a.get<float>(Bz * Sq * Sq * Sq * m.C);
  t.g_dec = a.get<float>(Bz * m.T * m.C);
  t.g_x = a.get<float>(rowsL * m.D);
  t.g_xn = a.get<float>(std::max(rowsL * m.D, Bz * m.n * m.C));
  t.g_ffg = a.get<float>(rowsL * 4 * m.D);
  t.g_ffh = a.get<float>(rowsL * 8 * m.D);
  t.g_att = a.get<float>(rowsL * inner);
  t.g_q = a.get<float>(rowsL * inner);
  t.g_kv = a.get<float>(rowsL * 2 * inner);
  t.g_ins = a.get<float>(Bz * m.n * m.C);
  t.g_ctx = a.get<float>(Bz * m.n * m.C);
  t.g_kv_c = a.get<float>(Bz * m.n * 2 * cq);
  t.g_qcb = a.get<float>(rowsL * cq);
  t.g_qc = a.get<float>((size_t)m.L * cq);
  t.g_latn = a.get<float>((size_t)m.L * m.D);
  t.g_qd = a.get<float>(Bz * m.T * cq);
  t.g_qn = a.get<float>(Bz * m.T * m.C);
  t.g_att_d = a.get<float>(Bz * m.T * cq);
  t.g_kv_d = a.get<float>(rowsL * 2 * cq);
  t.g_lang = a.get<float>(Bz * m.nl * m.C);
  t.g_patch = a.get<float>(Bz * m.T * 64);
  t.g_pfeat = a.get<float>(Bz * 64);
  t.dwt = a.get<float>(std::max(std::max(k3 * m.C * 64, (size_t)27 * 128 * 64), (size_t)27 * 64 * s3 * 64));
  t.tmp_small = a.get<float>(4096);
  t.im2col = a.get<float>(Bz * m.T * 27 * 64);
  t.tc_scratch_bytes = train_tc_scratch_bytes(m, B);
  t.tc_scratch = a.get<char>(t.tc_scratch_bytes);
}

static int train_supported(const vxb_qnet_desc* d, const Dims& m) {
  if (d->two_robots) {
    set_error("qnet training: the 2-robot encoder is inference-only in this library");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (d->final_input != VXB_FINAL_CAT) {
    set_error("qnet training: the no_skip_connection / no_perceiver ablations are inference-only in this library");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (d->iterations != 1) {
    set_error("qnet training: iterations must be 1 (what launch_utils.create_agent builds)");
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  if (m.S * m.s != m.V || m.k != m.s) {
    // the patchify adjoint below relies on non-overlapping windows (k == s); config 1's k=5,s=4 geometry is inference-only
    set_error("qnet training: needs voxel_patch_size == voxel_patch_stride (got k=%d, s=%d)", m.k, m.s);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  return VXB_OK;
}

struct TrainDropout { float input, attn, decoder; unsigned long long seed; };
// FF net.0 in training: the prepared planes of these weights hold the GEGLU-interleaved row order of the fused inference
// epilogue (vxb_qnet_prepare), so the training forward splits the fp32 weight on the fly instead of looking them up
static int lin_nolookup(Ctx& cx, const float* A, int lda, const float* W, int ldw, const float* bias, float* C, int ldc, int M,
                        int N, int K, int mode) {
  return linear(A, lda, W, ldw, bias, nullptr, 1, 0, C, ldc, M, N, K, 1.f, -1.f, mode, cx.st,
                cx.scratch.base ? &cx.scratch : nullptr, nullptr);
}

static bwd::AttnDropout layer_dropout(const TrainDropout& d, int block /*0 = cross, 1.. = layers, -1 = decoder*/) {
  bwd::AttnDropout a;
  a.p = block < 0 ? d.decoder : (block == 0 ? d.input : d.attn);
  a.seed = d.seed * 0x2545F4914F6CDD1Dull + (unsigned long long)(block + 2) * 0x9E3779B97F4A7C15ull;
  return a;
}

// attention forward with materialised probabilities + train-mode dropout on them (simA is the probability buffer)
static int attention_train(const float* q, int ldq, long long qbs, const float* k, const float* v, int ldkv, long long kvbs,
                           float* out, int ldo, long long obs, int B, int H, int Nq, int Nk, int dh, float scale, float* sim,
                           const bwd::AttnDropout& dr, cudaStream_t st, Arena* tc = nullptr) {
  const int Nkp = (Nk + 3) / 4 * 4;
  if (tc && dh == 64) {
    // tensor-core forward: the fused attention kernel of the inference path with the dropout mask applied to P in its softmax
    // warps (the probabilities are recomputed by the backward anyway): 42 ms of FFMA GEMMs + softmax + dropout passes -> ~6 ms
    Arena local(tc->base, tc->cap);
    umma::AttnDrop ad{dr.seed, dr.p > 0.f ? bwd::dropout_threshold(dr.p) : 0u, dr.p > 0.f ? 1.f / (1.f - dr.p) : 1.f, Nkp};
    const int rc = umma::attention_f32(q, ldq, qbs, k, v, ldkv, kvbs, out, ldo, obs, B, H, Nq, Nk, dh, scale, local, st, &ad);
    if (rc != VXB_E_WORKSPACE_TOO_SMALL) return rc;
  }
  GemmParams p;
  gemm_params_init(p);
  p.M = Nq; p.N = Nk; p.K = dh;
  p.A = q; p.lda = ldq; p.a_stride_zb = qbs; p.a_stride_zh = dh;
  p.W = k; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = sim; p.ldc = Nkp; p.c_stride_zb = (long long)H * Nq * Nkp; p.c_stride_zh = (long long)Nq * Nkp;
  p.Hz = H; p.alpha = scale;
  VXB_TRY((launch_simt_gemm<A_PLAIN, B_NT, O_PLAIN>(p, B * H, st)));
  softmax_rows_kernel<<<(unsigned)((size_t)B * H * Nq), 256, 0, st>>>(sim, Nk, Nkp);
  VXB_LAUNCH_CHECK();
  if (dr.p > 0.f) VXB_TRY(bwd::dropout_rows(sim, sim, (long long)B * H * Nq, Nk, Nkp, dr, st));
  gemm_params_init(p);
  p.M = Nq; p.N = dh; p.K = Nk;
  p.A = sim; p.lda = Nkp; p.a_stride_zb = (long long)H * Nq * Nkp; p.a_stride_zh = (long long)Nq * Nkp;
  p.W = v; p.ldw = ldkv; p.w_stride_zb = kvbs; p.w_stride_zh = dh;
  p.C = out; p.ldc = ldo; p.c_stride_zb = obs; p.c_stride_zh = dh;
  p.Hz = H;
  return launch_simt_gemm<A_PLAIN, B_NN, O_PLAIN>(p, B * H, st);
}

// ------------------------------------------------------------------------------------------------ forward (training)
static int qnet_forward_train_impl(const vxb_qnet_desc* d, const Dims& m, const void* const* params, const Prepared& pw, Work& w,
                                   TrainBufs& t, const float* grid, const float* proprio, const float* lang_tokens, int B,
                                   float* q_trans, float* rot_grip, float* collision, float* arm_out,
                                   const TrainDropout& drop, cudaStream_t st) {
  Ctx cx(d->math_mode, st, d->math_mode == VXB_MATH_F16X3 ? w.scratch : nullptr, w.scratch_bytes);
  cx.wp = pw.planes;
  auto P = [&](int slot) { return (const float*)params[slot]; };
  auto PL = [&](int layer, int slot) {
    return (const float*)params[VXB_P_FIXED_COUNT + layer * VXB_P_LAYER_STRIDE + slot];
  };
  const int mm = d->math_mode;
  const float slope = d->act_slope;
  const int cq = m.ch * m.cdh, lq = m.lh * m.ldh;
  const int rowsL = B * m.L;
  // (1) d0, ss0, max0                                                                            perceiver_lang_io.py:357-360
  VXB_TRY(input_preprocess_ss_run<10>(grid, P(VXB_P_INPRE_W), P(VXB_P_INPRE_B), slope, w.d0, B, m.V, m.V, m.V, 64, w.feats,
                                      m.flat, w.feats + 192, m.flat, w.ss_part, st, nullptr, nullptr, t.stats0));
  VXB_TRY(bwd::channel_argmax(w.d0, w.feats + 192, m.flat, B, (long long)m.V3, 64, t.arg0, st));
  // (2) patchify                                                                                  :363
  VXB_TRY(conv3d(w.d0, nullptr, 64, 0, pw.patch_wt, P(VXB_P_PATCH_B), w.patch, B, m.V, m.S, 64, m.k, m.s, slope, mm, st,
                 nullptr, nullptr));
  // (3) proprio, language, token assembly                                                         :370-422
  VXB_TRY(lin(cx, proprio, m.low, P(VXB_P_PROPRIO_W), m.low, P(VXB_P_PROPRIO_B), nullptr, 1, 0, w.pfeat, 64, B, 64, m.low, 1.f,
              slope, VXB_MATH_FP32_SIMT));
  if (d->no_language) {
    VXB_CUDA(cudaMemsetAsync(w.lang_lin, 0, (size_t)B * m.nl * m.C * sizeof(float), st));
    VXB_TRY(lin(cx, w.lang_lin, m.C, P(VXB_P_LANG_W), d->lang_emb_dim, P(VXB_P_LANG_B), nullptr, 1, 0, w.lang_lin, m.C,
                B * m.nl, m.C, 0, 1.f, -1.f, VXB_MATH_FP32_SIMT));
  } else {
    VXB_TRY(lin(cx, lang_tokens, d->lang_emb_dim, P(VXB_P_LANG_W), d->lang_emb_dim, P(VXB_P_LANG_B), nullptr, 1, 0,
                w.lang_lin, m.C, B * m.nl, m.C, d->lang_emb_dim, 1.f, -1.f, mm));
  }
  assemble_tokens_kernel<<<148 * 8, 256, 0, st>>>(w.lang_lin, w.patch, w.pfeat, nullptr, P(VXB_P_POS_ENCODING), w.ins, B, m.nl,
                                                  m.T, m.C, 64);
  VXB_LAUNCH_CHECK();
  Arena tc_arena(t.tc_scratch, t.tc_scratch_bytes);
  Arena* tcp = (mm == VXB_MATH_F16X3 && t.tc_scratch) ? &tc_arena : nullptr;      // tensor-core attention forward
  // (4) encoder cross attention block                                                              :431-432
  {
    BlockSaved& s = t.blk[0];
    VXB_TRY(layernorm(w.ins, P(VXB_P_CROSS_NORMCTX_W), P(VXB_P_CROSS_NORMCTX_B), w.ctx_n, (size_t)B * m.n, m.C, st));
    VXB_TRY(lin(cx, w.ctx_n, m.C, P(VXB_P_CROSS_KV_W), m.C, nullptr, nullptr, 1, 0, w.kv_c, 2 * cq, B * m.n, 2 * cq, m.C, 1.f,
                -1.f, mm));
    VXB_TRY(attention_train(pw.q_cross, cq, 0, w.kv_c, w.kv_c + cq, 2 * cq, (long long)m.n * 2 * cq, s.att, cq,
                            (long long)m.L * cq, B, m.ch, m.L, m.n, m.cdh, 1.f / sqrtf((float)m.cdh), t.simA,
                            layer_dropout(drop, 0), st, tcp));
    VXB_TRY(lin(cx, s.att, cq, P(VXB_P_CROSS_OUT_W), cq, P(VXB_P_CROSS_OUT_B), P(VXB_P_LATENTS), m.L, m.D, s.x_mid, m.D,
                rowsL, m.D, cq, 1.f, -1.f, mm));
    VXB_TRY(layernorm(s.x_mid, P(VXB_P_CROSS_FF_NORM_W), P(VXB_P_CROSS_FF_NORM_B), s.xn_f, (size_t)rowsL, m.D, st));
    VXB_TRY(lin_nolookup(cx, s.xn_f, m.D, P(VXB_P_CROSS_FF0_W), m.D, P(VXB_P_CROSS_FF0_B), s.ffh, 8 * m.D, rowsL, 8 * m.D, m.D,
                         mm));
    geglu_kernel<<<148 * 8, 256, 0, st>>>(s.ffh, s.ffg, (size_t)rowsL, 4 * m.D);
    VXB_LAUNCH_CHECK();
    float* xo = m.depth > 0 ? t.blk[1].x_in : t.x_out;
    VXB_TRY(lin(cx, s.ffg, 4 * m.D, P(VXB_P_CROSS_FF2_W), 4 * m.D, P(VXB_P_CROSS_FF2_B), s.x_mid, rowsL, m.D, xo, m.D, rowsL,
                m.D, 4 * m.D, 1.f, -1.f, mm));
  }
  // (5) latent self-attention stack                                                                :435-437
  for (int l = 0; l < m.depth; ++l) {
    BlockSaved& s = t.blk[l + 1];
    VXB_TRY(layernorm(s.x_in, PL(l, VXB_PL_ATTN_NORM_W), PL(l, VXB_PL_ATTN_NORM_B), s.xn_a, (size_t)rowsL, m.D, st));
    VXB_TRY(lin(cx, s.xn_a, m.D, PL(l, VXB_PL_Q_W), m.D, nullptr, nullptr, 1, 0, s.q, lq, rowsL, lq, m.D, 1.f, -1.f, mm));
    VXB_TRY(lin(cx, s.xn_a, m.D, PL(l, VXB_PL_KV_W), m.D, nullptr, nullptr, 1, 0, s.kv, 2 * lq, rowsL, 2 * lq, m.D, 1.f, -1.f,
                mm));
    VXB_TRY(attention_train(s.q, lq, (long long)m.L * lq, s.kv, s.kv + lq, 2 * lq, (long long)m.L * 2 * lq, s.att, lq,
                            (long long)m.L * lq, B, m.lh, m.L, m.L, m.ldh, 1.f / sqrtf((float)m.ldh), t.simA,
                            layer_dropout(drop, l + 1), st, tcp));
    VXB_TRY(lin(cx, s.att, lq, PL(l, VXB_PL_OUT_W), lq, PL(l, VXB_PL_OUT_B), s.x_in, rowsL, m.D, s.x_mid, m.D, rowsL, m.D, lq,
                1.f, -1.f, mm));
    VXB_TRY(layernorm(s.x_mid, PL(l, VXB_PL_FF_NORM_W), PL(l, VXB_PL_FF_NORM_B), s.xn_f, (size_t)rowsL, m.D, st));
    VXB_TRY(lin_nolookup(cx, s.xn_f, m.D, PL(l, VXB_PL_FF0_W), m.D, PL(l, VXB_PL_FF0_B), s.ffh, 8 * m.D, rowsL, 8 * m.D, m.D, mm));
    geglu_kernel<<<148 * 8, 256, 0, st>>>(s.ffh, s.ffg, (size_t)rowsL, 4 * m.D);
    VXB_LAUNCH_CHECK();
    float* xo = l + 1 < m.depth ? t.blk[l + 2].x_in : t.x_out;
    VXB_TRY(lin(cx, s.ffg, 4 * m.D, PL(l, VXB_PL_FF2_W), 4 * m.D, PL(l, VXB_PL_FF2_B), s.x_mid, rowsL, m.D, xo, m.D, rowsL, m.D,
                4 * m.D, 1.f, -1.f, mm));
  }
  // (6) decoder cross attention (queries = voxel rows of LN(ins), no residual)                     :440-448
  VXB_TRY(layernorm_batched(w.ins + (size_t)m.nl * m.C, (size_t)m.n * m.C, P(VXB_P_DEC_NORM_W), P(VXB_P_DEC_NORM_B), t.qn, B,
                            m.T, m.C, st));
  VXB_TRY(lin(cx, t.qn, m.C, P(VXB_P_DEC_Q_W), m.C, nullptr, nullptr, 1, 0, t.qd, cq, B * m.T, cq, m.C, 1.f, -1.f, mm));
  VXB_TRY(layernorm(t.x_out, P(VXB_P_DEC_NORMCTX_W), P(VXB_P_DEC_NORMCTX_B), t.xn_d, (size_t)rowsL, m.D, st));
  VXB_TRY(lin(cx, t.xn_d, m.D, P(VXB_P_DEC_KV_W), m.D, nullptr, nullptr, 1, 0, t.kv_d, 2 * cq, rowsL, 2 * cq, m.D, 1.f, -1.f,
              mm));
  VXB_TRY(attention_train(t.qd, cq, (long long)m.T * cq, t.kv_d, t.kv_d + cq, 2 * cq, (long long)m.L * 2 * cq, t.att_d, cq,
                          (long long)m.T * cq, B, m.ch, m.T, m.L, m.cdh, 1.f / sqrtf((float)m.cdh), t.simA,
                          layer_dropout(drop, -1), st, tcp));
  VXB_TRY(lin(cx, t.att_d, cq, P(VXB_P_DEC_OUT_W), cq, P(VXB_P_DEC_OUT_B), nullptr, 1, 0, w.dec, m.C, B * m.T, m.C, cq, 1.f,
              -1.f, mm));
  // (7) ss1 / max1                                                                                  :451
  VXB_TRY(spatial_softmax_run(w.dec, B, m.S, m.S, m.S, m.C, w.feats + 256, m.flat, w.feats + 256 + 3 * m.C, m.flat, w.ss_part,
                              st, t.stats1));
  VXB_TRY(bwd::channel_argmax(w.dec, w.feats + 256 + 3 * m.C, m.flat, B, m.T, m.C, t.arg1, st));
  // (8) up0: conv k (C -> 64) at S^3, folded upsample-conv                                          :454
  VXB_TRY(conv3d(w.dec, nullptr, m.C, 0, pw.up0_wt, P(VXB_P_UP0_B), w.low, B, m.S, m.S, 64, m.k, 1, slope, mm, st,
                 cx.scratch.base ? &cx.scratch : nullptr, cx.find(pw.up0_wt)));
  const umma::UpconvSparsity up_sp{pw.up1_kmask[0], nullptr};       // natural phase order: the 3-term planes of up1_fold
  VXB_TRY(upconv3d_folded(w.low, pw.up1_fold, P(VXB_P_UP1_B), w.u0, B, m.S, 64, 64, m.s, slope, mm, st,
                          cx.scratch.base ? &cx.scratch : nullptr, cx.find(pw.up1_fold), nullptr, nullptr, nullptr,
                          up_sparse(m) ? &up_sp : nullptr));
  // (9) final conv on cat[d0, u0]                                                                   :462
  if (mm == VXB_MATH_F16X3 && cx.scratch.base) {
    // input-stationary tcgen05 convolution (conv_umma.cuh) with the fp32 store epilogue: the GEMM-engine form re-fetches its
    // operand tile per tap from L2 (23.7 ms at B=16 against 18 ms)
    Arena local(cx.scratch.base, cx.scratch.cap);
    const long long prow = (long long)B * (m.V + 2) * (m.V + 2) * (m.V + 2);
    umma::Planes d0p, u0p;
    d0p.ld = u0p.ld = 64;
    d0p.hi = local.get<__nv_bfloat16>((size_t)prow * 64); d0p.lo = local.get<__nv_bfloat16>((size_t)prow * 64);
    u0p.hi = local.get<__nv_bfloat16>((size_t)prow * 64); u0p.lo = local.get<__nv_bfloat16>((size_t)prow * 64);
    if (!local.ok) {
      set_error("qnet_forward_train: scratch too small for the final convolution planes");
      return VXB_E_WORKSPACE_TOO_SMALL;
    }
    VXB_TRY(umma::pad_split(w.d0, B, m.V, 1, 64, d0p, st));
    VXB_TRY(umma::pad_split(w.u0, B, m.V, 1, 64, u0p, st));
    VXB_TRY(umma::conv3_planes(d0p, &u0p, 64, 64, pw.final_wc, P(VXB_P_FINAL_B), slope, w.u, B, m.V, st, nullptr));
  } else {
    VXB_TRY(conv3d(w.d0, w.u0, 64, 64, pw.final_wt, P(VXB_P_FINAL_B), w.u, B, m.V, m.V, 64, 3, 1, slope, mm, st,
                   cx.scratch.base ? &cx.scratch : nullptr, cx.find(pw.final_wt)));
  }
  // (10) trans decoder, ss_final / max, heads                                                       :465-483
  VXB_TRY(trans_stencil_run<64>(w.u, pw.trans_wt, P(VXB_P_TRANS_B), q_trans, B, m.V, st));
  const int off = 256 + 4 * m.C;
  VXB_TRY(spatial_softmax_run(w.u, B, m.V, m.V, m.V, 64, w.feats + off, m.flat, w.feats + off + 192, m.flat, w.ss_part, st,
                              t.statsF));
  VXB_TRY(bwd::channel_argmax(w.u, w.feats + off + 192, m.flat, B, (long long)m.V3, 64, t.argF, st));
  VXB_TRY(lin(cx, w.feats, m.flat, P(VXB_P_DENSE0_W), m.flat, P(VXB_P_DENSE0_B), nullptr, 1, 0, w.h0, 256, B, 256, m.flat, 1.f,
              slope, VXB_MATH_FP32_SIMT));
  VXB_TRY(lin(cx, w.h0, 256, P(VXB_P_DENSE1_W), 256, P(VXB_P_DENSE1_B), nullptr, 1, 0, w.h1, 64, B, 64, 256, 1.f, slope,
              VXB_MATH_FP32_SIMT));
  const int nout = 3 * m.R + m.G + m.Cc;
  VXB_TRY(lin(cx, w.h1, 64, P(VXB_P_RGC_W), 64, P(VXB_P_RGC_B), nullptr, 1, 0, w.rgc, nout, B, nout, 64, 1.f, -1.f,
              VXB_MATH_FP32_SIMT));
  VXB_CUDA(cudaMemcpy2DAsync(rot_grip, (size_t)(nout - m.Cc) * 4, w.rgc, (size_t)nout * 4, (size_t)(nout - m.Cc) * 4, B,
                             cudaMemcpyDeviceToDevice, st));
  VXB_CUDA(cudaMemcpy2DAsync(collision, (size_t)m.Cc * 4, w.rgc + (nout - m.Cc), (size_t)nout * 4, (size_t)m.Cc * 4, B,
                             cudaMemcpyDeviceToDevice, st));
  if (d->arm_pred_loss && arm_out) {
    VXB_TRY(lin(cx, w.feats, m.flat, P(VXB_P_DENSE2_W), m.flat, P(VXB_P_DENSE2_B), nullptr, 1, 0, w.h2, 64, B, 64, m.flat, 1.f,
                slope, VXB_MATH_FP32_SIMT));
    VXB_TRY(lin(cx, w.h2, 64, P(VXB_P_ARM_W), 64, P(VXB_P_ARM_B), nullptr, 1, 0, arm_out, 2, B, 2, 64, 1.f, -1.f,
                VXB_MATH_FP32_SIMT));
  }
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------ backward
// grads: HOST array parallel to `params` (vxb_param_slot order) of device pointers that receive d loss / d param
// (assigned, not accumulated); null entries are skipped.
struct Grads {
  float* const* g;
  float* at(int slot) const { return g[slot]; }
  float* layer(int l, int slot) const { return g[VXB_P_FIXED_COUNT + l * VXB_P_LAYER_STRIDE + slot]; }
};

// y = act(x W^T + b): given gy (overwritten by the pre-activation gradient), returns dW, db and optionally gx (=|+=)
static int linear_bwd(float* gy, const float* y_or_null, float slope, const float* x, int ldx, const float* W, int ldw,
                      float* dW, float* db, float* gx, int ldgx, bool gx_accumulate, long long M, int N, int K, cudaStream_t st) {
  // last layer of the chain (no input gradient wanted), 64 outputs, a handful of inputs, millions of rows -- input_preprocess:
  // one streaming kernel forms dW and db with the LeakyReLU adjoint applied on the way in (gy is left untouched)
  if (dW && !gx && M < (1ll << 31) && bwd::wgrad_tall64_ok(gy, N, N, K, (int)M) && (!y_or_null || (reinterpret_cast<uintptr_t>(y_or_null) & 7) == 0))
    return bwd::wgrad_tall64(gy, N, x, ldx, dW, K, K, M, false, st, (y_or_null && slope >= 0.f) ? y_or_null : nullptr, slope, db);
  if (y_or_null) VXB_TRY(bwd::lrelu_bwd(gy, y_or_null, M * N, slope, st));
  if (dW) VXB_TRY(bwd::gemm_tn(gy, N, x, ldx, dW, K, N, K, (int)M, false, st));
  if (db) VXB_TRY(bwd::colsum(gy, N, M, N, db, false, st));
  if (gx) VXB_TRY(bwd::gemm_nn(gy, N, W, ldw, gx, ldgx, (int)M, K, N, gx_accumulate, st));
  return VXB_OK;
}

// x_out = x_mid + FF(LN(x_mid)); on entry g_x = d loss / d x_out, on exit g_x = d loss / d x_mid
static int ff_bwd(const Dims& m, int B, TrainBufs& t, const BlockSaved& s, const float* nw, const float* W0, const float* W2,
                  float* d_nw, float* d_nb, float* dW0, float* db0, float* dW2, float* db2, cudaStream_t st) {
  const long long rows = (long long)B * m.L;
  // net.2: [rows, 4D] -> [rows, D]
  VXB_TRY(bwd::gemm_tn(t.g_x, m.D, s.ffg, 4 * m.D, dW2, 4 * m.D, m.D, 4 * m.D, (int)rows, false, st));
  VXB_TRY(bwd::colsum(t.g_x, m.D, rows, m.D, db2, false, st));
  VXB_TRY(bwd::gemm_nn(t.g_x, m.D, W2, 4 * m.D, t.g_ffg, 4 * m.D, (int)rows, 4 * m.D, m.D, false, st));
  VXB_TRY(bwd::geglu_bwd(t.g_ffg, s.ffh, t.g_ffh, rows, 4 * m.D, st));
  // net.0: [rows, D] -> [rows, 8D]
  VXB_TRY(bwd::gemm_tn(t.g_ffh, 8 * m.D, s.xn_f, m.D, dW0, m.D, 8 * m.D, m.D, (int)rows, false, st));
  VXB_TRY(bwd::colsum(t.g_ffh, 8 * m.D, rows, 8 * m.D, db0, false, st));
  VXB_TRY(bwd::gemm_nn(t.g_ffh, 8 * m.D, W0, m.D, t.g_xn, m.D, (int)rows, m.D, 8 * m.D, false, st));
  return bwd::layernorm_bwd(t.g_xn, s.x_mid, 0, (int)rows, nw, t.g_x, true, d_nw, d_nb, rows, m.D, st);
}

static int qnet_backward_impl(const vxb_qnet_desc* d, const Dims& m, const void* const* params, const Prepared& pw, Work& w,
                              TrainBufs& t, const float* grid, const float* proprio, const float* lang_tokens, int B,
                              const float* g_trans, const float* g_rot_grip, const float* g_collision, const float* g_arm,
                              const Grads& G, float* const* dbg, const TrainDropout& drop, cudaStream_t st) {
  auto P = [&](int slot) { return (const float*)params[slot]; };
  auto PL = [&](int layer, int slot) {
    return (const float*)params[VXB_P_FIXED_COUNT + layer * VXB_P_LAYER_STRIDE + slot];
  };
  auto DBG = [&](int i, const float* src, size_t n) -> int {
    if (dbg && dbg[i]) VXB_CUDA(cudaMemcpyAsync(dbg[i], src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return VXB_OK;
  };
  bwd::g_tc = bwd::TensorCtx{d->math_mode, d->math_mode == VXB_MATH_F16X3 ? (void*)t.tc_scratch : nullptr, t.tc_scratch_bytes};
  const float slope = d->act_slope;
  const int cq = m.ch * m.cdh, lq = m.lh * m.ldh;
  const long long rowsL = (long long)B * m.L;
  const int nout = 3 * m.R + m.G + m.Cc, nrg = nout - m.Cc;
  const int k3 = m.k * m.k * m.k, s3 = m.s * m.s * m.s;
  const size_t full = (size_t)B * m.V3 * 64;
  const int off = 256 + 4 * m.C;

  // ---- dgrad weights from the current parameters
  bwd::conv_dgrad_weight_kernel<<<148 * 2, 256, 0, st>>>(P(VXB_P_FINAL_W), t.wd_final, 64, 128, 27);
  bwd::conv_dgrad_weight_kernel<<<148 * 4, 256, 0, st>>>(P(VXB_P_UP0_W), t.wd_up0, 64, m.C, k3);
  bwd::fold_gemm_weight_kernel<<<148 * 8, 256, 0, st>>>(pw.up1_fold, t.wd_fold, s3, 64, 64);
  bwd::patch_dgrad_weight_kernel<<<148 * 2, 256, 0, st>>>(P(VXB_P_PATCH_W), t.wd_patch, 64, 64, k3);
  VXB_LAUNCH_CHECK();

  // ---- heads                                                                       perceiver_lang_io.py:472-483
  VXB_CUDA(cudaMemcpy2DAsync(t.g_rgc, (size_t)nout * 4, g_rot_grip, (size_t)nrg * 4, (size_t)nrg * 4, B, cudaMemcpyDeviceToDevice, st));
  VXB_CUDA(cudaMemcpy2DAsync(t.g_rgc + nrg, (size_t)nout * 4, g_collision, (size_t)m.Cc * 4, (size_t)m.Cc * 4, B,
                             cudaMemcpyDeviceToDevice, st));
  VXB_TRY(linear_bwd(t.g_rgc, nullptr, -1.f, w.h1, 64, P(VXB_P_RGC_W), 64, G.at(VXB_P_RGC_W), G.at(VXB_P_RGC_B), t.g_h1, 64,
                     false, B, nout, 64, st));
  VXB_TRY(linear_bwd(t.g_h1, w.h1, slope, w.h0, 256, P(VXB_P_DENSE1_W), 256, G.at(VXB_P_DENSE1_W), G.at(VXB_P_DENSE1_B), t.g_h0,
                     256, false, B, 64, 256, st));
  VXB_TRY(linear_bwd(t.g_h0, w.h0, slope, w.feats, m.flat, P(VXB_P_DENSE0_W), m.flat, G.at(VXB_P_DENSE0_W), G.at(VXB_P_DENSE0_B),
                     t.g_feats, m.flat, false, B, 256, m.flat, st));
  if (d->arm_pred_loss && g_arm) {
    float* g_armc = t.tmp_small;   // [B,2] copy: linear_bwd may rewrite its gy
    VXB_CUDA(cudaMemcpyAsync(g_armc, g_arm, (size_t)B * 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    VXB_TRY(linear_bwd(g_armc, nullptr, -1.f, w.h2, 64, P(VXB_P_ARM_W), 64, G.at(VXB_P_ARM_W), G.at(VXB_P_ARM_B), t.g_h2, 64, false,
                       B, 2, 64, st));
    VXB_TRY(linear_bwd(t.g_h2, w.h2, slope, w.feats, m.flat, P(VXB_P_DENSE2_W), m.flat, G.at(VXB_P_DENSE2_W),
                       G.at(VXB_P_DENSE2_B), t.g_feats, m.flat, true, B, 64, m.flat, st));
  }
  VXB_TRY(DBG(0, t.g_feats, (size_t)B * m.flat));

  // ---- ss_final / max-pool over u, trans decoder                                   :465-470
  VXB_TRY(bwd::ss_bwd(w.u, t.statsF, w.feats + off, m.flat, t.g_feats + off, m.flat, t.g_feats + off + 192, m.flat, t.argF, t.g_u,
                      false, B, m.V, m.V, m.V, 64, st));
  VXB_CUDA(cudaMemsetAsync(t.dwt, 0, (size_t)27 * 64 * sizeof(float), st));
  if ((long long)B * m.V3 >= (1ll << 32)) {
    set_error("qnet_backward: B * V^3 = %lld voxels exceed the 32-bit voxel index of the trans_decoder adjoint", (long long)B * m.V3);
    return VXB_E_UNSUPPORTED_SHAPE;
  }
  bwd::trans_bwd_kernel<64><<<148 * 3, 128, 0, st>>>(g_trans, w.u, pw.trans_wt, t.g_u, 1, t.dwt, B, m.V);
  VXB_LAUNCH_CHECK();
  if (G.at(VXB_P_TRANS_W)) {
    bwd::wgrad_to_torch_kernel<<<8, 256, 0, st>>>(t.dwt, G.at(VXB_P_TRANS_W), 1, 64, 27);
    VXB_LAUNCH_CHECK();
  }
  if (G.at(VXB_P_TRANS_B)) VXB_TRY(bwd::colsum(g_trans, 1, (long long)B * m.V3, 1, G.at(VXB_P_TRANS_B), false, st));
  VXB_TRY(DBG(1, t.g_u, full));

  // ---- final conv 128 -> 64 on cat[d0, u0]                                          :462
  VXB_TRY(bwd::lrelu_bwd(t.g_u, w.u, (long long)full, slope, st));            // g_u := gradient of the pre-activation
  if (G.at(VXB_P_FINAL_B)) VXB_TRY(bwd::colsum(t.g_u, 64, (long long)B * m.V3, 64, G.at(VXB_P_FINAL_B), false, st));
  if (G.at(VXB_P_FINAL_W)) {
    int rc = 1;
    if (bwd::g_tc.mm == VXB_MATH_F16X3 && bwd::g_tc.scratch) {
      Arena local(bwd::g_tc.scratch, bwd::g_tc.scratch_bytes);
      rc = umma::conv3_wgrad_f32(w.d0, w.u0, t.g_u, t.dwt, B, m.V, local, st);
      if (rc == VXB_E_WORKSPACE_TOO_SMALL) rc = 1;
      if (rc < 0) return rc;
    }
    if (rc == 1) VXB_TRY(bwd::conv_wgrad(w.d0, w.u0, 64, 64, t.g_u, 64, t.dwt, B, m.V, m.V, 3, 1, st));
    bwd::wgrad_to_torch_kernel<<<148 * 2, 256, 0, st>>>(t.dwt, G.at(VXB_P_FINAL_W), 64, 128, 27);
    VXB_LAUNCH_CHECK();
  }
  {
    const bwd::FoldDst dst[2] = {{t.g_d0, 64, 0}, {t.g_u0, 64, 0}};
    VXB_TRY(bwd::conv_dgrad_fold(t.g_u, 64, t.wd_final, 128, t.gxp_big, B, m.V, 3, dst, false, st));
  }
  VXB_TRY(DBG(2, t.g_u0, full));

  // ---- up0, second half: conv_k o upsample_s in its folded (polyphase) form         network_utils.py:245-251
  VXB_TRY(bwd::lrelu_bwd(t.g_u0, w.u0, (long long)full, slope, st));
  if (G.at(VXB_P_UP1_B)) VXB_TRY(bwd::colsum(t.g_u0, 64, (long long)B * m.V3, 64, G.at(VXB_P_UP1_B), false, st));
  bwd::phase_gather_kernel<<<148 * 16, 256, 0, st>>>(t.g_u0, t.g_ph, B, m.S, m.s, 64);
  VXB_LAUNCH_CHECK();
  if (G.at(VXB_P_UP1_W)) {
    // dwt[(nb, ci)][(r, co)] = im2col(low)^T g_ph: one GEMM with K = B * S^3 (explicit low-resolution im2col, 27 x 64 columns)
    bwd::im2col3_kernel<<<148 * 8, 256, 0, st>>>(w.low, t.im2col, B, m.S, 64);
    VXB_LAUNCH_CHECK();
    VXB_TRY(bwd::gemm_tn(t.im2col, 27 * 64, t.g_ph, s3 * 64, t.dwt, s3 * 64, 27 * 64, s3 * 64, B * m.T, false, st));
    bwd::fold_upconv_weights_bwd_kernel<<<148 * 8, 256, 0, st>>>(t.dwt, G.at(VXB_P_UP1_W), 64, 64, m.k, m.s);
    VXB_LAUNCH_CHECK();
  }
  // dgrad as GEMM + col2im: T[q][(nb, ci)] = g_ph[q] . wt2[(nb, ci)], then every low-resolution voxel gathers the taps that
  // touched it (replicate-padding adjoint included)
  VXB_TRY(bwd::gemm_nt(t.g_ph, s3 * 64, t.wd_fold, s3 * 64, t.im2col, 27 * 64, B * m.T, 27 * 64, s3 * 64, false, st));
  bwd::col2im3_fold_kernel<<<148 * 8, 256, 0, st>>>(t.im2col, t.g_low, B, m.S, 64);
  VXB_LAUNCH_CHECK();
  VXB_TRY(DBG(3, t.g_low, (size_t)B * m.T * 64));
  // ---- up0, first half: conv k (C -> 64) at S^3
  VXB_TRY(bwd::lrelu_bwd(t.g_low, w.low, (long long)B * m.T * 64, slope, st));
  if (G.at(VXB_P_UP0_B)) VXB_TRY(bwd::colsum(t.g_low, 64, (long long)B * m.T, 64, G.at(VXB_P_UP0_B), false, st));
  if (G.at(VXB_P_UP0_W)) {
    VXB_TRY(bwd::conv_wgrad(w.dec, nullptr, m.C, 0, t.g_low, 64, t.dwt, B, m.S, m.S, m.k, 1, st));
    bwd::wgrad_to_torch_kernel<<<148 * 4, 256, 0, st>>>(t.dwt, G.at(VXB_P_UP0_W), 64, m.C, k3);
    VXB_LAUNCH_CHECK();
  }
  {
    bwd::FoldDst dst[8];
    for (int j = 0; j < m.C / 64; ++j) dst[j] = bwd::FoldDst{t.g_dec, m.C, 64 * j};
    VXB_TRY(bwd::conv_dgrad_fold(t.g_low, 64, t.wd_up0, m.C, t.gxp_low, B, m.S, m.k, dst, false, st));
  }
  // ---- ss1 / max-pool over dec                                                       :451
  VXB_TRY(bwd::ss_bwd(w.dec, t.stats1, w.feats + 256, m.flat, t.g_feats + 256, m.flat, t.g_feats + 256 + 3 * m.C, m.flat, t.arg1,
                      t.g_dec, true, B, m.S, m.S, m.S, m.C, st));
  VXB_TRY(DBG(4, t.g_dec, (size_t)B * m.T * m.C));

  // ---- decoder cross attention                                                       :440-448
  VXB_TRY(linear_bwd(t.g_dec, nullptr, -1.f, t.att_d, cq, P(VXB_P_DEC_OUT_W), cq, G.at(VXB_P_DEC_OUT_W), G.at(VXB_P_DEC_OUT_B),
                     t.g_att_d, cq, false, (long long)B * m.T, m.C, cq, st));
  VXB_TRY(bwd::attention_bwd(t.qd, cq, (long long)m.T * cq, t.kv_d, t.kv_d + cq, 2 * cq, (long long)m.L * 2 * cq, t.g_att_d, cq,
                             (long long)m.T * cq, t.g_qd, cq, (long long)m.T * cq, t.g_kv_d, t.g_kv_d + cq, 2 * cq,
                             (long long)m.L * 2 * cq, B, m.ch, m.T, m.L, m.cdh, 1.f / sqrtf((float)m.cdh), t.simA, t.simB,
                             t.simC, layer_dropout(drop, -1), st));
  VXB_TRY(linear_bwd(t.g_qd, nullptr, -1.f, t.qn, m.C, P(VXB_P_DEC_Q_W), m.C, G.at(VXB_P_DEC_Q_W), nullptr, t.g_qn, m.C, false,
                     (long long)B * m.T, cq, m.C, st));
  VXB_CUDA(cudaMemsetAsync(t.g_ins, 0, (size_t)B * m.n * m.C * sizeof(float), st));   // language rows get nothing from the decoder
  VXB_TRY(bwd::layernorm_bwd(t.g_qn, w.ins + (size_t)m.nl * m.C, (size_t)m.n * m.C, m.T, P(VXB_P_DEC_NORM_W),
                             t.g_ins + (size_t)m.nl * m.C, false, G.at(VXB_P_DEC_NORM_W), G.at(VXB_P_DEC_NORM_B),
                             (long long)B * m.T, m.C, st));
  VXB_TRY(linear_bwd(t.g_kv_d, nullptr, -1.f, t.xn_d, m.D, P(VXB_P_DEC_KV_W), m.D, G.at(VXB_P_DEC_KV_W), nullptr, t.g_xn, m.D,
                     false, rowsL, 2 * cq, m.D, st));
  VXB_TRY(bwd::layernorm_bwd(t.g_xn, t.x_out, 0, (int)rowsL, P(VXB_P_DEC_NORMCTX_W), t.g_x, false, G.at(VXB_P_DEC_NORMCTX_W),
                             G.at(VXB_P_DEC_NORMCTX_B), rowsL, m.D, st));
  VXB_TRY(DBG(5, t.g_x, (size_t)rowsL * m.D));

  // ---- latent self-attention stack, reversed                                          :435-437
  for (int l = m.depth - 1; l >= 0; --l) {
    const BlockSaved& s = t.blk[l + 1];
    VXB_TRY(ff_bwd(m, B, t, s, PL(l, VXB_PL_FF_NORM_W), PL(l, VXB_PL_FF0_W), PL(l, VXB_PL_FF2_W), G.layer(l, VXB_PL_FF_NORM_W),
                   G.layer(l, VXB_PL_FF_NORM_B), G.layer(l, VXB_PL_FF0_W), G.layer(l, VXB_PL_FF0_B), G.layer(l, VXB_PL_FF2_W),
                   G.layer(l, VXB_PL_FF2_B), st));
    // x_mid = x_in + to_out(attn(...)): g_x is d/d x_mid
    VXB_TRY(bwd::gemm_tn(t.g_x, m.D, s.att, lq, G.layer(l, VXB_PL_OUT_W), lq, m.D, lq, (int)rowsL, false, st));
    VXB_TRY(bwd::colsum(t.g_x, m.D, rowsL, m.D, G.layer(l, VXB_PL_OUT_B), false, st));
    VXB_TRY(bwd::gemm_nn(t.g_x, m.D, PL(l, VXB_PL_OUT_W), lq, t.g_att, lq, (int)rowsL, lq, m.D, false, st));
    VXB_TRY(bwd::attention_bwd(s.q, lq, (long long)m.L * lq, s.kv, s.kv + lq, 2 * lq, (long long)m.L * 2 * lq, t.g_att, lq,
                               (long long)m.L * lq, t.g_q, lq, (long long)m.L * lq, t.g_kv, t.g_kv + lq, 2 * lq,
                               (long long)m.L * 2 * lq, B, m.lh, m.L, m.L, m.ldh, 1.f / sqrtf((float)m.ldh), t.simA, t.simB,
                               t.simC, layer_dropout(drop, l + 1), st));
    VXB_TRY(bwd::gemm_tn(t.g_q, lq, s.xn_a, m.D, G.layer(l, VXB_PL_Q_W), m.D, lq, m.D, (int)rowsL, false, st));
    VXB_TRY(bwd::gemm_tn(t.g_kv, 2 * lq, s.xn_a, m.D, G.layer(l, VXB_PL_KV_W), m.D, 2 * lq, m.D, (int)rowsL, false, st));
    VXB_TRY(bwd::gemm_nn(t.g_q, lq, PL(l, VXB_PL_Q_W), m.D, t.g_xn, m.D, (int)rowsL, m.D, lq, false, st));
    VXB_TRY(bwd::gemm_nn(t.g_kv, 2 * lq, PL(l, VXB_PL_KV_W), m.D, t.g_xn, m.D, (int)rowsL, m.D, 2 * lq, true, st));
    VXB_TRY(bwd::layernorm_bwd(t.g_xn, s.x_in, 0, (int)rowsL, PL(l, VXB_PL_ATTN_NORM_W), t.g_x, true,
                               G.layer(l, VXB_PL_ATTN_NORM_W), G.layer(l, VXB_PL_ATTN_NORM_B), rowsL, m.D, st));
  }
  // ---- encoder cross attention block                                                  :431-432
  {
    const BlockSaved& s = t.blk[0];
    VXB_TRY(ff_bwd(m, B, t, s, P(VXB_P_CROSS_FF_NORM_W), P(VXB_P_CROSS_FF0_W), P(VXB_P_CROSS_FF2_W), G.at(VXB_P_CROSS_FF_NORM_W),
                   G.at(VXB_P_CROSS_FF_NORM_B), G.at(VXB_P_CROSS_FF0_W), G.at(VXB_P_CROSS_FF0_B), G.at(VXB_P_CROSS_FF2_W),
                   G.at(VXB_P_CROSS_FF2_B), st));
    // x_mid = latents (broadcast) + to_out(attn)
    VXB_TRY(bwd::batch_sum(t.g_x, (long long)m.L * m.D, B, (long long)m.L * m.D, G.at(VXB_P_LATENTS), false, st));
    VXB_TRY(bwd::gemm_tn(t.g_x, m.D, s.att, cq, G.at(VXB_P_CROSS_OUT_W), cq, m.D, cq, (int)rowsL, false, st));
    VXB_TRY(bwd::colsum(t.g_x, m.D, rowsL, m.D, G.at(VXB_P_CROSS_OUT_B), false, st));
    VXB_TRY(bwd::gemm_nn(t.g_x, m.D, P(VXB_P_CROSS_OUT_W), cq, t.g_att, cq, (int)rowsL, cq, m.D, false, st));
    VXB_TRY(bwd::attention_bwd(pw.q_cross, cq, 0, w.kv_c, w.kv_c + cq, 2 * cq, (long long)m.n * 2 * cq, t.g_att, cq,
                               (long long)m.L * cq, t.g_qcb, cq, (long long)m.L * cq, t.g_kv_c, t.g_kv_c + cq, 2 * cq,
                               (long long)m.n * 2 * cq, B, m.ch, m.L, m.n, m.cdh, 1.f / sqrtf((float)m.cdh), t.simA, t.simB,
                               t.simC, layer_dropout(drop, 0), st));
    // q = to_q(LN(latents)) is shared by the batch
    VXB_TRY(bwd::batch_sum(t.g_qcb, (long long)m.L * cq, B, (long long)m.L * cq, t.g_qc, false, st));
    VXB_TRY(bwd::gemm_tn(t.g_qc, cq, pw.lat_norm, m.D, G.at(VXB_P_CROSS_Q_W), m.D, cq, m.D, m.L, false, st));
    VXB_TRY(bwd::gemm_nn(t.g_qc, cq, P(VXB_P_CROSS_Q_W), m.D, t.g_latn, m.D, m.L, m.D, cq, false, st));
    VXB_TRY(bwd::layernorm_bwd(t.g_latn, P(VXB_P_LATENTS), 0, m.L, P(VXB_P_CROSS_NORM_W), G.at(VXB_P_LATENTS), true,
                               G.at(VXB_P_CROSS_NORM_W), G.at(VXB_P_CROSS_NORM_B), m.L, m.D, st));
    // context = LN_ctx(ins)
    VXB_TRY(bwd::gemm_tn(t.g_kv_c, 2 * cq, w.ctx_n, m.C, G.at(VXB_P_CROSS_KV_W), m.C, 2 * cq, m.C, B * m.n, false, st));
    VXB_TRY(bwd::gemm_nn(t.g_kv_c, 2 * cq, P(VXB_P_CROSS_KV_W), m.C, t.g_ctx, m.C, B * m.n, m.C, 2 * cq, false, st));
    VXB_TRY(bwd::layernorm_bwd(t.g_ctx, w.ins, 0, B * m.n, P(VXB_P_CROSS_NORMCTX_W), t.g_ins, true, G.at(VXB_P_CROSS_NORMCTX_W),
                               G.at(VXB_P_CROSS_NORMCTX_B), (long long)B * m.n, m.C, st));
  }
  VXB_TRY(DBG(6, t.g_ins, (size_t)B * m.n * m.C));

  // ---- token assembly: pos_encoding, language projection, proprio, patch tokens        :370-422
  VXB_TRY(bwd::batch_sum(t.g_ins, (long long)m.n * m.C, B, (long long)m.n * m.C, G.at(VXB_P_POS_ENCODING), false, st));
  bwd::disassemble_tokens_kernel<<<148 * 8, 256, 0, st>>>(t.g_ins, t.g_lang, t.g_patch, B, m.nl, m.T, m.C, 64);
  VXB_LAUNCH_CHECK();
  bwd::proprio_token_sum_kernel<<<B, 256, 0, st>>>(t.g_ins, t.g_pfeat, m.nl, m.T, m.C, 64, 64);
  VXB_LAUNCH_CHECK();
  if (d->no_language) {
    VXB_CUDA(cudaMemsetAsync(G.at(VXB_P_LANG_W), 0, (size_t)m.C * d->lang_emb_dim * sizeof(float), st));
    VXB_TRY(bwd::colsum(t.g_lang, m.C, (long long)B * m.nl, m.C, G.at(VXB_P_LANG_B), false, st));
  } else {
    VXB_TRY(linear_bwd(t.g_lang, nullptr, -1.f, lang_tokens, d->lang_emb_dim, P(VXB_P_LANG_W), d->lang_emb_dim,
                       G.at(VXB_P_LANG_W), G.at(VXB_P_LANG_B), nullptr, 0, false, (long long)B * m.nl, m.C, d->lang_emb_dim, st));
  }
  VXB_TRY(linear_bwd(t.g_pfeat, w.pfeat, slope, proprio, m.low, P(VXB_P_PROPRIO_W), m.low, G.at(VXB_P_PROPRIO_W),
                     G.at(VXB_P_PROPRIO_B), nullptr, 0, false, B, 64, m.low, st));

  // ---- patchify (k == s: non-overlapping windows)                                       :363
  VXB_TRY(bwd::lrelu_bwd(t.g_patch, w.patch, (long long)B * m.T * 64, slope, st));
  if (G.at(VXB_P_PATCH_B)) VXB_TRY(bwd::colsum(t.g_patch, 64, (long long)B * m.T, 64, G.at(VXB_P_PATCH_B), false, st));
  if (G.at(VXB_P_PATCH_W)) {
    VXB_TRY(bwd::conv_wgrad(w.d0, nullptr, 64, 0, t.g_patch, 64, t.dwt, B, m.V, m.S, m.k, m.s, st));
    bwd::wgrad_to_torch_kernel<<<148 * 4, 256, 0, st>>>(t.dwt, G.at(VXB_P_PATCH_W), 64, 64, k3);
    VXB_LAUNCH_CHECK();
  }
  {
    // padded-input gradient of window (o, t): gxp[s o + t] = W_t^T gz[o]  -> one GEMM per tap, scattered as phases
    float* gxp = t.g_u;                                  // g_u is dead by now; same size ([B, (S s)^3, 64])
    GemmParams p;
    gemm_params_init(p);
    p.M = B * m.T; p.N = 64; p.K = 64;
    p.src0 = t.g_patch; p.src1 = nullptr; p.C0 = 64; p.C1 = 0;
    p.Di = m.S; p.Do = m.S; p.kk = 1; p.cstride = 1; p.pad = 0;
    p.W = t.wd_patch; p.ldw = 64; p.w_stride_zb = 64 * 64;
    p.C = gxp; p.ldc = 64;
    p.ps = m.s;
    VXB_TRY((launch_simt_gemm<A_CONV, B_NT, O_PHASE>(p, s3, st)));
    VXB_TRY(bwd::fold_pad(gxp, 64, 0, m.S * m.s, m.k / 2, t.g_d0, 64, m.V, B, true, st));
  }
  // ---- ss0 / max-pool over d0, input_preprocess                                          :357-360
  VXB_TRY(bwd::ss_bwd(w.d0, t.stats0, w.feats, m.flat, t.g_feats, m.flat, t.g_feats + 192, m.flat, t.arg0, t.g_d0, true, B, m.V,
                      m.V, m.V, 64, st));
  VXB_TRY(DBG(7, t.g_d0, full));
  VXB_TRY(linear_bwd(t.g_d0, w.d0, slope, grid, d->initial_dim, P(VXB_P_INPRE_W), d->initial_dim, G.at(VXB_P_INPRE_W),
                     G.at(VXB_P_INPRE_B), nullptr, 0, false, (long long)B * m.V3, 64, d->initial_dim, st));
  return VXB_OK;
}

}  // namespace vxb

/*
 * voxactb.h -- C ABI of libvoxactb.so: the B200 (sm_100a) voxel-policy hot path.
 *
 * Drop-in boundary for the VoxAct-B / PerAct hot path:
 *   - vxb_voxelize_f32      replaces VoxelGrid.coords_to_bounding_voxel_grid
 *                           (reference peract/voxel/voxel_grid.py:148-198)
 *   - vxb_qnet_forward_f32  replaces PerceiverVoxelLangEncoder.forward
 *                           (reference peract/agents/peract_bc/perceiver_lang_io.py:345-485)
 *   - vxb_qnet_prepare      one-off weight re-layout for the above (conv weights to
 *                           tap-major, upsample-conv folding, bf16 hi/lo split)
 *   - vxb_select_action_f32 replaces QFunction._argmax_3d / choose_highest_action and the
 *                           act() tail (reference qattention_peract_bc_agent.py:57-80,709-724)
 *
 * Conventions: plain C types only; every pointer is a DEVICE pointer unless it says "host";
 * the caller owns every buffer (outputs and workspaces, sized by the *_bytes queries);
 * all work is enqueued asynchronously on the caller's stream (a cudaStream_t passed as void*);
 * return value 0 = OK, negative = vxb_status; vxb_last_error() gives a thread-local message.
 * There is no CPU fallback anywhere in this library.
 */
#ifndef VOXACTB_H_
#define VOXACTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VXB_VERSION 100

typedef enum vxb_status {
  VXB_OK = 0,
  VXB_E_BADARG = -1,
  VXB_E_UNSUPPORTED_SHAPE = -2,
  VXB_E_WORKSPACE_TOO_SMALL = -3,
  VXB_E_CUDA = -4,
  VXB_E_NO_DEVICE = -5,
  VXB_E_NCCL = -6
} vxb_status;

int vxb_version(void);
const char* vxb_last_error(void);
/* 0 if a CUDA device of compute capability 10.x is current, else VXB_E_NO_DEVICE. */
int vxb_check_device(void);

/* ------------------------------------------------------------------ voxelizer (K1) */
/* layout of the voxel grid output */
#define VXB_LAYOUT_CHANNELS_LAST 0 /* [B,V,V,V,3+F+3+1]  (what voxel_grid.py:196-198 returns) */

size_t vxb_voxelize_workspace_bytes(int B, int N, int V, int F);

/*
 * coords  [B,N,3] fp32 world-frame points, feats [B,N,F] fp32 (may be NULL when F==0),
 * bounds  [Bb,6]  fp32 (min xyz, max xyz), Bb is 1 (shared) or B (per-sample VLM crop),
 * out     [B,V,V,V,3+F+3+1] fp32,
 * out_idx NULL or [B,N,3] int32: the clamped (V+2)-grid voxel index of every point
 *         (voxel_grid.py:159-163), for the bit-exact parity test.
 */
int vxb_voxelize_f32(const float* coords, const float* feats, const float* bounds, int Bb,
                     int B, int N, int F, int V, float* out, int layout, int32_t* out_idx,
                     void* ws, size_t ws_bytes, void* stream);

/*
 * Raw-depth entry: the depth -> world point cloud back-projection that the reference does on
 * the host per camera per step (PyRep pyrep/objects/vision_sensor.py:155-175, 381-412) is fused into the scatter kernel, so
 * the path ingests depth images + camera matrices and the 3 MB/sample point cloud never crosses PCIe.
 * depth    [B,cams,H,W] fp32 metres; proj_inv [B,cams,3,4] FLOAT64 = inv(K [R^T | -R^T C])[0:3] (one 4x4 inverse per
 * camera on the host, as vision_sensor.py:165-171 does); rgb [B,cams,F,H,W] planar image features (NULL when F == 0).
 * Point n = cam*H*W + y*W + x: pc = (x d, y d, d) in fp32, world = proj_inv . (pc, 1) in float64, rounded to fp32 -- the
 * reference's arithmetic -- then voxelized exactly like vxb_voxelize_f32.  out_points: NULL or [B,N,3] (parity tests).
 */
int vxb_voxelize_depth_f32(const float* depth, const double* proj_inv, const float* rgb, const float* bounds, int Bb,
                           int B, int cams, int H, int W, int F, int V, float* out, int layout, float* out_points,
                           int32_t* out_idx, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ Q-network */
typedef struct vxb_qnet_desc {
  int32_t struct_bytes;     /* sizeof(vxb_qnet_desc), for ABI checking */
  int32_t voxel_size;       /* V */
  int32_t patch_size;       /* voxel_patch_size k (odd) */
  int32_t patch_stride;     /* voxel_patch_stride s; V % s == 0 and conv output == V/s required */
  int32_t initial_dim;      /* 10 */
  int32_t im_channels;      /* 64 */
  int32_t low_dim_size;     /* proprio width per arm (4 or 7) */
  int32_t two_robots;       /* 0: PerceiverVoxelLangEncoder, 1: ...2RobotsEncoder (C = 3*im) */
  int32_t lang_seq_len;     /* 77 */
  int32_t lang_emb_dim;     /* 512 */
  int32_t num_latents;      /* L */
  int32_t latent_dim;       /* D */
  int32_t depth;            /* self-attention layers */
  int32_t iterations;       /* cross-attention iterations */
  int32_t cross_heads, cross_dim_head, latent_heads, latent_dim_head;
  int32_t final_dim;        /* 64 */
  int32_t num_rotation_classes, num_grip_classes, num_collision_classes;
  int32_t arm_pred_loss;    /* 1: dense2/arm_ff head present and evaluated */
  int32_t no_language;      /* 1: language tokens zeroed (perceiver_lang_io.py:376-378) */
  float   act_slope;        /* 0.02 for 'lrelu', 0 for 'relu' */
  int32_t math_mode;        /* VXB_MATH_* */
  int32_t final_input;      /* input of the final 3x3x3 convolution (perceiver_lang_io.py:456-462): VXB_FINAL_CAT = cat[d0, u0]
                             * (128 channels, the default), VXB_FINAL_U0 = u0 only (no_skip_connection), VXB_FINAL_D0 = d0 only
                             * (no_perceiver); the two ablations are inference-only and run the split-16x3 convolution */
} vxb_qnet_desc;
#define VXB_FINAL_CAT 0
#define VXB_FINAL_U0  1
#define VXB_FINAL_D0  2

#define VXB_MATH_FP32_SIMT 0  /* fp32 FFMA everywhere (reference arithmetic, slow path for parity) */
#define VXB_MATH_F16X3     1  /* tcgen05 split-16-bit (hi*hi + hi*lo + lo*hi; fp16 planes), fp32 accumulate in TMEM */
#define VXB_MATH_BF16X3    VXB_MATH_F16X3  /* round-1 name (the planes were bf16 then); kept for source compatibility */
#define VXB_MATH_F16F8C    2  /* as VXB_MATH_F16X3, but the 3x3x3 final convolution forms x.w as fp16 hi*hi plus ONE E4M3 MMA that
                               * carries both 2^-11 correction terms (conv_f8c.cuh): 2 MMA units per product instead of 3, same
                               * accuracy class (the corrections only need ~4 bits).  Training entries treat it as VXB_MATH_F16X3. */

/* Parameter slots: device pointers to contiguous fp32 tensors with the reference's shapes
 * (perceiver_lang_io.py:137-334; state-dict names in the comments). */
enum vxb_param_slot {
  VXB_P_POS_ENCODING = 0,        /* pos_encoding [1,77+S^3,C] */
  VXB_P_LATENTS,                 /* latents [L,D] */
  VXB_P_INPRE_W, VXB_P_INPRE_B,  /* input_preprocess.conv3d [64,10,1,1,1],[64] */
  VXB_P_PATCH_W, VXB_P_PATCH_B,  /* patchify.conv3d [64,64,k,k,k],[64] */
  VXB_P_LANG_W, VXB_P_LANG_B,    /* lang_preprocess [C,512],[C] */
  VXB_P_PROPRIO_W, VXB_P_PROPRIO_B,   /* proprio_preprocess.linear [64,low],[64] (right arm for 2 robots) */
  VXB_P_PROPRIO2_W, VXB_P_PROPRIO2_B, /* 2 robots: proprio_preprocess_left_arm.linear, else NULL */
  VXB_P_CROSS_NORM_W, VXB_P_CROSS_NORM_B,       /* cross_attend_blocks.0.norm */
  VXB_P_CROSS_NORMCTX_W, VXB_P_CROSS_NORMCTX_B, /* cross_attend_blocks.0.norm_context */
  VXB_P_CROSS_Q_W,               /* cross_attend_blocks.0.fn.to_q [ch*dh, D] */
  VXB_P_CROSS_KV_W,              /* cross_attend_blocks.0.fn.to_kv [2*ch*dh, C] */
  VXB_P_CROSS_OUT_W, VXB_P_CROSS_OUT_B,         /* cross_attend_blocks.0.fn.to_out */
  VXB_P_CROSS_FF_NORM_W, VXB_P_CROSS_FF_NORM_B, /* cross_attend_blocks.1.norm */
  VXB_P_CROSS_FF0_W, VXB_P_CROSS_FF0_B,         /* cross_attend_blocks.1.fn.net.0 [8D,D] */
  VXB_P_CROSS_FF2_W, VXB_P_CROSS_FF2_B,         /* cross_attend_blocks.1.fn.net.2 [D,4D] */
  VXB_P_DEC_NORM_W, VXB_P_DEC_NORM_B,           /* decoder_cross_attn.norm (C) */
  VXB_P_DEC_NORMCTX_W, VXB_P_DEC_NORMCTX_B,     /* decoder_cross_attn.norm_context (D) */
  VXB_P_DEC_Q_W, VXB_P_DEC_KV_W, VXB_P_DEC_OUT_W, VXB_P_DEC_OUT_B, /* decoder_cross_attn.fn.* */
  VXB_P_UP0_W, VXB_P_UP0_B,      /* up0.conv_up.0.conv3d [64,C,k,k,k] */
  VXB_P_UP1_W, VXB_P_UP1_B,      /* up0.conv_up.2.conv3d [64,64,k,k,k] */
  VXB_P_FINAL_W, VXB_P_FINAL_B,  /* final.conv3d [64,128,3,3,3] */
  VXB_P_TRANS_W, VXB_P_TRANS_B,  /* trans_decoder.conv3d [1,64,3,3,3] */
  VXB_P_TRANS2_W, VXB_P_TRANS2_B,/* 2 robots: trans_decoder_left_arm.conv3d, else NULL */
  VXB_P_DENSE0_W, VXB_P_DENSE0_B,/* dense0.linear [256,flat] */
  VXB_P_DENSE1_W, VXB_P_DENSE1_B,/* dense1.linear [64,256] */
  VXB_P_RGC_W, VXB_P_RGC_B,      /* rot_grip_collision_ff.linear [3R+G+Cc,64] */
  VXB_P_DENSE2_W, VXB_P_DENSE2_B,/* arm head: dense2.linear [64,flat]; 2 robots: dense0_left_arm */
  VXB_P_ARM_W, VXB_P_ARM_B,      /* arm head: arm_ff.linear [2,64];    2 robots: rot_grip_collision_ff_left_arm */
  VXB_P_DENSE1L_W, VXB_P_DENSE1L_B, /* 2 robots: dense1_left_arm.linear, else NULL */
  VXB_P_FIXED_COUNT,
  /* followed by depth x VXB_P_LAYER_STRIDE per-layer slots */
  VXB_PL_ATTN_NORM_W = 0, VXB_PL_ATTN_NORM_B,   /* layers.i.0.norm */
  VXB_PL_Q_W, VXB_PL_KV_W, VXB_PL_OUT_W, VXB_PL_OUT_B, /* layers.i.0.fn.{to_q,to_kv,to_out} */
  VXB_PL_FF_NORM_W, VXB_PL_FF_NORM_B,           /* layers.i.1.norm */
  VXB_PL_FF0_W, VXB_PL_FF0_B, VXB_PL_FF2_W, VXB_PL_FF2_B, /* layers.i.1.fn.net.{0,2} */
  VXB_P_LAYER_STRIDE
};

/* number of entries the params array must hold for this descriptor */
int vxb_qnet_num_params(const vxb_qnet_desc* d);

/* bytes of the "prepared weights" arena and of the per-call workspace for batch B */
size_t vxb_qnet_prepared_bytes(const vxb_qnet_desc* d);
size_t vxb_qnet_workspace_bytes(const vxb_qnet_desc* d, int B);

/* Re-layout / fold / split the weights into `prepared` (device, 256-byte aligned). Must be re-run
 * whenever a parameter changes. params: HOST array of device pointers (vxb_param_slot order). */
int vxb_qnet_prepare(const vxb_qnet_desc* d, const void* const* params, void* prepared,
                     size_t prepared_bytes, void* stream);

/*
 * grid        [B,V,V,V,initial_dim] fp32 channels-last (the voxelizer's native output; the
 *             reference's [B,10,V,V,V] argument is the permuted view of this memory),
 * proprio     [B,low_dim_size]; proprio2 = left arm for two_robots else NULL,
 * lang_tokens [B,77,512],
 * q_trans     [B,1,V,V,V]; q_trans2 = left arm grid for two_robots else NULL,
 * rot_grip    [B,3R+G], collision [B,Cc] (NULL allowed when num_rotation_classes==0),
 * rot_grip2/collision2: left-arm heads for two_robots else NULL,
 * arm_out     [B,2] when arm_pred_loss else NULL.
 */
int vxb_qnet_forward_f32(const vxb_qnet_desc* d, const void* const* params, const void* prepared,
                         const float* grid, const float* proprio, const float* proprio2,
                         const float* lang_tokens, int B,
                         float* q_trans, float* q_trans2, float* rot_grip, float* collision,
                         float* rot_grip2, float* collision2, float* arm_out,
                         void* ws, size_t ws_bytes, void* stream);

/* number of this library's kernels enqueued by the most recent vxb_qnet_forward_f32 on the calling
 * thread, and by one vxb_voxelize_f32 call (bench.py's gpu_launches) */
int vxb_last_launch_count(void);
int vxb_voxelize_launches(void);

/* stage timing of vxb_qnet_forward_f32 with CUDA events on the caller's stream (bench.py's live
 * roofline numbers).  vxb_profile_read ADDS the per-stage milliseconds of every forward since the
 * last read into ms[0..vxb_profile_stage_count()) and returns how many forwards it consumed. */
int vxb_profile_stage_count(void);
const char* vxb_profile_stage_name(int i);
int vxb_profile_enable(int on);
int vxb_profile_read(double* ms /* host */);

/* ------------------------------------------------------------------ training step of the Q-network (row a18)
 * Device side of `total_loss.backward()` in QAttentionPerActBCAgent.update (reference
 * qattention_peract_bc_agent.py:484-582) for PerceiverVoxelLangEncoder.forward (perceiver_lang_io.py:345-485):
 * vxb_qnet_forward_train_f32 runs the forward and keeps the activations in `ws`; vxb_qnet_backward_f32, called
 * with the SAME ws / params / prepared / inputs / opts, writes d loss / d parameter for every parameter.  The voxel
 * grid is detached (agent:107-108): nothing is propagated past input_preprocess.  Single-arm encoder (optionally
 * with the arm head), iterations == 1, voxel_patch_size == voxel_patch_stride. */
typedef struct vxb_train_opts {
  int32_t struct_bytes;          /* sizeof(vxb_train_opts) */
  float input_dropout;           /* nn.Dropout on the attention probabilities of the encoder cross attention */
  float attn_dropout;            /* ... of the latent self-attention layers */
  float decoder_dropout;         /* ... of the decoder cross attention (perceiver_lang_io.py:127-128,258,265,284) */
  uint64_t seed;                 /* counter-based mask stream; the backward regenerates the forward's masks from it */
} vxb_train_opts;

size_t vxb_qnet_train_workspace_bytes(const vxb_qnet_desc* d, int B);
int vxb_qnet_forward_train_f32(const vxb_qnet_desc* d, const void* const* params, const void* prepared,
                               const float* grid, const float* proprio, const float* lang_tokens, int B,
                               float* q_trans, float* rot_grip, float* collision, float* arm_out,
                               const vxb_train_opts* opts, void* ws, size_t ws_bytes, void* stream);
/* g_trans [B,V^3], g_rot_grip [B,3R+G], g_collision [B,Cc], g_arm [B,2] or NULL: d loss / d output.
 * grads: HOST array parallel to `params` (vxb_param_slot order) of device buffers with the parameters' shapes;
 * every gradient is ASSIGNED (not accumulated).  debug: NULL, or a HOST array of 8 device buffers (NULL entries
 * allowed) receiving activation gradients for the parity tests: 0 feats [B,flat], 1 u [B,V^3,64], 2 u0 [B,V^3,64],
 * 3 low [B,S^3,64], 4 dec [B,S^3,C], 5 final latents [B,L,D], 6 tokens [B,77+S^3,C], 7 d0 [B,V^3,64]
 * (1-3 and 7: gradient of the block OUTPUT, i.e. before the activation adjoint). */
int vxb_qnet_backward_f32(const vxb_qnet_desc* d, const void* const* params, const void* prepared,
                          const float* grid, const float* proprio, const float* lang_tokens, int B,
                          const float* g_trans, const float* g_rot_grip, const float* g_collision, const float* g_arm,
                          float* const* grads, const vxb_train_opts* opts, float* const* debug, void* ws,
                          size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ action selection */
/*
 * softmax-free argmax of the translation grid (softmax is monotone: agent:709-718), per-axis-group
 * argmax of rot/grip, collision argmax, and the metric attention coordinate
 * bounds_min + res*idx + res/2 (agent:724).  q_trans [B,V^3]; rot_grip [B,3R+2]; collision [B,2];
 * bounds [Bb,6].  Outputs: coords [B,3] int32, rot_grip_idx [B,4] int32, coll_idx [B] int32,
 * attention_xyz [B,3] fp32.
 */
size_t vxb_select_action_workspace_bytes(int B, int V);
int vxb_select_action_f32(const float* q_trans, const float* rot_grip, const float* collision,
                          const float* bounds, int Bb, int B, int V, int R,
                          int32_t* coords, int32_t* rot_grip_idx, int32_t* coll_idx,
                          float* attention_xyz, void* ws, size_t ws_bytes, void* stream);

/* act() tail: the 9-D continuous action QAttentionStackAgent.act assembles on the host
 * (reference qattention_stack_agent.py:78-89): action [B,9] = [attention_xyz(3), quaternion x,y,z,w (4) =
 * Rotation.from_euler('xyz', rot_idx * rotation_resolution - 180, degrees=True).as_quat() (helpers/utils.py:103-105),
 * gripper index (1), ignore-collision index (1)], from vxb_select_action_f32's outputs, on the device. */
int vxb_act_tail_f32(const int32_t* rot_grip_idx /*[B,4]*/, const int32_t* coll_idx /*[B]*/,
                     const float* attention_xyz /*[B,3]*/, float rotation_resolution, float* action /*[B,9]*/, int B,
                     void* stream);

/* SE(3) augmentation of a planar point cloud: perturb_se3 of the reference
 * (peract/voxel/augmentation.py:7-65) as one streaming kernel.  pcd, out [B,3,N] fp32 (one camera, N = H*W; out may alias pcd);
 * xform [B,15] = keyframe gripper position a(3), rot_shift[0:3,0:3] row-major (9), c(3) = clamp(a + trans_shift, bounds):
 * p' = (p - a) . R + c.  The sampling / rejection of the perturbation (augmentation.py:68-185) stays on the host
 * (voxactb_b200/augmentation.py), it touches a few floats per sample. */
int vxb_se3_perturb_f32(const float* pcd, const float* xform, float* out, int B, long long N, void* stream);

/* ------------------------------------------------------------------ building blocks (exported for the per-op parity tests) */
/* number of tcgen05 (split 16-bit x3) GEMM kernels launched so far by this process: lets tests prove that
 * VXB_MATH_F16X3 really ran on the tensor cores and did not fall back to the FFMA path */
long long vxb_umma_launch_count(void);

/* C[M,N] = act(alpha * A[M,K] * W[N,K]^T + bias[N]) (+ residual[(m % res_rows),N]); row-major fp32.
 * Replaces nn.Linear / DenseBlock (perceiver_lang_io.py:80-90,100-104,229-238,321-334; network_utils.py:257-289).
 * ws (vxb_linear_workspace_bytes) holds the 16-bit hi/lo operand planes of the tcgen05 path. */
size_t vxb_linear_workspace_bytes(int M, int N, int K);
int vxb_linear_f32(const float* A, int lda, const float* W, int ldw, const float* bias,
                   const float* residual, int res_rows, float* C, int ldc,
                   int M, int N, int K, float alpha, float act_slope /* <0: no activation */,
                   int math_mode, void* ws, size_t ws_bytes, void* stream);
/* rows of length n: y = (x-mean)/sqrt(var+1e-5)*w+b -- nn.LayerNorm of PreNorm (perceiver_lang_io.py:56-71) */
int vxb_layernorm_f32(const float* x, const float* w, const float* b, float* y, int rows, int n,
                      void* stream);
/* spatial soft-argmax (T=0.01) + max over P = Dd*Hh*Ww positions of channels-last x [B,P,C]:
 * ss [B,3C] ordered (c, [x,y,z]) with the reference's meshgrid axis convention
 * (network_utils.py:782-808), mx [B,C]. ws: vxb_spatial_softmax_workspace_bytes. */
size_t vxb_spatial_softmax_workspace_bytes(int B, int P, int C);
int vxb_spatial_softmax_f32(const float* x, int B, int Dd, int Hh, int Ww, int C, float* ss,
                            int ss_stride, float* mx, int mx_stride, void* ws, size_t ws_bytes,
                            void* stream);
/* channels-last conv3d, replicate padding k/2, stride s, weight in PyTorch layout [Co,Ci,k,k,k];
 * x [B,Di,Di,Di,Ci] -> y [B,Do,Do,Do,Co]; ws >= vxb_conv3d_workspace_bytes.
 * Replaces Conv3DBlock (network_utils.py:128-170) as used at perceiver_lang_io.py:217-226,302-311. */
size_t vxb_conv3d_workspace_bytes(int B, int Di, int Ci, int Co, int k);
int vxb_conv3d_f32(const float* x, const float* w, const float* bias, float* y, int B, int Di,
                   int Ci, int Co, int k, int s, float act_slope, int math_mode, void* ws,
                   size_t ws_bytes, void* stream);
/* fused conv(k,pad k/2,replicate) o trilinear-upsample(x s, align_corners=False): the second half of
 * Conv3DUpsampleBlock (network_utils.py:245-251) evaluated as s^3 polyphase 3x3x3 convolutions on
 * the low-resolution tensor.  x [B,S,S,S,Ci] -> y [B,S*s,S*s,S*s,Co]. */
size_t vxb_upconv3d_workspace_bytes(int B, int S, int Ci, int Co, int k, int s);
int vxb_upconv3d_f32(const float* x, const float* w, const float* bias, float* y, int B, int S,
                     int Ci, int Co, int k, int s, float act_slope, int math_mode, void* ws,
                     size_t ws_bytes, void* stream);
/* softmax(scale * Q K^T) V per (batch, head); q [B,Nq,H*dh] (ldq), k/v rows [B,Nk,*] (ldkv).
 * Replaces the einsum / softmax / einsum core of Attention.forward (perceiver_lang_io.py:107-132). */
size_t vxb_attention_workspace_bytes(int B, int H, int Nq, int Nk);
int vxb_attention_f32(const float* q, int ldq, long long q_batch_stride, const float* k,
                      const float* v, int ldkv, long long kv_batch_stride, float* out, int ldo,
                      long long o_batch_stride, int B, int H, int Nq, int Nk, int dh, float scale,
                      int math_mode, void* ws, size_t ws_bytes, void* stream);

/* ---- training tail building blocks (the Q-network backward is not built yet) ----
 * Per-sample cross entropy of logits [B, N] (row pitch ld) against label indices, replacing
 * nn.CrossEntropyLoss(reduction='none') on one-hot labels (reference qattention_peract_bc_agent.py:217,
 * 391-392, 517-578): loss[b] = logsumexp(logits[b]) - logits[b, labels[b]];
 * grad (optional, row pitch ldg) = grad_scale * (softmax(logits[b]) - onehot(labels[b])). */
size_t vxb_ce_loss_workspace_bytes(int B, int N);
int vxb_ce_loss_f32(const float* logits, long long ld, const int32_t* labels, int B, int N, float grad_scale,
                    float* loss, float* grad, long long ldg, void* ws, size_t ws_bytes, void* stream);

/* Fused multi-tensor optimizer steps.  params/grads/exp_avg/exp_avg_sq/sizes are HOST arrays of n_tensors
 * device pointers / element counts; state tensors are updated in place.
 * LAMB: reference peract/helpers/optim/lamb.py:60-122 (no bias correction, weight decay added to the
 *       step, trust ratio clamp(|w|,0,10)/|step|, 1 when either norm is 0).
 * Adam: torch.optim.Adam with L2 weight decay as the agent builds it (agent:263-268); step counts from 1. */
size_t vxb_optimizer_workspace_bytes(int n_tensors, const long long* sizes /* host */);
int vxb_lamb_step_f32(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, const long long* sizes, float lr, float beta1, float beta2,
                      float eps, float weight_decay, void* ws, size_t ws_bytes, void* stream);
int vxb_adam_step_f32(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                      float* const* exp_avg_sq, const long long* sizes, int step, float lr, float beta1,
                      float beta2, float eps, float weight_decay, void* ws, size_t ws_bytes, void* stream);

/* ---- gradient all-reduce of the training step over NCCL (NVLink 5 / NVSwitch), replacing the DistributedDataParallel wrap
 * of the reference (qattention_peract_bc_agent.py:50-54, process group from run_seed_fn.py:34).  The library resolves
 * libnccl.so.2 at run time (dlopen; the copy PyTorch has loaded) and owns its communicator:
 *   rank 0 calls vxb_nccl_unique_id and ships the 128 bytes to the other ranks (any side channel, e.g. a
 *   torch.distributed broadcast); every rank then calls vxb_nccl_init(id, rank, world, &comm).
 * vxb_allreduce_grads: flat[0:n] <- scale * sum over ranks (DDP: scale = 1 / world), issued as `bucket_elems`-sized
 * ncclAllReduce calls (0 = one call) inside one NCCL group on `stream`. */
int vxb_nccl_unique_id(char* id128 /* host, 128 bytes */);
int vxb_nccl_init(const char* id128, int rank, int world, void** comm /* out: opaque communicator */);
int vxb_nccl_destroy(void* comm);
int vxb_allreduce_grads(void* comm, float* flat, size_t n, float scale, size_t bucket_elems, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VOXACTB_H_ */

// Training-tail building blocks: the per-sample cross-entropy losses of
// QAttentionPerActBCAgent.update with their logit gradients, and fused multi-tensor LAMB / Adam steps.
//   loss    reference qattention_peract_bc_agent.py:217,391-392,517-578 (nn.CrossEntropyLoss(reduction='none')
//           on labels.argmax(-1): here the label INDEX is passed, no 10^6-wide one-hot is ever built)
//   LAMB    reference peract/helpers/optim/lamb.py:60-122 (no bias correction, L2 added to the step,
//           trust ratio clamp(|w|, 0, 10) / |step|)
//   Adam    torch.optim.Adam with L2 weight decay (agent:263-268)
// The backward pass of the Q-network itself is not part of this file (not built yet).
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace vxb {

// ------------------------------------------------------------------------------------------------ cross entropy
constexpr int CE_THREADS = 256;

// partial (max, sum exp(x - max)) per (chunk, row)
static __global__ void __launch_bounds__(CE_THREADS)
ce_partial_kernel(const float* __restrict__ x, long long ld, int N, int chunk, float* __restrict__ part) {
  const int b = blockIdx.y, ck = blockIdx.x, chunks = gridDim.x;
  const float* row = x + (long long)b * ld;
  const int beg = ck * chunk, end = min(N, beg + chunk);
  float m = -INFINITY, s = 0.f;
  for (int i = beg + threadIdx.x; i < end; i += CE_THREADS) {
    const float v = row[i];
    if (v > m) { s *= expf(m - v); m = v; }
    s += expf(v - m);
  }
  __shared__ float sm[CE_THREADS], ss[CE_THREADS];
  sm[threadIdx.x] = m; ss[threadIdx.x] = s;
  __syncthreads();
  for (int o = CE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float m2 = sm[threadIdx.x + o], s2 = ss[threadIdx.x + o];
      const float M = fmaxf(sm[threadIdx.x], m2);
      const float a = (sm[threadIdx.x] == -INFINITY) ? 0.f : ss[threadIdx.x] * expf(sm[threadIdx.x] - M);
      const float c = (m2 == -INFINITY) ? 0.f : s2 * expf(m2 - M);
      sm[threadIdx.x] = M; ss[threadIdx.x] = a + c;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[(b * chunks + ck) * 2] = sm[0]; part[(b * chunks + ck) * 2 + 1] = ss[0]; }
}

// lse[b] = log sum exp; loss[b] = lse - x[label]
static __global__ void ce_finish_kernel(const float* __restrict__ x, long long ld, const int32_t* __restrict__ labels,
                                        int B, int N, int chunks, const float* __restrict__ part,
                                        float* __restrict__ lse, float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float M = -INFINITY;
  for (int k = 0; k < chunks; ++k) M = fmaxf(M, part[(b * chunks + k) * 2]);
  float S = 0.f;
  for (int k = 0; k < chunks; ++k) {
    const float m = part[(b * chunks + k) * 2];
    if (m != -INFINITY) S += part[(b * chunks + k) * 2 + 1] * expf(m - M);
  }
  const float l = M + logf(S);
  lse[b] = l;
  // an out-of-range label poisons the sample (NaN loss and gradient) instead of being clamped silently:
  // nn.CrossEntropyLoss raises "Target out of bounds" there; the Python wrapper raises for host-side labels
  const int lab = labels[b];
  loss[b] = (lab >= 0 && lab < N) ? l - x[(long long)b * ld + lab] : __int_as_float(0x7fc00000);
}

// grad[b, i] = scale * (softmax(x)[b, i] - [i == label])
static __global__ void __launch_bounds__(CE_THREADS)
ce_grad_kernel(const float* __restrict__ x, long long ld, const int32_t* __restrict__ labels, int N,
               const float* __restrict__ lse, float scale, float* __restrict__ g, long long ldg) {
  const int b = blockIdx.y;
  const float l = lse[b];
  const int lab = labels[b];
  if (lab < 0 || lab >= N) scale = __int_as_float(0x7fc00000);
  for (int i = blockIdx.x * CE_THREADS + threadIdx.x; i < N; i += gridDim.x * CE_THREADS) {
    const float p = expf(x[(long long)b * ld + i] - l);
    g[(long long)b * ldg + i] = scale * (p - (i == lab ? 1.f : 0.f));
  }
}

static int ce_chunks(int N) { return std::max(1, std::min(256, N / 4096)); }

// ------------------------------------------------------------------------------------------------ optimizers
// Work is cut into blocks of OPT_CHUNK elements; blk_tensor / blk_begin map a block to its tensor slice.
constexpr int OPT_CHUNK = 4096;
constexpr int OPT_THREADS = 256;

struct OptTables {
  float** w;
  const float** g;
  float** m;
  float** v;
  const long long* n;
  const int* blk_tensor;
  const long long* blk_begin;
  double* norms;   // [tensors][2] = (sum w^2, sum step^2)
};

__device__ __forceinline__ double block_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = OPT_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  return sh[0];
}

// LAMB pass 1: moments in place, per-tensor sum w^2 and sum step^2   (lamb.py:92-108)
static __global__ void __launch_bounds__(OPT_THREADS)
lamb_moments_kernel(OptTables t, float beta1, float beta2, float eps, float wd) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long beg = t.blk_begin[blockIdx.x], end = min(t.n[ti], beg + OPT_CHUNK);
  float* w = t.w[ti]; const float* g = t.g[ti]; float* m = t.m[ti]; float* v = t.v[ti];
  double sw = 0.0, ss = 0.0;
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float gi = g[i];
    const float mi = m[i] * beta1 + gi * (1.f - beta1);
    const float vi = v[i] * beta2 + gi * gi * (1.f - beta2);
    m[i] = mi; v[i] = vi;
    const float wi = w[i];
    const float st = mi / (sqrtf(vi) + eps) + wd * wi;
    sw += (double)wi * wi;
    ss += (double)st * st;
  }
  __shared__ double sh[OPT_THREADS];
  const double a = block_sum(sw, sh);
  __syncthreads();
  const double c = block_sum(ss, sh);
  if (threadIdx.x == 0) { atomicAdd(&t.norms[2 * ti], a); atomicAdd(&t.norms[2 * ti + 1], c); }
}

// LAMB pass 2: w -= lr * trust * step, trust = clamp(|w|, 0, 10) / |step| (1 when either norm is 0)  (lamb.py:103-120)
static __global__ void __launch_bounds__(OPT_THREADS)
lamb_apply_kernel(OptTables t, float lr, float eps, float wd) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long beg = t.blk_begin[blockIdx.x], end = min(t.n[ti], beg + OPT_CHUNK);
  const float wn = fminf(fmaxf((float)sqrt(t.norms[2 * ti]), 0.f), 10.f);
  const float an = (float)sqrt(t.norms[2 * ti + 1]);
  const float trust = (wn == 0.f || an == 0.f) ? 1.f : wn / an;
  float* w = t.w[ti]; const float* m = t.m[ti]; const float* v = t.v[ti];
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float wi = w[i];
    const float st = m[i] / (sqrtf(v[i]) + eps) + wd * wi;
    w[i] = wi - lr * trust * st;
  }
}

// torch.optim.Adam (L2 weight decay folded into the gradient, bias-corrected)
static __global__ void __launch_bounds__(OPT_THREADS)
adam_kernel(OptTables t, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
  const int ti = t.blk_tensor[blockIdx.x];
  const long long beg = t.blk_begin[blockIdx.x], end = min(t.n[ti], beg + OPT_CHUNK);
  float* w = t.w[ti]; const float* g = t.g[ti]; float* m = t.m[ti]; float* v = t.v[ti];
  const float step_size = lr / bc1;
  for (long long i = beg + threadIdx.x; i < end; i += OPT_THREADS) {
    const float wi = w[i];
    const float gi = g[i] + wd * wi;
    const float mi = m[i] * beta1 + gi * (1.f - beta1);
    const float vi = v[i] * beta2 + gi * gi * (1.f - beta2);
    m[i] = mi; v[i] = vi;
    w[i] = wi - step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

struct OptLayout {
  size_t off_w, off_g, off_m, off_v, off_n, off_bt, off_bb, off_norms, total;
  int blocks;
};
static OptLayout opt_layout(int nt, const long long* sizes) {
  OptLayout L;
  long long blocks = 0;
  for (int i = 0; i < nt; ++i) blocks += (sizes[i] + OPT_CHUNK - 1) / OPT_CHUNK;
  L.blocks = (int)blocks;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return r; };
  L.off_w = take(sizeof(void*) * nt); L.off_g = take(sizeof(void*) * nt);
  L.off_m = take(sizeof(void*) * nt); L.off_v = take(sizeof(void*) * nt);
  L.off_n = take(sizeof(long long) * nt);
  L.off_bt = take(sizeof(int) * (size_t)blocks);
  L.off_bb = take(sizeof(long long) * (size_t)blocks);
  L.off_norms = take(sizeof(double) * 2 * nt);
  L.total = o;
  return L;
}

static int opt_upload(int nt, float* const* params, const float* const* grads, float* const* m, float* const* v,
                      const long long* sizes, void* ws, size_t ws_bytes, cudaStream_t st, OptTables& t, OptLayout& L) {
  VXB_CHECK_ARG(nt > 0 && params && grads && m && v && sizes && ws, "optimizer: bad arguments");
  L = opt_layout(nt, sizes);
  if (ws_bytes < L.total) {
    set_error("optimizer: workspace too small (%zu < %zu)", ws_bytes, L.total);
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  std::vector<int> bt((size_t)L.blocks);
  std::vector<long long> bb((size_t)L.blocks);
  size_t k = 0;
  for (int i = 0; i < nt; ++i)
    for (long long b = 0; b < sizes[i]; b += OPT_CHUNK) { bt[k] = i; bb[k] = b; ++k; }
  char* base = (char*)ws;
  // the staging vectors die at return: copy synchronously with respect to the host (small tables)
  VXB_CUDA(cudaMemcpyAsync(base + L.off_w, params, sizeof(void*) * nt, cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_g, grads, sizeof(void*) * nt, cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_m, m, sizeof(void*) * nt, cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_v, v, sizeof(void*) * nt, cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_n, sizes, sizeof(long long) * nt, cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_bt, bt.data(), sizeof(int) * bt.size(), cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemcpyAsync(base + L.off_bb, bb.data(), sizeof(long long) * bb.size(), cudaMemcpyHostToDevice, st));
  VXB_CUDA(cudaMemsetAsync(base + L.off_norms, 0, sizeof(double) * 2 * nt, st));
  VXB_CUDA(cudaStreamSynchronize(st));
  t.w = (float**)(base + L.off_w); t.g = (const float**)(base + L.off_g);
  t.m = (float**)(base + L.off_m); t.v = (float**)(base + L.off_v);
  t.n = (const long long*)(base + L.off_n);
  t.blk_tensor = (const int*)(base + L.off_bt); t.blk_begin = (const long long*)(base + L.off_bb);
  t.norms = (double*)(base + L.off_norms);
  return VXB_OK;
}

}  // namespace vxb

using namespace vxb;

extern "C" size_t vxb_ce_loss_workspace_bytes(int B, int N) {
  return align_up((size_t)B * ce_chunks(N) * 2 * sizeof(float), 256) + align_up((size_t)B * sizeof(float), 256);
}

extern "C" int vxb_ce_loss_f32(const float* logits, long long ld, const int32_t* labels, int B, int N,
                               float grad_scale, float* loss, float* grad, long long ldg, void* ws,
                               size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(logits && labels && loss && ws && B > 0 && N > 0 && ld >= N, "ce_loss: bad arguments");
  VXB_CHECK_ARG(!grad || ldg >= N, "ce_loss: bad gradient leading dimension");
  if (ws_bytes < vxb_ce_loss_workspace_bytes(B, N)) {
    set_error("ce_loss: workspace too small");
    return VXB_E_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = ce_chunks(N);
  const int chunk = (N + chunks - 1) / chunks;
  float* part = (float*)ws;
  float* lse = (float*)((char*)ws + align_up((size_t)B * chunks * 2 * sizeof(float), 256));
  ce_partial_kernel<<<dim3(chunks, B), CE_THREADS, 0, st>>>(logits, ld, N, chunk, part);
  VXB_LAUNCH_CHECK();
  ce_finish_kernel<<<cdiv(B, 64), 64, 0, st>>>(logits, ld, labels, B, N, chunks, part, lse, loss);
  VXB_LAUNCH_CHECK();
  if (grad) {
    const int gx = std::max(1, std::min(cdiv(N, CE_THREADS), 148 * 8 / std::max(1, B) + 1));
    ce_grad_kernel<<<dim3(gx, B), CE_THREADS, 0, st>>>(logits, ld, labels, N, lse, grad_scale, grad, ldg);
    VXB_LAUNCH_CHECK();
  }
  return VXB_OK;
}

extern "C" size_t vxb_optimizer_workspace_bytes(int n_tensors, const long long* sizes) {
  if (n_tensors <= 0 || !sizes) return 0;
  return opt_layout(n_tensors, sizes).total;
}

extern "C" int vxb_lamb_step_f32(int n_tensors, float* const* params, const float* const* grads,
                                 float* const* exp_avg, float* const* exp_avg_sq, const long long* sizes,
                                 float lr, float beta1, float beta2, float eps, float weight_decay,
                                 void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OptTables t;
  OptLayout L;
  VXB_TRY(opt_upload(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, ws, ws_bytes, st, t, L));
  lamb_moments_kernel<<<L.blocks, OPT_THREADS, 0, st>>>(t, beta1, beta2, eps, weight_decay);
  VXB_LAUNCH_CHECK();
  lamb_apply_kernel<<<L.blocks, OPT_THREADS, 0, st>>>(t, lr, eps, weight_decay);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

extern "C" int vxb_adam_step_f32(int n_tensors, float* const* params, const float* const* grads,
                                 float* const* exp_avg, float* const* exp_avg_sq, const long long* sizes,
                                 int step, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 void* ws, size_t ws_bytes, void* stream) {
  VXB_CHECK_ARG(step >= 1, "adam: step counts from 1");
  cudaStream_t st = (cudaStream_t)stream;
  OptTables t;
  OptLayout L;
  VXB_TRY(opt_upload(n_tensors, params, grads, exp_avg, exp_avg_sq, sizes, ws, ws_bytes, st, t, L));
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<L.blocks, OPT_THREADS, 0, st>>>(t, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt);
  VXB_LAUNCH_CHECK();
  return VXB_OK;
}

// ------------------------------------------------------------------------------------------------ NCCL gradient all-reduce
// DDP-equivalent gradient averaging (reference qattention_peract_bc_agent.py:50-54 wraps the Q-network in
// DistributedDataParallel; run_seed_fn.py:34 creates the process group): one all-reduce of the flat fp32 gradient arena
// per step over NVLink / NVSwitch.  libnccl is resolved at run time (dlopen of the copy the host process -- PyTorch --
// already loaded), so the library has no link-time NCCL dependency and CPU-only hosts can still load it.
#include <dlfcn.h>
#include <mutex>

struct Id128 { char bytes[128]; };   // ncclUniqueId (passed by value to ncclCommInitRank)
namespace {
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static bool nccl_load() {
  std::call_once(g_nccl_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return;
    g_nccl.handle = h;
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(h, "ncclAllReduce");
    g_nccl.GroupStart = (int (*)())dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  });
  return g_nccl.handle && g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce &&
         g_nccl.GroupStart && g_nccl.GroupEnd;
}
static int nccl_fail(const char* what, int rc) {
  set_error("%s failed: NCCL error %d (%s)", what, rc, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return VXB_E_NCCL;
}
#define VXB_NCCL(call, what)                    \
  do {                                          \
    const int _r = (call);                      \
    if (_r != 0) return nccl_fail(what, _r);    \
  } while (0)

extern "C" int vxb_nccl_unique_id(char* id128) {
  VXB_CHECK_ARG(id128, "nccl_unique_id: null pointer");
  if (!nccl_load()) {
    set_error("libnccl.so.2 could not be loaded");
    return VXB_E_NCCL;
  }
  VXB_NCCL(g_nccl.GetUniqueId(id128), "ncclGetUniqueId");
  return VXB_OK;
}

extern "C" int vxb_nccl_init(const char* id128, int rank, int world, void** comm) {
  VXB_CHECK_ARG(id128 && comm && world > 0 && rank >= 0 && rank < world, "nccl_init: bad arguments");
  if (!nccl_load()) {
    set_error("libnccl.so.2 could not be loaded");
    return VXB_E_NCCL;
  }
  Id128 id;
  memcpy(id.bytes, id128, 128);
  VXB_NCCL(g_nccl.CommInitRank(comm, world, id, rank), "ncclCommInitRank");
  return VXB_OK;
}

extern "C" int vxb_nccl_destroy(void* comm) {
  if (!comm) return VXB_OK;
  if (!nccl_load()) return VXB_E_NCCL;
  VXB_NCCL(g_nccl.CommDestroy(comm), "ncclCommDestroy");
  return VXB_OK;
}

static __global__ void scale_kernel(float* __restrict__ x, size_t n, float s) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}

extern "C" int vxb_allreduce_grads(void* comm, float* flat, size_t n, float scale, size_t bucket_elems, void* stream) {
  VXB_CHECK_ARG(comm && flat && n > 0, "allreduce_grads: bad arguments");
  if (!nccl_load()) {
    set_error("libnccl.so.2 could not be loaded");
    return VXB_E_NCCL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (bucket_elems == 0) bucket_elems = n;
  const int kFloat32 = 7, kSum = 0;
  VXB_NCCL(g_nccl.GroupStart(), "ncclGroupStart");
  for (size_t o = 0; o < n; o += bucket_elems) {
    const size_t cnt = std::min(bucket_elems, n - o);
    const int r = g_nccl.AllReduce(flat + o, flat + o, cnt, kFloat32, kSum, comm, st);
    if (r != 0) {
      g_nccl.GroupEnd();
      return nccl_fail("ncclAllReduce", r);
    }
  }
  VXB_NCCL(g_nccl.GroupEnd(), "ncclGroupEnd");
  if (scale != 1.f) {
    scale_kernel<<<148 * 8, 256, 0, st>>>(flat, n, scale);
    VXB_LAUNCH_CHECK();
  }
  return VXB_OK;
}

// Multi-tensor optimiser kernels for the training-step glue (row a19): Adam (torch.optim.Adam semantics, as
// train_only_inner_loop.py:573-575 / train.py:517 use it) and the Reptile-style outer update
// theta += eta * (omega - theta) of train.py:524-534, split into "delta = omega - theta into a flat buffer"
// (so the delta can be all-reduced over NCCL when every GPU holds a different style task) and
// "theta += scale * delta".  HBM-streaming, float4-vectorised, one launch for all parameter tensors.
#include "../../include/mst_b200.h"
#include "common.cuh"

namespace mst {

constexpr int OPT_CHUNK = 2048 * 4;  // elements per CTA-chunk (256 threads x 8 float4)

// device-side table: for chunk c, tensor index and offset inside the tensor are found from chunk_start[] (prefix sums)
MST_DEVINL int find_tensor(const int* __restrict__ chunk_start, int n_tensors, int chunk) {
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// dev_state (optional, CUDA-graph-capturable mode): {float lr, int step}; step has already been incremented for this update
struct AdamDevState {
  float lr;
  int step;
};
__global__ void adam_advance_kernel(AdamDevState* st) { st->step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(const MstTensorTable tb, float lr, float beta1, float beta2, float eps,
                                                   float weight_decay, float bc1, float bc2_sqrt, const AdamDevState* dev_state) {
  if (dev_state) {  // the captured graph must not bake the step count or the learning rate into its launch arguments
    lr = dev_state->lr;
    const double t = (double)dev_state->step;
    bc1 = (float)(1.0 - pow((double)beta1, t));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  const int t = find_tensor(tb.chunk_start, tb.n_tensors, blockIdx.x);
  const long long base = (long long)(blockIdx.x - tb.chunk_start[t]) * OPT_CHUNK;
  const long long n = tb.numel[t];
  float* __restrict__ p = reinterpret_cast<float*>(tb.a[t]);
  const float* __restrict__ g = reinterpret_cast<const float*>(tb.b[t]);
  float* __restrict__ m = reinterpret_cast<float*>(tb.c[t]);
  float* __restrict__ v = reinterpret_cast<float*>(tb.d[t]);
  const float step_size = lr / bc1;
  for (long long i = base + threadIdx.x * 4; i < base + OPT_CHUNK && i < n; i += 256 * 4) {
    if (i + 3 < n && (((uintptr_t)(p + i) | (uintptr_t)(g + i) | (uintptr_t)(m + i) | (uintptr_t)(v + i)) & 15) == 0) {
      float4 pp = *reinterpret_cast<float4*>(p + i), gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gr = ga[e] + weight_decay * pa[e];
        ma[e] = ma[e] + (1.0f - beta1) * (gr - ma[e]);        // torch: exp_avg.lerp_(grad, 1-beta1)
        va[e] = beta2 * va[e] + (1.0f - beta2) * gr * gr;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
        const float denom = sqrtf(va[e]) / bc2_sqrt + eps;
        pa[e] -= step_size * (ma[e] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long j = i; j < i + 4 && j < n; ++j) {
        const float gr = g[j] + weight_decay * p[j];
        const float mj = m[j] + (1.0f - beta1) * (gr - m[j]);
        const float vj = beta2 * v[j] + (1.0f - beta2) * gr * gr;
        m[j] = mj; v[j] = vj;
        p[j] -= step_size * (mj / (sqrtf(vj) / bc2_sqrt + eps));
      }
    }
  }
}

// mode 0: flat[off + i] = omega[i] - theta[i]          (a = theta, b = omega)
// mode 1: theta[i] += scale * flat[off + i]            (a = theta)
__global__ void __launch_bounds__(256) reptile_kernel(const MstTensorTable tb, float* __restrict__ flat, float scale, int mode) {
  const int t = find_tensor(tb.chunk_start, tb.n_tensors, blockIdx.x);
  const long long base = (long long)(blockIdx.x - tb.chunk_start[t]) * OPT_CHUNK;
  const long long n = tb.numel[t];
  float* __restrict__ theta = reinterpret_cast<float*>(tb.a[t]);
  const float* __restrict__ omega = reinterpret_cast<const float*>(tb.b[t]);
  float* __restrict__ f = flat + tb.flat_offset[t];
  for (long long i = base + threadIdx.x; i < base + OPT_CHUNK && i < n; i += 256) {
    if (mode == 0) f[i] = omega[i] - theta[i];
    else theta[i] += scale * f[i];
  }
}

}  // namespace mst

using namespace mst;

extern "C" int mst_opt_chunk_elems(void) { return OPT_CHUNK; }

extern "C" int mst_adam_step(const MstTensorTable* tb, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             void* stream) {
  if (!tb || tb->n_tensors <= 0 || tb->n_chunks <= 0 || !tb->chunk_start || !tb->numel || !tb->a || !tb->b || !tb->c || !tb->d || step < 1)
    return MST_ERR_BAD_ARG;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<tb->n_chunks, 256, 0, (cudaStream_t)stream>>>(*tb, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, nullptr);
  return (int)cudaGetLastError();
}

extern "C" int mst_adam_step_dev(const MstTensorTable* tb, float beta1, float beta2, float eps, float weight_decay, void* dev_state,
                                 int advance, void* stream) {
  if (!tb || tb->n_tensors <= 0 || tb->n_chunks <= 0 || !tb->chunk_start || !tb->numel || !tb->a || !tb->b || !tb->c || !tb->d || !dev_state)
    return MST_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (advance) adam_advance_kernel<<<1, 1, 0, st>>>(reinterpret_cast<AdamDevState*>(dev_state));
  adam_kernel<<<tb->n_chunks, 256, 0, st>>>(*tb, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, reinterpret_cast<const AdamDevState*>(dev_state));
  return (int)cudaGetLastError();
}

extern "C" int mst_reptile_delta(const MstTensorTable* tb, float* flat, void* stream) {
  if (!tb || !flat || tb->n_tensors <= 0 || tb->n_chunks <= 0 || !tb->chunk_start || !tb->numel || !tb->flat_offset || !tb->a || !tb->b)
    return MST_ERR_BAD_ARG;
  reptile_kernel<<<tb->n_chunks, 256, 0, (cudaStream_t)stream>>>(*tb, flat, 0.f, 0);
  return (int)cudaGetLastError();
}

extern "C" int mst_reptile_apply(const MstTensorTable* tb, float* flat, float scale, void* stream) {
  if (!tb || !flat || tb->n_tensors <= 0 || tb->n_chunks <= 0 || !tb->chunk_start || !tb->numel || !tb->flat_offset || !tb->a)
    return MST_ERR_BAD_ARG;
  reptile_kernel<<<tb->n_chunks, 256, 0, (cudaStream_t)stream>>>(*tb, flat, scale, 1);
  return (int)cudaGetLastError();
}

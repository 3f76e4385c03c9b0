// Multi-tensor SGD with momentum and weight decay in ONE launch (torch.optim.SGD as configured at
// tool/train.py:140,274-276 runs ~33 foreach kernels over the 161 parameter tensors):
//   g' = g + wd * w;   buf = first ? g' : momentum * buf + (1 - dampening) * g';   w -= lr * (nesterov ? g' + momentum*buf : buf)
// Hyper-parameters are per parameter GROUP (the trainer rewrites the 8 group learning rates every iteration,
// tool/train.py:299-304) and travel by value in the launch parameters; tensors are described by a device-resident item
// table (built once) plus a device array of gradient pointers (gradients are fresh tensors every step).
#include "host_common.h"

namespace sb {

constexpr int kSgdChunk = 4096;   // elements per block

__global__ void __launch_bounds__(256)
sgd_multi_kernel(const semseg_sgd_item* __restrict__ items, const unsigned long long* __restrict__ grads, int n_items,
                 const semseg_sgd_hyper h) {
  __shared__ int s_item;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;
    const int b = static_cast<int>(blockIdx.x);
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].chunk0 <= b) lo = mid; else hi = mid - 1;
    }
    s_item = lo;
  }
  __syncthreads();
  const semseg_sgd_item it = items[s_item];
  const float* __restrict__ g = reinterpret_cast<const float*>(grads[s_item]);
  if (g == nullptr) return;                      // parameter without a gradient this step (torch skips it too)
  const float lr = h.lr[it.group], mom = h.momentum[it.group], wd = h.weight_decay[it.group], damp = h.dampening[it.group];
  const long long base = static_cast<long long>(static_cast<int>(blockIdx.x) - it.chunk0) * kSgdChunk;
  const long long end = min(base + kSgdChunk, it.n);
  auto upd = [&](float w, float gg, float b) -> float2 {
    gg = fmaf(wd, w, gg);
    b = it.first ? gg : fmaf(mom, b, (1.f - damp) * gg);
    const float step = h.nesterov ? fmaf(mom, b, gg) : (mom != 0.f ? b : gg);
    return make_float2(w - lr * step, b);
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(it.w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(it.buf)) & 15) == 0;
  if (vec) {
    for (long long i = base + 4LL * threadIdx.x; i + 3 < end; i += 4LL * blockDim.x) {
      float4 w = *reinterpret_cast<const float4*>(it.w + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 b = it.first ? make_float4(0, 0, 0, 0) : *reinterpret_cast<const float4*>(it.buf + i);
      float2 r;
      r = upd(w.x, gg.x, b.x); w.x = r.x; b.x = r.y;
      r = upd(w.y, gg.y, b.y); w.y = r.x; b.y = r.y;
      r = upd(w.z, gg.z, b.z); w.z = r.x; b.z = r.y;
      r = upd(w.w, gg.w, b.w); w.w = r.x; b.w = r.y;
      *reinterpret_cast<float4*>(it.w + i) = w;
      *reinterpret_cast<float4*>(it.buf + i) = b;
    }
    const long long tail = base + ((end - base) & ~3LL);
    for (long long i = tail + threadIdx.x; i < end; i += blockDim.x) {
      const float2 r = upd(it.w[i], g[i], it.first ? 0.f : it.buf[i]);
      it.w[i] = r.x;
      it.buf[i] = r.y;
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
      const float2 r = upd(it.w[i], g[i], it.first ? 0.f : it.buf[i]);
      it.w[i] = r.x;
      it.buf[i] = r.y;
    }
  }
}

}  // namespace sb

extern "C" int semseg_sgd_chunk_elems(void) { return sb::kSgdChunk; }

extern "C" int semseg_sgd_multi(const semseg_sgd_item* items_dev, const void* grad_ptrs_dev, int n_items, int n_chunks,
                                const semseg_sgd_hyper* hyper, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SB_CHECK_ARG(items_dev && grad_ptrs_dev && hyper && n_items > 0 && n_chunks > 0, "sgd_multi: bad args");
  sb::sgd_multi_kernel<<<static_cast<unsigned>(n_chunks), 256, 0, stream>>>(
      items_dev, static_cast<const unsigned long long*>(grad_ptrs_dev), n_items, *hyper);
  SB_LAUNCHED();
  return SEMSEG_OK;
}

// hcj_internal.h — definitions shared by the translation units that implement the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <vector>

#include "../../include/hcjpeg.h"

#define CU_TRY(expr)                                      \
  do {                                                    \
    cudaError_t e_ = (expr);                              \
    if (e_ != cudaSuccess) return HCJ_ERR_CUDA - (int)e_; \
  } while (0)

namespace hcj {
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
struct FreeBlock {
  void *p;
  size_t cap;
};
}  // namespace hcj

struct hcj_ctx {
  int device = 0;
  int sm_count = 148;  // of `device` (hcjk::configure_device)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t enc0 = nullptr, enc1 = nullptr;  // around the device work of the latest hcj_encode_batch
  bool enc_timed = false;
  float enc_ms = 0.f;  // kernel time of the latest hcj_encode_batch, summed over its chunks
  cudaStream_t copy_stream = nullptr;  // D2H of finished chunks overlaps the kernels of the next chunk
  cudaStream_t up_stream = nullptr;    // H2D of the next chunk's files overlaps both
  std::vector<cudaEvent_t> up_events;
  void *hdr_buf = nullptr;  // parsed headers of the batch being created (39 KB each): kept between calls, never zero-filled
  size_t hdr_cap = 0;
  std::vector<cudaEvent_t> chunk_events;
  std::vector<hcj::FreeBlock> pool;  // device memory recycled between batches (grow-only)

  int alloc(void **p, size_t bytes) {
    bytes = hcj::align_up(std::max<size_t>(bytes, 256), 256);
    int best = -1;
    for (size_t i = 0; i < pool.size(); i++)
      if (pool[i].cap >= bytes && (best < 0 || pool[i].cap < pool[best].cap)) best = (int)i;
    if (best >= 0 && pool[best].cap <= bytes * 2 + (1 << 20)) {
      *p = pool[best].p;
      pool.erase(pool.begin() + best);
      return HCJ_OK;
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give cached blocks back and retry once
      for (auto &f : pool) cudaFree(f.p);
      pool.clear();
      (void)cudaGetLastError();
      e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) return e == cudaErrorMemoryAllocation ? HCJ_ERR_OUT_OF_MEMORY : HCJ_ERR_CUDA - (int)e;
    sizes[*p] = bytes;
    return HCJ_OK;
  }
  void release(void *p) {
    if (!p) return;
    auto it = sizes.find(p);
    pool.push_back({p, it == sizes.end() ? 0 : it->second});
  }
  std::map<void *, size_t> sizes;
};


// Host-side helpers of the set-up path: a thread-pool-free parallel loop and a
// pinned, double-buffered host -> device copy for the big operator arrays
// (a pageable cudaMemcpy of 13 GB runs at ~7 GB/s; staged through pinned
// buffers filled by several threads it follows the PCIe link).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <memory>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>
#include <cstdlib>
#ifdef __linux__
#include <sched.h>
#endif

namespace eqd {

// std::vector whose resize() leaves new elements uninitialised: the first touch of the
// big set-up arrays then happens inside the parallel copy loops (page faults spread
// over the threads) instead of in a serial zero fill
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U> struct rebind { using other = default_init_allocator<U>; };
  using std::allocator<T>::allocator;
  template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new ((void*)p) U; }
  template <class U, class... A> void construct(U* p, A&&... a) { ::new ((void*)p) U(std::forward<A>(a)...); }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

// worker threads of the set-up loops: EQD_HOST_THREADS when the host sets it (several ranks sharing the cores of one
// box should: cores / ranks), else the cores this process may run on, at most 16
inline int host_threads() {
  static const int nt = [] {
    if (const char* e = getenv("EQD_HOST_THREADS")) { const int v = atoi(e); if (v >= 1) return std::min(v, 64); }
    int n = (int)std::thread::hardware_concurrency();
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = std::min(n, (int)CPU_COUNT(&set));
#endif
    return std::max(1, std::min(n, 16));
  }();
  return nt;
}

// fn(begin, end) over [0, n) split into contiguous chunks, one per thread
template <class F>
void parallel_range(size_t n, F&& fn, size_t serialBelow = 1 << 14) {
  const int nt = host_threads();
  if (n < serialBelow || nt == 1) { fn((size_t)0, n); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const size_t b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([=, &fn] { fn(b, e); });
  }
  for (auto& x : th) x.join();
}

inline void parallel_memcpy(void* dst, const void* src, size_t bytes) {
  parallel_range(bytes, [&](size_t b, size_t e) { std::memcpy((char*)dst + b, (const char*)src + b, e - b); }, 1 << 22);
}

// Pinned double buffer, allocated once per process and reused by every upload.
class StagedCopy {
 public:
  static constexpr size_t kChunk = 128u << 20;
  ~StagedCopy() {
    for (int i = 0; i < 2; ++i) {
      if (buf_[i]) cudaFreeHost(buf_[i]);
      if (ev_[i]) cudaEventDestroy(ev_[i]);
    }
  }
  // blocking host -> device copy of `bytes` on stream s; returns a cudaError_t
  cudaError_t h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (bytes < (8u << 20)) {
      cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
      return e != cudaSuccess ? e : cudaStreamSynchronize(s);
    }
    cudaError_t e = init();
    if (e != cudaSuccess) return e;
    int k = 0;
    for (size_t off = 0; off < bytes; off += kChunk, k ^= 1) {
      const size_t n = std::min(kChunk, bytes - off);
      if ((e = cudaEventSynchronize(ev_[k])) != cudaSuccess) return e;   // the buffer's previous DMA has drained
      parallel_memcpy(buf_[k], (const char*)src + off, n);
      if ((e = cudaMemcpyAsync((char*)dst + off, buf_[k], n, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
      if ((e = cudaEventRecord(ev_[k], s)) != cudaSuccess) return e;
    }
    return cudaStreamSynchronize(s);
  }

  // One of the two pinned buffers (kChunk bytes), for staged device -> host reads.  The caller
  // must have drained the stream the uploads used (the buffers may still be DMA sources).
  cudaError_t pinned(int i, void** out) {
    cudaError_t e = init();
    if (e != cudaSuccess) return e;
    *out = buf_[i & 1];
    return cudaSuccess;
  }

 private:
  cudaError_t init() {
    for (int i = 0; i < 2; ++i) {
      if (!buf_[i]) {
        cudaError_t e = cudaHostAlloc(&buf_[i], kChunk, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming)) != cudaSuccess) return e;
      }
    }
    return cudaSuccess;
  }
  void* buf_[2] = {nullptr, nullptr};
  cudaEvent_t ev_[2] = {nullptr, nullptr};
};

}  // namespace eqd

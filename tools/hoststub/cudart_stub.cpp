// Host-only stand-in for the CUDA runtime -- TEST / PROFILING INFRASTRUCTURE, never shipped
// and never loaded by the product (eqdyna_b200/ does not reference it).
//
// The development container has no GPU.  Linking the step library's objects against this
// stub instead of libcudart gives a library whose *host-side* set-up code (eqd_set_mesh,
// tile planner, slot tables, staging, uploads: the part of bench.py's `e2e` that is not
// stepping) runs here: "device" memory is host memory, copies are memcpy, kernel launches
// do nothing.  tools/hoststub/setup_probe.py uses it to (1) time the set-up phases with
// EQD_VERBOSE=1 and (2) hash every buffer the host uploaded, so that a change to the
// set-up path can be checked for byte-identical uploads without a GPU.
// Numbers computed by kernels (operators, masses, steps) are NOT produced.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

extern "C" {

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
struct dim3 { unsigned x, y, z; };

static std::mutex g_mu;
static std::map<void*, size_t> g_allocs;   // live "device" buffers

cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
cudaError_t cudaSetDevice(int) { return 0; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return 0; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return 0; }
cudaError_t cudaDeviceSynchronize() { return 0; }
cudaError_t cudaGetLastError() { return 0; }
const char* cudaGetErrorString(cudaError_t) { return "stub"; }
// "Device" memory comes from one arena that is faulted in up front (EQD_STUB_ARENA_GB, default 6),
// so that an allocation costs what cudaMalloc costs -- nothing -- instead of page faults at the
// first copy.  Bump allocation; the arena is rewound when the last buffer is freed.
static char* g_arena = nullptr;
static size_t g_arenaSize = 0, g_arenaTop = 0;
static void arena_init() {
  if (g_arena) return;
  const char* e = std::getenv("EQD_STUB_ARENA_GB");
  g_arenaSize = (size_t)((e ? std::atof(e) : 6.0) * (1ull << 30));
  g_arena = (char*)std::malloc(g_arenaSize);
  if (g_arena) std::memset(g_arena, 0, g_arenaSize);
}
cudaError_t cudaMalloc(void** p, size_t n) {
  std::lock_guard<std::mutex> g(g_mu);
  arena_init();
  const size_t need = (n + 255) & ~(size_t)255;
  if (!g_arena || g_arenaTop + need > g_arenaSize) { *p = nullptr; return 2; }
  *p = g_arena + g_arenaTop;
  g_arenaTop += need ? need : 256;
  g_allocs[*p] = n;
  return 0;
}
cudaError_t cudaFree(void* p) {
  if (!p) return 0;
  std::lock_guard<std::mutex> g(g_mu);
  g_allocs.erase(p);
  if (g_allocs.empty()) g_arenaTop = 0;
  return 0;
}
void stub_prefault() { std::lock_guard<std::mutex> g(g_mu); arena_init(); }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, int) { std::memcpy(d, s, n); return 0; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
cudaError_t cudaMemcpyToSymbol(const void*, const void*, size_t, size_t, int) { return 0; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (void*)1; return 0; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)2; return 0; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (void*)1; return 0; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return 0; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return 0; }   // a made-up millisecond: callers divide by it
cudaError_t cudaFuncSetAttribute(const void*, int, int) { return 0; }
struct cudaIpcMemHandle_st { char reserved[64]; };
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_st*, void*) { return 1; }   // no peers in the stand-in
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_st, unsigned) { return 1; }
cudaError_t cudaIpcCloseMemHandle(void*) { return 0; }
cudaError_t cudaDeviceSetLimit(int, size_t) { return 0; }
cudaError_t cudaDeviceGetLimit(size_t* v, int) { *v = 0; return 0; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 3; return 0; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) { *n = 3; return 0; }
cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return 0; }   // kernels do not run
unsigned __cudaPushCallConfiguration(dim3, dim3, size_t, void*) { return 0; }
cudaError_t __cudaPopCallConfiguration(dim3*, dim3*, size_t*, void*) { return 0; }
void** __cudaRegisterFatBinary(void*) { static void* h; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char*, char*, const char*, int, void*, void*, void*, void*, int*) {}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}

// FNV-1a over every live "device" buffer, in order of size then content hash (addresses vary)
// -> a fingerprint of everything the host uploaded.  out[0] = buffers, out[1] = bytes, out[2] = hash.
void stub_fingerprint(uint64_t* out) {
  std::lock_guard<std::mutex> g(g_mu);
  std::multimap<size_t, uint64_t> hs;
  uint64_t bytes = 0;
  for (auto& kv : g_allocs) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* p = (const unsigned char*)kv.first;
    const uint64_t* q = (const uint64_t*)p;
    size_t n8 = kv.second / 8;
    for (size_t i = 0; i < n8; ++i) { h ^= q[i]; h *= 1099511628211ull; }
    for (size_t i = n8 * 8; i < kv.second; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    hs.insert({kv.second, h});
    bytes += kv.second;
  }
  // buffers of equal size are combined commutatively
  uint64_t tot = 1469598103934665603ull;
  size_t last = (size_t)-1; uint64_t acc = 0;
  for (auto& kv : hs) {
    if (kv.first != last) { tot ^= acc; tot *= 1099511628211ull; tot ^= kv.first; tot *= 1099511628211ull; acc = 0; last = kv.first; }
    acc += kv.second;
  }
  tot ^= acc; tot *= 1099511628211ull;
  out[0] = g_allocs.size(); out[1] = bytes; out[2] = tot;
}
// per-buffer listing: sizes[i], hashes[i] for i < cap; returns the number of live buffers
int stub_list(uint64_t* sizes, uint64_t* hashes, int cap) {
  std::lock_guard<std::mutex> g(g_mu);
  std::multimap<size_t, uint64_t> hs;
  for (auto& kv : g_allocs) {
    uint64_t h = 1469598103934665603ull;
    const uint64_t* q = (const uint64_t*)kv.first;
    for (size_t i = 0; i < kv.second / 8; ++i) { h ^= q[i]; h *= 1099511628211ull; }
    hs.insert({kv.second, h});
  }
  int k = 0;
  for (auto& kv : hs) { if (k < cap) { sizes[k] = kv.first; hashes[k] = kv.second; } ++k; }
  return k;
}

}  // extern "C"

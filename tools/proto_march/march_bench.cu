// PROTOTYPE microbenchmark (not part of the product): k_march_reg of march_kernel.cuh on a synthetic
// rectilinear block, stand-alone -- to be the FIRST thing run on a B200 in round 2, before any integration:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o march_bench march_bench.cu
//   ./march_bench --check                 # small block: GPU result against the host reading of the same source
//   ./march_bench 512 160 256 20 [ctas_per_sm] [variant 0|1|2]  # cells in x, y, z (multiples of 32, 16, 4), timed launches
//   ./march_bench --check 1 ; ./march_bench --check 2           # the same check for the other schedules
// variant 0: node planes loaded synchronously; 1: next-but-one plane prefetched through registers; 2: also the next
// element's operator values (3 CTAs per SM by launch bounds: pass ctas_per_sm = 3).
//
// The block is cut into complete 32 x 4 x 16 bundles; node ids are the structured ones (ix*ny*nz + iz*ny + iy),
// listed plane by plane as the kernel expects.  Operators and nodal fields are pseudo-random: the kernel's
// time does not depend on the values.  Prints the average launch time, elements per second and the bytes the
// kernel must move (15 operator / stress values + 6 stress writes per element, 48 B gathered and 24 B written
// per bundle node), against which `ncu` can be read.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "march_kernel.cuh"

using namespace march;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); std::exit(3); } \
  } while (0)

namespace {

double rnd(uint64_t& s) {   // xorshift, (0,1)
  s ^= s << 13; s ^= s >> 7; s ^= s << 17;
  return (double)(s >> 11) / 9007199254740992.0;
}

struct Problem {
  int ncx, ncy, ncz, LX = 32;
  std::vector<Bundle> rec;
  std::vector<int> tnode;
  size_t S = 0, Nn = 0, PFS = 0;
  std::vector<double> ax, ay, az, s0, s3, s5, lam, mu, det, stress, vel, disp;
};

void build(Problem& P) {
  const int nx = P.ncx + 1, ny = P.ncy + 1, nz = P.ncz + 1;
  P.Nn = (size_t)nx * ny * nz;
  for (int bx0 = 0; bx0 + P.LX <= P.ncx; bx0 += P.LX)
    for (int bz0 = 0; bz0 + BZ <= P.ncz; bz0 += BZ)
      for (int by0 = 0; by0 + BY <= P.ncy; by0 += BY) {
        P.rec.push_back(Bundle{(int)P.S, (int)P.tnode.size(), P.LX, 0});
        P.S += (size_t)P.LX * NCOL;
        for (int p = 0; p <= P.LX; ++p)
          for (int iz = 0; iz <= BZ; ++iz)
            for (int iy = 0; iy <= BY; ++iy) P.tnode.push_back(((bx0 + p) * nz + (bz0 + iz)) * ny + by0 + iy);
      }
  P.PFS = P.tnode.size();
  uint64_t seed = 88172645463325252ull;
  auto fill = [&](std::vector<double>& v, size_t n, double lo, double hi) {
    v.resize(n);
    for (double& x : v) x = lo + (hi - lo) * rnd(seed);
  };
  fill(P.ax, P.S, 2.0e-3, 3.0e-3); fill(P.ay, P.S, 2.0e-3, 3.0e-3); fill(P.az, P.S, 2.0e-3, 3.0e-3);
  fill(P.s0, P.S, 1.0e9, 2.0e9); fill(P.s3, P.S, 1.0e9, 2.0e9); fill(P.s5, P.S, 1.0e9, 2.0e9);
  fill(P.lam, P.S, 3.0e10, 3.3e10); fill(P.mu, P.S, 3.0e10, 3.3e10); fill(P.det, P.S, 1.0e5, 1.3e5);
  fill(P.stress, 6 * P.S, -1.0e6, 1.0e6);
  fill(P.vel, 3 * P.Nn, -1.0, 1.0); fill(P.disp, 3 * P.Nn, -1.0e-2, 1.0e-2);
}

template <class T>
T* to_device(const std::vector<T>& v) {
  T* p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)));
  CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return p;
}

Args device_args(const Problem& P, double** pfOut, double** stressOut) {
  Args A{};
  A.nBundles = (int)P.rec.size(); A.rec = to_device(P.rec); A.tnode = to_device(P.tnode);
  A.S = P.S; A.NnS = P.Nn; A.PFS = P.PFS;
  A.ax = to_device(P.ax); A.ay = to_device(P.ay); A.az = to_device(P.az);
  A.ss0 = to_device(P.s0); A.ss3 = to_device(P.s3); A.ss5 = to_device(P.s5);
  A.lam = to_device(P.lam); A.mu = to_device(P.mu); A.det = to_device(P.det);
  A.stress = to_device(P.stress); A.vel = to_device(P.vel); A.disp = to_device(P.disp);
  double* pf = nullptr;
  CK(cudaMalloc(&pf, 3 * P.PFS * sizeof(double)));
  CK(cudaMemset(pf, 0, 3 * P.PFS * sizeof(double)));
  A.pf = pf; *pfOut = pf; *stressOut = A.stress;
  A.dt = 0.008; A.rdampk = 0.1 * 0.008; A.w = 8.0;
  return A;
}

// the host reading of the same schedule (as tools/proto_march/march_proto.cpp does)
void host_run(const Problem& P, std::vector<double>& pf, std::vector<double>& stress) {
  pf.assign(3 * P.PFS, 0.0);
  stress = P.stress;
  Args A{};
  A.nBundles = (int)P.rec.size(); A.rec = P.rec.data(); A.tnode = P.tnode.data(); A.S = P.S; A.NnS = P.Nn; A.PFS = P.PFS;
  A.ax = P.ax.data(); A.ay = P.ay.data(); A.az = P.az.data(); A.ss0 = P.s0.data(); A.ss3 = P.s3.data(); A.ss5 = P.s5.data();
  A.lam = P.lam.data(); A.mu = P.mu.data(); A.det = P.det.data(); A.stress = stress.data();
  A.vel = P.vel.data(); A.disp = P.disp.data(); A.pf = pf.data(); A.dt = 0.008; A.rdampk = 0.1 * 0.008; A.w = 8.0;
  std::vector<Shared> smv(1);
  Shared& sm = smv[0];
  std::vector<Regs> R(NT);
  for (int b = 0; b < A.nBundles; ++b) {
    const Bundle B = A.rec[b];
#define MK_RUN(body) do { for (int tid = 0; tid < NT; ++tid) { body; } } while (0)
    MARCH_BUNDLE(MK_RUN, A, B, sm, R[tid]);
#undef MK_RUN
  }
}

}  // namespace

int main(int argc, char** argv) {
  const bool check = argc > 1 && !std::strcmp(argv[1], "--check");
  Problem P;
  int iters = 20, ctasPerSm = 4;
  int variant = 0;
  if (check) { P.ncx = 64; P.ncy = 32; P.ncz = 8; iters = 1; variant = argc > 2 ? std::atoi(argv[2]) : 0; }
  else {
    P.ncx = argc > 1 ? std::atoi(argv[1]) : 512; P.ncy = argc > 2 ? std::atoi(argv[2]) : 160; P.ncz = argc > 3 ? std::atoi(argv[3]) : 256;
    iters = argc > 4 ? std::atoi(argv[4]) : 20;
    ctasPerSm = argc > 5 ? std::atoi(argv[5]) : 4;
    variant = argc > 6 ? std::atoi(argv[6]) : 0;
  }
  if (P.ncx % 32 || P.ncy % BY || P.ncz % BZ || P.ncx <= 0) { std::fprintf(stderr, "cells must be multiples of 32, %d, %d\n", BY, BZ); return 4; }
  build(P);
  int dev = 0, sms = 148;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *dPf = nullptr, *dStress = nullptr;
  Args A = device_args(P, &dPf, &dStress);
  const int grid = std::min(A.nBundles, ctasPerSm * sms);
  std::printf("march_bench: %d x %d x %d cells = %zu elements in %d bundles, %zu bundle nodes (%.2f per element), grid %d x %d threads\n",
              P.ncx, P.ncy, P.ncz, P.S, A.nBundles, P.PFS, (double)P.PFS / P.S, grid, NT);
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  auto launch = [&] {
    if (variant == 2) k_march_reg_pf2<<<grid, NT, 0, s>>>(A);
    else if (variant == 1) k_march_reg_pf<<<grid, NT, 0, s>>>(A);
    else k_march_reg<<<grid, NT, 0, s>>>(A);
  };
  std::printf("schedule %d: %s\n", variant, variant == 2 ? "node planes and operator values prefetched through registers"
                                              : variant == 1 ? "node planes prefetched through registers" : "node planes loaded synchronously");
  launch();   // warm-up (also the checked launch)
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));
  if (check) {
    std::vector<double> pf(3 * P.PFS), st(6 * P.S), pfH, stH;
    CK(cudaMemcpy(pf.data(), dPf, pf.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), dStress, st.size() * sizeof(double), cudaMemcpyDeviceToHost));
    host_run(P, pfH, stH);
    double fmax = 0, fdev = 0, smax = 0, sdev = 0;
    for (size_t k = 0; k < pf.size(); ++k) { fmax = std::max(fmax, std::fabs(pfH[k])); fdev = std::max(fdev, std::fabs(pf[k] - pfH[k])); }
    for (size_t k = 0; k < st.size(); ++k) { smax = std::max(smax, std::fabs(stH[k])); sdev = std::max(sdev, std::fabs(st[k] - stH[k])); }
    std::printf("check: partial forces rel dev %.3e, stresses rel dev %.3e (FMA contraction on the device: expect ~1e-15)\n", fdev / fmax, sdev / smax);
    return (fdev / fmax < 1e-11 && sdev / smax < 1e-11) ? 0 : 1;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, s));
  for (int it = 0; it < iters; ++it) launch();
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  const double bytes = (double)P.S * (15 * 8 + 6 * 8) + (double)P.PFS * (4 + 48 + 24);
  std::printf("k_march_reg: %.3f ms per launch, %.3e elements/s, %.0f B per element to move => %.0f GB/s\n", ms, P.S / (ms * 1e-3),
              bytes / P.S, bytes / (ms * 1e-3) / 1e9);
  return 0;
}

// Launch wrappers of eqd_kernels.cu (host-callable).
#pragma once
#include <cuda_runtime.h>

#include "eqd_dev.cuh"

namespace eqd {
void upload_qtab(const QTab* t16);
void launch_advance(StepState* st, double dt, cudaStream_t s);
void launch_node_update(const NodeArgs& A, cudaStream_t s);
void launch_node_update_special(const NodeArgs& A, const int* list, int n, cudaStream_t s);
void launch_assemble_special(const NodeArgs& A, const int* list, int n, cudaStream_t s);
void launch_materialize_accel(const NodeArgs& A, double* out, cudaStream_t s);
// sweep tiles [A.tile0, A.tile0 + A.ntiles) of the class's launch order
void launch_elem_reg(const ElemArgs& A, bool split, bool plastic, bool q, bool body, int chg, cudaStream_t s);
void launch_elem_pml(const ElemArgs& A, bool body, int chg, cudaStream_t s);
size_t tile_smem_bytes(int cls, bool q, int LS);  // dynamic shared memory of the class's tile kernel
void launch_store_offfault(const int* idhist, int n, double* out, const double* vel, const double* disp, int NnS,
                           const StepState* st, cudaStream_t s);
void launch_sample_gm(const int* surf, int nSurf, const double* vel, int NnS, double* out, cudaStream_t s);
void launch_sample_src(const double* fric, int PS, int n, double* out, cudaStream_t s);
void launch_pack(const double* src, const uint32_t* idx, int n, double* buf, cudaStream_t s);
void launch_unpack_add(double* dst, const uint32_t* idx, int n, const double* buf, cudaStream_t s);
void launch_halo_send(const HaloAxisArgs& A, cudaStream_t s);
void launch_halo_recv(const HaloAxisArgs& A, cudaStream_t s);
void launch_thermop(const FaultArgs& A, cudaStream_t s);
void launch_fault(const FaultArgs& A, cudaStream_t s);
void launch_gather_rows(const double* src, int K, const int* refId, int S, double* dst, int row, cudaStream_t s);
void launch_aos_to_soa(const double* src, int K, int n, const int* dstIdx, int cls, double* dst, int S, int k0, int nk,
                       cudaStream_t s);
// eqd_ops.cu
void launch_elem_ops(const OpsArgs& A, cudaStream_t s);
void launch_tile_mass(const TileMassArgs& A, int ntiles, cudaStream_t s);
void launch_node_mass(const NodeMassArgs& A, cudaStream_t s);
}  // namespace eqd

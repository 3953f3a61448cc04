// nvcc translation unit of the prototype (compile check / resource usage only; see march_kernel.cuh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -c march_kernel.cu
#include "march_kernel.cuh"

namespace march {
void launch_march(const Args& A, int grid, int variant, cudaStream_t s) {
  if (variant == 2) k_march_reg_pf2<<<grid, NT, 0, s>>>(A);
  else if (variant == 1) k_march_reg_pf<<<grid, NT, 0, s>>>(A);
  else k_march_reg<<<grid, NT, 0, s>>>(A);
}
}  // namespace march

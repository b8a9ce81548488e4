// Geometry helpers of the alignment kernel (poa_kernels.cu).  `cols` = read columns per thread
// (8, or 16 with 256 threads, or 4 with 512); a pass covers threads*cols columns.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "poa_cell.h"
#include "poa_task.h"

namespace svs {

int poa_cols_per_thread(int threads, int cols);
size_t poa_dp_smem_bytes(int threads, int ring_rows, int cols);
int poa_dp_cols_per_pass(int threads, int cols);

}  // namespace svs

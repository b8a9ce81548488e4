// Launchers of the alignment kernels (poa_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "poa_cell.h"
#include "poa_task.h"

namespace svs {

size_t poa_dp_smem_bytes(int threads, int ring_rows);
int poa_dp_cols_per_pass(int threads);
cudaError_t poa_dp_configure(int threads, int ring_rows);
cudaError_t poa_dp_launch(const PoaTask* d_tasks, int n_tasks, const Scores& s, int threads,
                          int ring_rows, cudaStream_t stream);
cudaError_t poa_tb_launch(const PoaTask* d_tasks, int n_tasks, const Scores& s, cudaStream_t stream);

}  // namespace svs

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
// persistent scheduler kernel (512 threads, one CTA per SM, fused traceback)
cudaError_t poa_persistent_configure(int threads, int ring_rows);
cudaError_t poa_persistent_launch(const PoaTask* d_tasks, int n_tasks, int* d_counter, uint8_t* slot_base,
                                  uint64_t slot_bytes, int n_sm, const Scores& s, int ring_rows, cudaStream_t stream);
cudaError_t poa_tb_launch(const PoaTask* d_tasks, int n_tasks, const Scores& s, cudaStream_t stream);

}  // namespace svs

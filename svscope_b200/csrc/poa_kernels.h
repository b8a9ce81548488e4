// Launchers of the alignment kernels (poa_kernels.cu).  `cols` = read columns per thread
// (8, or 16 with 256 threads); a pass covers threads*cols columns.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "poa_cell.h"
#include "poa_task.h"

namespace svs {

int poa_cols_per_thread(int threads, int cols);
size_t poa_dp_smem_bytes(int threads, int ring_rows, int cols);
int poa_dp_cols_per_pass(int threads, int cols);
cudaError_t poa_dp_configure(int threads, int ring_rows, int cols);
cudaError_t poa_dp_launch(const PoaTask* d_tasks, int n_tasks, const Scores& s, int threads, int ring_rows, int cols,
                          cudaStream_t stream);
// persistent scheduler kernel (one CTA per SM, fused traceback): 512 threads x 8 or 256 x 16 columns
int poa_persistent_ctas_per_sm(int threads, int ring_rows, int cols);
bool poa_persistent_supported(int threads, int ring_rows, int cols);
cudaError_t poa_persistent_configure(int threads, int ring_rows, int cols);
cudaError_t poa_persistent_launch(const PoaTask* d_tasks, int n_tasks, int* d_counter, uint8_t* slot_base,
                                  uint64_t slot_bytes, int* slot_flags, int n_sm, const Scores& s, int threads,
                                  int ring_rows, int cols, cudaStream_t stream);
cudaError_t poa_tb_launch(const PoaTask* d_tasks, int n_tasks, const Scores& s, cudaStream_t stream);

}  // namespace svs

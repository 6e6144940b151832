// diag.cuh -- interface of the time-parallel kernels for decoupled (diagonal) models (diag.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace eks {
size_t diag_optimize_workspace_bytes(int dtype, int n_blocks, int B, int T);
int diag_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q, const void* C,
                  const void* y_base, long long y_seq_stride, const long long* y_off, const void* ymean,
                  const void* Rconst, int t_begin, int n, int n_blocks, const int* block_off, const int* members,
                  const void* s_log0, double lr, double lo, double hi, double tol, int cap, void* s_log_out,
                  void* last_loss_out, int* iters_out, void* trace, int trace_cap, void* workspace,
                  size_t workspace_bytes, cudaStream_t st);
// lag-statistics optimiser (diag_lag.cu): same contract, one pass over the observations + one persistent launch
size_t diag_lag_workspace_bytes(int dtype, int n_blocks, int B, int T);
int diag_lag_optimize(int dtype, int B, int T, const void* m0, const void* S0, const void* A, const void* Q,
                      const void* C, const void* y_base, long long y_seq_stride, const long long* y_off,
                      const void* ymean, const void* Rconst, int t_begin, int n, int n_blocks, const int* block_off,
                      const int* members, const void* s_log0, double lr, double lo, double hi, double tol, int cap,
                      void* s_log_out, void* last_loss_out, int* iters_out, void* trace, int trace_cap, void* workspace,
                      size_t workspace_bytes, cudaStream_t st);
}

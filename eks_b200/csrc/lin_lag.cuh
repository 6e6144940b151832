// lin_lag.cuh -- interface of the lag-statistics optimiser for linear models with A = I (lin_lag.cu).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
#include "generic.cuh"
namespace eks {
bool lin_lag_applicable(int dtype, int D, int O, int n_spans, int n);
size_t lin_lag_workspace_bytes(int dtype, int n_blocks, int B, int O, int T);
// Runs the whole optimisation when the closed form applies to every block (*used = 1); otherwise leaves the outputs to
// the caller's run-parallel path (*used = 0).  Synchronises the stream (reads the per-block flags).
template <class P>
int lin_lag_optimize(const GArgs<P>& a, void* workspace, size_t workspace_bytes, cudaStream_t st, int* used);
}  // namespace eks

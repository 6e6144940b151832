// diag.cuh -- interface of the time-parallel kernels for decoupled (diagonal) models (diag.cu).
#pragma once
#include <cstddef>
namespace eks {
size_t diag_optimize_workspace_bytes(int dtype, int n_blocks);
}

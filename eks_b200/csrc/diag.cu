#include "common.cuh"
#include "diag.cuh"
namespace eks {
size_t diag_optimize_workspace_bytes(int dtype, int n_blocks) { (void)dtype; (void)n_blocks; return 16; }
}

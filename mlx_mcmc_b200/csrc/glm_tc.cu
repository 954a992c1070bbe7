// tcgen05 + TMA implementation of the two GLM contractions (3xTF32).  Placeholder until the kernel lands:
// reports "not available" so the SIMT path is used.
#include "glm.cuh"

namespace b2m {
bool tc_available() { return false; }
int tc_gemm_resid(GlmModel &, int64_t, cudaStream_t) { set_error("tcgen05 path not built"); return 1; }
int tc_gemm_grad(GlmModel &, int64_t, cudaStream_t) { set_error("tcgen05 path not built"); return 1; }
}  // namespace b2m

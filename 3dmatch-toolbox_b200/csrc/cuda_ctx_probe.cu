// cuda_ctx_probe.cu -- creates a CUDA context on device 0 and exits: the fixed cost every CUDA process pays, measured
// by scripts/realdata_e2e.py next to the two demo binaries.
#include <cuda_runtime.h>
int main() { return cudaFree(0) == cudaSuccess ? 0 : 1; }

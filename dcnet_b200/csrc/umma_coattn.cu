// tcgen05 / TMEM / TMA co-attention forward.  Placeholder selector for the first build step: reports "unsupported"
// so dcnet_coattn_fwd takes the fp32 composition; replaced by the fused kernel below once it is parity-green.
#include "common.cuh"

bool umma_coattn_supported(int C, int N) { (void)C; (void)N; return false; }

int umma_coattn_fwd(const float*, const int*, const int*, const int*, int, float*, float*, int, int, float, void*, size_t,
                    cudaStream_t) {
  return dcnet_set_error(-2, "umma_coattn_fwd: not built");
}

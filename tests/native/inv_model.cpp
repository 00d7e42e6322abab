// Host build of csrc/inv.cuh (the SAME source the CUDA kernels compile) for the CPU test-suite (tests/test_inv_model.py).
#include "../../msm_blst_b200/csrc/inv.cuh"

extern "C" int inv_model_batch(const uint32_t *in, uint32_t *out, int n) {
    int max_rounds = 0;
    for (int k = 0; k < n; k++) {
        int r = msmb200::s30_inverse_words(out + 12 * k, in + 12 * k, [](bool done) { return done; });
        if (r > max_rounds) max_rounds = r;
    }
    return max_rounds;
}

// G1 instantiation of the MSM engine (F = Fp). Also hosts the Fp / Fp2 field-op test kernels.
#include "engine_impl.cuh"

namespace msmb200 {

// BLS12_381_G1 generator, affine Montgomery (reference src/e1.c:20-32)
static const uint64_t kG1[12] = {
    0x5cb38790fd530c16ULL, 0x7817fc679976fff5ULL, 0x154f95c7143ba1c1ULL, 0xf0ae6acdf3d0e747ULL, 0xedce6ecc21dbf440ULL, 0x120177419e0bfb75ULL,
    0xbaac93d50ce72271ULL, 0x8c22631a7918fd8eULL, 0xdd595f13570725ceULL, 0x51ac582950405194ULL, 0x0e1c8c3fad0059c0ULL, 0x0bbc3efc5008a26aULL};
template <> const uint32_t *generator_words<fpc_t>() { return (const uint32_t *)kG1; }

static int field_op_g(int field, int op, const void *a, const void *b, void *out, size_t n) {
    if (field == 1) field_op_kernel<fp_t><<<blocks_for(n, 128), 128>>>(op, (const fp_t *)a, (const fp_t *)b, (fp_t *)out, n);
    else field_op_kernel<fp2_t><<<blocks_for(n, 128), 128>>>(op, (const fp2_t *)a, (const fp2_t *)b, (fp2_t *)out, n);
    return cudaGetLastError() == cudaSuccess ? 0 : MSMB200_ECUDA;
}

// Register-only microbenchmarks (the kernels of microbench.cu) that give bench.py its integer roofline denominator live.
// ptxas folds `acc += x * y` with loop-invariant x, y into one multiplication plus adds (the round-1 figure of
// 61.5 MAC/clk/SM was that artefact), so both multiplicands here come from accumulators that change every iteration:
// one IMAD.WIDE.U32 per multiply-accumulate, 8 independent chains per thread, 32 warps per SM. The kernel runs for tens
// of milliseconds and reports its own cycle count, so the result is MACs per clock per SM plus the clock it ran at.
static __global__ void peak_imad_wide_kernel(uint64_t *out, uint32_t a, int iters, long long *cyc) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = ((uint64_t)(threadIdx.x * 2654435761u + k) << 32) | (a * (k + 1) + threadIdx.x);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"((uint32_t)acc[(k + 1) & 7]), "r"((uint32_t)(acc[(k + 3) & 7] >> 32)));
    }
    const long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
static __global__ void peak_fp_mul_kernel(fp_t *out, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    fp_t x, y;
    fp_set_one(x);
    fp_set_one(y);
    x.l[0] ^= tid; y.l[1] ^= tid * 2654435761u;
#pragma unroll 1
    for (int i = 0; i < iters; i++) { fp_mul(x, x, y); fp_mul(y, y, x); }
    x.l[0] ^= y.l[0];
    out[tid] = x;
}
// out[0] = IMAD.WIDE multiply-accumulates per second, out[1] = dependent fp_mul per second (the out-of-line multiplier the
// kernels call), out[2] = multiply-accumulates per clock per SM, out[3] = SM clock (MHz) during the IMAD kernel
int measure_peaks(double out[4]) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return MSMB200_ECUDA;
    void *buf = nullptr;
    long long *d_cyc = nullptr;
    if (cudaMalloc(&buf, (size_t)sms * 8 * 256 * 48) != cudaSuccess || cudaMalloc((void **)&d_cyc, (size_t)sms * 4 * sizeof(long long)) != cudaSuccess) return MSMB200_ECUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int it1 = 300000, it2 = 256;   // ~55 ms of IMAD.WIDE at 32 MAC/clk/SM
    float ms1 = 0, best2 = 1e30f, ms;
    peak_imad_wide_kernel<<<sms * 4, 256>>>((uint64_t *)buf, 3u, 1000, d_cyc);   // warm-up
    cudaEventRecord(e0);
    peak_imad_wide_kernel<<<sms * 4, 256>>>((uint64_t *)buf, 5u, it1, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms1, e0, e1);
    std::vector<long long> cyc((size_t)sms * 4);
    cudaMemcpy(cyc.data(), d_cyc, cyc.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg_cyc = 0;
    for (long long v : cyc) avg_cyc += (double)v;
    avg_cyc /= (double)cyc.size();
    for (int r = 0; r < 7; r++) {
        cudaEventRecord(e0);
        peak_fp_mul_kernel<<<sms * 8, 128>>>((fp_t *)buf, it2);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2 && ms < best2) best2 = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(d_cyc);
    if (cudaGetLastError() != cudaSuccess) return MSMB200_ECUDA;
    out[0] = 8.0 * it1 * (double)sms * 4 * 256 / (ms1 * 1e-3);
    out[1] = 2.0 * it2 * (double)sms * 8 * 128 / (best2 * 1e-3);
    out[2] = 8.0 * it1 * 4 * 256 / avg_cyc;      // 4 blocks of 256 threads per SM, all co-resident
    out[3] = avg_cyc / (ms1 * 1e3);
    return MSMB200_OK;
}

#ifdef MSMB200_BA_TIMING
extern "C" int msmb200_debug_ba_timing(unsigned long long *out) {
    return cudaMemcpyFromSymbol(out, ba_dbg, sizeof(ba_dbg)) == cudaSuccess ? 0 : -1;
}
#endif

static const GroupOps kOps = {
    sizeof(aff_t<fp_t>), sizeof(jac_t<fp_t>), sizeof(xyzz_t<fp_t>),
    msm_impl<fp_t, fpc_t>, generate_fix_points_impl<fpc_t>, table_build_impl<fpc_t>, sum_partials_impl<fpc_t>, combine_bits_impl<fpc_t>, tile_impl<fp_t, fpc_t>,
    pippenger_impl<fp_t, fpc_t>, shim_aux_impl, field_op_g, point_op_impl<fp_t, fpc_t>, digits_impl<fp_t>, resident_blocks_impl<fp_t, fpc_t>, wbits_precompute_impl<fpc_t>, table_io_impl<fpc_t>, checksum_impl};
const GroupOps *group_ops_g1() { return &kOps; }

}  // namespace msmb200

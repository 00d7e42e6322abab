// Register-only microbenchmarks that define the integer-multiply roofline used by bench.py
// ("IMAD peak must be measured on the box"): 32x32->64-bit multiply-accumulates per second in the
// two encodings a field kernel can use, plus the sustained throughput of fp_mul / xyzz_add_affine themselves.
// Standalone binary; prints one JSON object.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "ec.cuh"

using namespace msmb200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;

// A: mad.wide.u32 (IMAD.WIDE.U32), 8 independent 64-bit accumulators
__global__ void k_imad_wide(uint64_t *out, uint32_t a, uint32_t b, long long *cyc) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x), "r"(y));
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// B: carry-chained pairs mad.lo.cc/madc.hi.cc (IMAD.WIDE.U32.X) exactly as the field multiplier issues them
__global__ void k_imad_chain(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
    uint32_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
        acc[0] = mad_lo_cc(x, y, acc[0]);
        acc[1] = madc_hi_cc(x, y, acc[1]);
#pragma unroll
        for (int k = 2; k < 16; k += 2) {
            acc[k] = madc_lo_cc(x, y, acc[k]);
            acc[k + 1] = madc_hi_cc(x, y, acc[k + 1]);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// C: 32-bit mad.lo (IMAD) + mad.hi (IMAD.HI) as separate instructions: 2 issue slots per full MAC
__global__ void k_imad_lohi(uint32_t *out, uint32_t a, uint32_t b, long long *cyc) {
    uint32_t acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x + k;
    uint32_t x = a + threadIdx.x, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(x), "r"(y));
            asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[k + 1]) : "r"(x), "r"(y));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// D: dependent chain of fp_mul (x = x*y), the way a point addition uses it
__global__ void k_fp_mul(fp_t *out, const fp_t *in, int iters, long long *cyc) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    fp_t x = in[tid & 1023], y = in[(tid + 7) & 1023];
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) { fp_mul_inline(x, x, y); fp_mul_inline(y, y, x); }
    long long t1 = clock64();
    out[tid] = x;
    out[tid].l[0] ^= y.l[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// E: xyzz_add_affine chain (G1): acc += P repeatedly with varying P
__global__ void k_madd_g1(xyzz_t<fp_t> *out, const aff_t<fp_t> *pts, int iters, long long *cyc) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    xyzz_t<fp_t> acc;
    xyzz_set_inf(acc);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        aff_t<fp_t> p = pts[(tid * 7 + i) & 1023];
        xyzz_add_affine(acc, p, (i & 1) != 0);
    }
    long long t1 = clock64();
    out[tid] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class K, class... Args>
static void run(const char *name, double ops_per_thread, int blocks, int threads, K kern, long long *d_cyc, bool last, Args... args) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) kern<<<blocks, threads>>>(args..., d_cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        kern<<<blocks, threads>>>(args..., d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (auto c : cyc) avg += (double)c; avg /= blocks;
    double total = ops_per_thread * (double)blocks * threads;
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
    double resident_blocks_per_sm = (double)blocks / pr.multiProcessorCount;
    double per_clk_sm = ops_per_thread * threads * resident_blocks_per_sm / avg;  // valid when all blocks co-resident
    printf("  \"%s\": {\"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"Gops_s\": %.1f, \"ops_per_clk_per_sm\": %.2f, \"eff_mhz\": %.0f}%s\n",
           name, blocks, threads, best, total / best / 1e6, per_clk_sm, avg / best / 1e3, last ? "" : ",");
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int sms = pr.multiProcessorCount;
    long long *d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 64));
    void *d_out; CK(cudaMalloc(&d_out, (size_t)sms * 64 * 1024 * 192));
    // field inputs: arbitrary residues (Montgomery form of something) — take multiples of ONE via adds on host? use raw
    std::vector<fp_t> h(1024);
    uint64_t s = 88172645463325252ull;
    for (auto &f : h) { for (int i = 0; i < 12; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; f.l[i] = (uint32_t)s; } f.l[11] &= 0x0fffffff; }
    fp_t *d_in; CK(cudaMalloc(&d_in, sizeof(fp_t) * 2048));
    CK(cudaMemcpy(d_in, h.data(), sizeof(fp_t) * 1024, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_in + 1024, h.data(), sizeof(fp_t) * 1024, cudaMemcpyHostToDevice));
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", pr.name, sms, pr.clockRate);
    // MAC-peak kernels: 4 blocks/SM x 256 threads = 32 warps/SM
    run("imad_wide_u32", 8.0 * ITERS, sms * 4, 256, k_imad_wide, d_cyc, false, (uint64_t *)d_out, 3u, 5u);
    run("imad_wide_x_chain", 8.0 * ITERS, sms * 4, 256, k_imad_chain, d_cyc, false, (uint32_t *)d_out, 3u, 5u);
    run("imad_lo_hi_pair", 8.0 * ITERS, sms * 4, 256, k_imad_lohi, d_cyc, false, (uint32_t *)d_out, 3u, 5u);
    run("imad_wide_u32_8w", 8.0 * ITERS, sms * 1, 256, k_imad_wide, d_cyc, false, (uint64_t *)d_out, 3u, 5u);
    // fp_mul: 2 muls per iteration; vary warps per SM
    const int it = 512;
    run("fp_mul_4w", 2.0 * it, sms * 1, 128, k_fp_mul, d_cyc, false, (fp_t *)d_out, (const fp_t *)d_in, it);
    run("fp_mul_8w", 2.0 * it, sms * 2, 128, k_fp_mul, d_cyc, false, (fp_t *)d_out, (const fp_t *)d_in, it);
    run("fp_mul_16w", 2.0 * it, sms * 4, 128, k_fp_mul, d_cyc, false, (fp_t *)d_out, (const fp_t *)d_in, it);
    run("fp_mul_32w", 2.0 * it, sms * 8, 128, k_fp_mul, d_cyc, false, (fp_t *)d_out, (const fp_t *)d_in, it);
    // mixed add: 1 add per iteration (10 Fp-mul-equivalents)
    run("xyzz_madd_g1_8w", 1.0 * it, sms * 2, 128, k_madd_g1, d_cyc, false, (xyzz_t<fp_t> *)d_out, (const aff_t<fp_t> *)d_in, it);
    run("xyzz_madd_g1_16w", 1.0 * it, sms * 4, 128, k_madd_g1, d_cyc, true, (xyzz_t<fp_t> *)d_out, (const aff_t<fp_t> *)d_in, it);
    printf("}\n");
    return 0;
}

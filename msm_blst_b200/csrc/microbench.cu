// Register-only microbenchmarks that define the integer-multiply roofline used by bench.py
// ("IMAD peak must be measured on the box"): 32x32->64-bit multiply-accumulates per second in the
// two encodings a field kernel can use, plus the sustained throughput of fp_mul / xyzz_add_affine themselves.
// Standalone binary; prints one JSON object.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "ec.cuh"

using namespace msmb200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)


// How the variants differ (SASS of each checked with cuobjdump; ptxas folds `acc += x * y` with loop-invariant x, y into
// ONE multiplication plus 64-bit adds — the round-1 "peak" kernel measured IADD3, not IMAD.WIDE — so every variant here
// takes at least one multiplicand from a value that changes every iteration):
//   A  c += lo(c) * y          y shared by all instructions (operand-reuse cache can serve it)
//   B  c += lo(c) * hi(c')     both multiplicands change, one is the accumulator's own low word
//   C  c += lo(c') * hi(c'')   4 distinct source registers per instruction (what a field multiplier issues)
//   I  c += lo(c) * imm        multiplicand is an immediate (the m * p rows of the Montgomery reduction)
//   L  32-bit IMAD (lo only)   a = a * y + z
#define WIDE(acc, x, y) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(y))
__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }
template <int VARIANT> __global__ void k_imad_wide(uint64_t *out, uint32_t a, uint32_t b, int iters, long long *cyc) {
    uint64_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = ((uint64_t)(threadIdx.x * 2654435761u + k) << 32) | (a * (k + 1) + threadIdx.x);
    const uint32_t y = b | 1u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (VARIANT == 0) WIDE(acc[k], lo32(acc[k]), y);
            else if (VARIANT == 1) WIDE(acc[k], lo32(acc[k]), hi32(acc[(k + 1) & 7]));
            else if (VARIANT == 2) WIDE(acc[k], lo32(acc[(k + 1) & 7]), hi32(acc[(k + 3) & 7]));
            else asm volatile("mad.wide.u32 %0, %1, 0x1eabfffe, %0;" : "+l"(acc[k]) : "r"(lo32(acc[k])));
        }
    }
    long long t1 = clock64();
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_imad_lo(uint32_t *out, uint32_t a, uint32_t b, int iters, long long *cyc) {
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = a * (k + 1) + threadIdx.x;
    const uint32_t y = b | 1u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(y), "r"(acc[(k + 1) & 7]));
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// P: pure 32x32->64 products (IMAD.WIDE.U32 with RZ addend) consumed by ONE 3-input LOP3 each; H: IMAD.HI.U32 only;
// U: ALU-pipe reference, 3-input integer adds (the two adds per element fuse into one IADD3)
template <int VARIANT> __global__ void k_prod(uint32_t *out, uint32_t a, uint32_t b, int iters, long long *cyc) {
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a * (2 * k + 1) + threadIdx.x * 2654435761u;
    const uint32_t y = b | 1u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (VARIANT == 0) {
                uint64_t t;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[k]), "r"(x[(k + 3) & 7]));
                x[k] = lo32(t) ^ hi32(t) ^ y;
            } else if (VARIANT == 1) {
                uint32_t t;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(x[k]), "r"(x[(k + 3) & 7]));
                x[k] = t ^ x[(k + 1) & 7] ^ y;
            } else {
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 1) & 7]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(y));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// X: carry-chained IMAD.WIDE.U32.X rows exactly as the field multiplier issues them (two interleaved chains of 6, the
// multiplicands a[k] distinct, b_i shared by the row and replaced every row)
__global__ void k_imad_chain(uint32_t *out, uint32_t a, uint32_t b, int iters, long long *cyc) {
    uint32_t ev[12], od[12], x[12];
#pragma unroll
    for (int k = 0; k < 12; k++) { ev[k] = threadIdx.x + k; od[k] = a + k; x[k] = a * (2 * k + 1) + threadIdx.x; }
    uint32_t y = b | 1u;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        od[0] = mad_lo_cc(x[1], y, od[0]);
        od[1] = madc_hi_cc(x[1], y, od[1]);
#pragma unroll
        for (int k = 2; k < 12; k += 2) {
            od[k] = madc_lo_cc(x[k + 1], y, od[k]);
            od[k + 1] = madc_hi_cc(x[k + 1], y, od[k + 1]);
        }
        ev[0] = mad_lo_cc(x[0], y, ev[0]);
        ev[1] = madc_hi_cc(x[0], y, ev[1]);
#pragma unroll
        for (int k = 2; k < 12; k += 2) {
            ev[k] = madc_lo_cc(x[k], y, ev[k]);
            ev[k + 1] = madc_hi_cc(x[k], y, ev[k + 1]);
        }
        y = ev[0] ^ od[1];  // next row's b_i depends on this row: nothing is loop-invariant
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) s ^= ev[k] ^ od[k];
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


// ---- random table gather (the access pattern of the bucket accumulation): every thread reads BYTES bytes (16-byte vector
// loads) at offset OFF inside an entry of STRIDE bytes chosen by a hash of (thread, iteration) from a table of `entries`
// entries; UNROLL independent entries in flight per thread. Reports useful GB/s; run under
// `ncu --metrics dram__bytes_read.sum` to see what the memory system really fetched per entry.
template <int BYTES, int UNROLL>
__global__ void k_gather(const uint4 *__restrict__ table, size_t entries, int stride16, int off16, int iters, uint4 *out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    uint64_t h = (uint64_t)tid * 0x9E3779B97F4A7C15ull + 12345;
#pragma unroll 1
    for (int i = 0; i < iters; i += UNROLL) {
        uint4 v[UNROLL][BYTES / 16];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            h = h * 6364136223846793005ull + 1442695040888963407ull;
            const size_t e = (size_t)((h >> 20) % entries);
            const uint4 *p = table + e * (size_t)stride16 + off16;
#pragma unroll
            for (int k = 0; k < BYTES / 16; k++) v[u][k] = __ldg(p + k);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int k = 0; k < BYTES / 16; k++) { acc.x ^= v[u][k].x; acc.y += v[u][k].y; acc.z ^= v[u][k].z; acc.w += v[u][k].w; }
    }
    out[tid] = acc;
}
template <int BYTES, int UNROLL>
static void run_gather(const char *name, const uint4 *table, size_t table_bytes, int stride, int off, uint4 *out, int sms, bool last) {
    const size_t entries = table_bytes / stride;
    const int blocks = sms * 8, threads = 256, iters = 512;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_gather<BYTES, UNROLL><<<blocks, threads>>>(table, entries, stride / 16, off / 16, iters, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0));
        k_gather<BYTES, UNROLL><<<blocks, threads>>>(table, entries, stride / 16, off / 16, iters, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double n = (double)blocks * threads * iters;
    printf("  \"%s\": {\"bytes_per_entry\": %d, \"stride\": %d, \"offset\": %d, \"in_flight_per_thread\": %d, \"ms\": %.4f, \"M_entries_per_s\": %.1f, \"useful_GB_s\": %.1f}%s\n",
           name, BYTES, stride, off, UNROLL, best, n / best / 1e3, n * BYTES / best / 1e6, last ? "" : ",");
}
static int gather_main(int sms) {
    const size_t bytes = (size_t)8 << 30;   // 8 GB: far beyond the 126 MB L2, like the 7.25 GB CHES table
    uint4 *table, *out;
    CK(cudaMalloc(&table, bytes));
    CK(cudaMemset(table, 1, bytes));
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 16));
    printf("{\n");
    run_gather<96, 2>("entry96_stride96_full", table, bytes, 96, 0, out, sms, false);        // the shipped table: {x, y} at 96 k
    run_gather<48, 4>("entry96_stride96_x_only", table, bytes, 96, 0, out, sms, false);      // forward pass: x only
    run_gather<96, 2>("entry96_stride128_full", table, bytes, 128, 0, out, sms, false);      // entries padded to one 128-byte line
    run_gather<48, 4>("entry96_stride128_x_only", table, bytes, 128, 0, out, sms, false);
    run_gather<48, 4>("entry48_stride64_aligned", table, bytes, 64, 0, out, sms, false);     // separate x table, 64-byte slots
    run_gather<32, 4>("sector32_stride32", table, bytes, 32, 0, out, sms, false);
    run_gather<64, 4>("line64_stride64", table, bytes, 64, 0, out, sms, false);
    run_gather<128, 2>("line128_stride128", table, bytes, 128, 0, out, sms, false);
    run_gather<192, 1>("g2_entry192_stride192", table, bytes, 192, 0, out, sms, false);
    run_gather<192, 1>("g2_entry192_stride256", table, bytes, 256, 0, out, sms, true);
    printf("}\n");
    return 0;
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
    int sms = pr.multiProcessorCount;
    if (argc > 1 && std::string(argv[1]) == "gather") return gather_main(sms);
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
    // MAC-peak kernels: 4 blocks/SM x 256 threads = 32 warps/SM, >= 50 ms each so that the SM clock has settled; the
    // per-clock figure comes from clock64 inside the kernel, eff_mhz = cycles / wall time is the clock it ran at
    const int big = 400000;
    run("imad_wide_A_shared_multiplicand", 8.0 * big, sms * 4, 256, k_imad_wide<0>, d_cyc, false, (uint64_t *)d_out, 3u, 5u, big);
    run("imad_wide_B_two_changing", 8.0 * big, sms * 4, 256, k_imad_wide<1>, d_cyc, false, (uint64_t *)d_out, 3u, 5u, big);
    run("imad_wide_C_four_distinct_regs", 8.0 * big, sms * 4, 256, k_imad_wide<2>, d_cyc, false, (uint64_t *)d_out, 3u, 5u, big);
    run("imad_wide_I_immediate", 8.0 * big, sms * 4, 256, k_imad_wide<3>, d_cyc, false, (uint64_t *)d_out, 3u, 5u, big);
    run("imad_wide_C_8w", 8.0 * big, sms * 1, 256, k_imad_wide<2>, d_cyc, false, (uint64_t *)d_out, 3u, 5u, big);
    run("imad_wide_P_product_plus_lop3", 8.0 * big, sms * 4, 256, k_prod<0>, d_cyc, false, (uint32_t *)d_out, 3u, 5u, big);
    run("imad_hi_32bit", 8.0 * big, sms * 4, 256, k_prod<1>, d_cyc, false, (uint32_t *)d_out, 3u, 5u, big);
    run("alu_iadd3", 8.0 * big, sms * 4, 256, k_prod<2>, d_cyc, false, (uint32_t *)d_out, 3u, 5u, big);
    run("imad_lo_32bit", 8.0 * big, sms * 4, 256, k_imad_lo, d_cyc, false, (uint32_t *)d_out, 3u, 5u, big);
    run("imad_wide_x_chain_rows", 12.0 * (big / 2), sms * 4, 256, k_imad_chain, d_cyc, false, (uint32_t *)d_out, 3u, 5u, big / 2);
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

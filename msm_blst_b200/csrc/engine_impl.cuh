// Host orchestration of the MSM pipeline for one group (F = fp_t: G1, F = fp2_t: G2). Included by
// engine_g1.cu / engine_g2.cu, which instantiate it once each (separate translation units so the two
// groups compile in parallel). All work is enqueued on the context's stream; nothing here computes on the CPU.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "engine.hpp"
#include "msm_kernels.cuh"

namespace msmb200 {

static inline unsigned blocks_for(size_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }

// Default accumulator (msmb200_set_accumulator(ctx, 0)): batch-affine rounds (2) once the problem is large enough that
// its rounds are throughput-bound, XYZZ work items (1) below — measured crossover on B200 (profiles/r2_accum_sweep.log and,
// on the final kernels, gpurun_out/r2ar log: sort + accumulate + reduce, the reduction being cheaper on affine bucket sums):
// G1 between 6.8 M and 12.6 M entries (n=2^19: 2.96 vs 3.03 ms, n=2^20: 5.68 vs 5.16 ms), G2 at about 0.5 M entries
// (n=2^15: 1.44 vs 1.38 ms, n=2^16: 2.58 vs 2.46 ms, n=2^17: 3.03 vs 2.86 ms).
template <class F> static inline int default_accumulator(size_t entries) {
    return entries >= (sizeof(F) > 48 ? ((size_t)1 << 19) : ((size_t)1 << 23)) ? 2 : 1;
}

struct Layout {
    size_t m;            // number of (key, val) entries
    uint32_t nbw;        // buckets per window (local index 0 unused)
    uint32_t nwindows;
    uint32_t wbits;      // doublings between windows
    const int *bucket_vals;  // device, or nullptr for dense (value = local index)
    int d_max;
    const int *chunk_first;  // device (sparse only): first local index whose value exceeds c * vspan, c = 0..nchunks
    uint32_t vspan;          // value span of one reduction chunk (power of two)
    uint32_t nchunks;        // reduction chunks per window (all of them)
    uint32_t chunk_lo = 0, chunk_cnt = 0;  // the slice this context reduces (bucket-range sharding); cnt 0 = all
    const ReducePlan *plan = nullptr;      // digit-splitting reduction plan (nullptr: chunked running sums only)
    uint32_t tstride16 = 0;                // table entry stride in 16-byte units (0: packed, sizeof(aff_t<F>) / 16)
};

static inline bool use_split_reduce(const Ctx *c, const Layout &L) {
    int mode = c->reduce_mode;
    if (c->reduce_env) mode = c->reduce_env;
    // bucket-range sharding reduces a RANGE of chunks: only the chunked reducer honours chunk_lo / chunk_cnt, so a sharded
    // context uses it (the digit-splitting plan covers all buckets and would redo the full reduction on every rank)
    return L.plan && L.plan->valid && mode != 1 && c->shard_world <= 1 && (L.nwindows == 1 || L.plan->nbits_w <= L.wbits);
}
// dense plans (value == local index) are built on first use and cached per (buckets per window, windows)
static inline const ReducePlan *dense_plan(Ctx *c, ReducePlan &slot, size_t nbw, uint32_t nwindows, uint32_t *key_windows) {
    if (!slot.valid || slot.key_nbw != nbw || *key_windows != nwindows) {
        if (build_reduce_plan(c, slot, nullptr, nbw, nwindows) != MSMB200_OK) return nullptr;
        *key_windows = nwindows;
    }
    return &slot;
}

// slice of the reduction chunks owned by this context and the matching local bucket-index range [lo, hi)
static inline void shard_slice(const Ctx *c, uint32_t nchunks, uint32_t *chunk_lo, uint32_t *chunk_cnt) {
    uint32_t lo = (uint32_t)((uint64_t)nchunks * c->shard_rank / c->shard_world);
    uint32_t hi = (uint32_t)((uint64_t)nchunks * (c->shard_rank + 1) / c->shard_world);
    *chunk_lo = lo;
    *chunk_cnt = hi - lo;
}

// value span of a reduction chunk: aim at ~32 K chunks in total, between 8 and 64 values per chunk
static inline uint32_t pick_vspan(const Ctx *c, size_t max_value, uint32_t nwindows) {
    if (c->vspan_env) return (uint32_t)c->vspan_env;
    // ~32 K chunks in total (measured optimum on B200: G1 n=2^21 -> 64, G2 n=2^18 -> 16), 8 <= v <= 64, power of two
    uint32_t v = 8;
    while (v < 64 && ((max_value + 1) * nwindows + v - 1) / v > 32768 + 1024) v <<= 1;
    return v;
}

static __global__ void iota_kernel(uint32_t *out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}

// Batch-affine bucket accumulation (batch_affine.cuh). Entries are in c->keys / c->vals / c->ranks with the bucket
// histogram in `count`; on return c->bucket_sum[b] holds the affine sum of every non-empty bucket. Steps: per-round
// totals (one small read-back: the number of rounds and the slot counts size every later launch exactly), the
// exclusive scans of all rounds in one pass, the counting-sort scatter, the slot descriptors, then one arithmetic
// kernel per round. Records ev[2] between planning and arithmetic ("sort" / "accumulate" phases).
template <class F>
static int ba_accumulate(Ctx *c, const aff_t<F> *d_table, uint32_t tstride16, uint32_t *count, size_t nb, size_t m) {
    cudaStream_t st = c->stream;
    if (ensure(c, c->ba_totals, (3 * BA_RMAX + 1) * 4)) return MSMB200_ECUDA;
    if (!c->h_totals) MSM_CUDA(c, cudaMallocHost((void **)&c->h_totals, (3 * BA_RMAX + 1) * 4));
    MSM_CUDA(c, cudaMemsetAsync(c->ba_totals.p, 0, (3 * BA_RMAX + 1) * 4, st));
    ba_totals_kernel<<<(unsigned)std::min<size_t>(blocks_for(nb, 256), (size_t)c->sms * 4), 256, 0, st>>>(count, nb, (uint32_t *)c->ba_totals.p);
    MSM_CUDA(c, cudaMemcpyAsync(c->h_totals, c->ba_totals.p, (3 * BA_RMAX + 1) * 4, cudaMemcpyDeviceToHost, st));
    MSM_CUDA(c, cudaStreamSynchronize(st));
    const uint32_t *T = c->h_totals;
    BaRounds rd{};
    uint32_t R = 0;
    while (R < BA_RMAX && T[3 * R] != 0) R++;
    const uint32_t Rs = std::max<uint32_t>(R, 1);  // round 0 always exists: single-entry buckets are copied there
    rd.R = Rs;
    size_t a_total = 0, c_total = 0, e_odd = 1, e_even = 1;
    for (uint32_t r = 0; r < Rs; r++) {
        rd.aoff[r] = (uint32_t)a_total; a_total += T[3 * r];
        rd.coff[r] = (uint32_t)c_total; c_total += T[3 * r + 1];
        if (r >= 1) (r & 1 ? e_odd : e_even) = std::max<size_t>(r & 1 ? e_odd : e_even, T[3 * r + 2]);
    }
    rd.aoff[Rs] = (uint32_t)a_total; rd.coff[Rs] = (uint32_t)c_total;
    const size_t ntiles = (nb + BA_TILE - 1) / BA_TILE, nbs = ntiles * BA_TILE, NR = 3 * (size_t)Rs;
    const size_t scratch_stride = ((size_t)T[0] + 31) & ~(size_t)31;
    if (ensure(c, c->ba_tile_sums, ntiles * NR * 4) || ensure(c, c->ba_bases, NR * nbs * 4) || ensure(c, c->ba_adesc, (a_total + 1) * 16) ||
        ensure(c, c->ba_cdesc, (c_total + 1) * 8) || ensure(c, c->ba_heavy, (m / BA_HEAVY + 2) * 4) || ensure(c, c->pts_a, e_odd * sizeof(aff_t<F>)) ||
        ensure(c, c->pts_b, e_even * sizeof(aff_t<F>)) || ensure(c, c->ba_scratch, (scratch_stride + 1) * sizeof(F)) ||
        ensure(c, c->bucket_sum, nb * sizeof(aff_t<F>)) || ensure(c, c->sorted, (m + 2) * 4))
        return MSMB200_ECUDA;
    if (c->iota_n < nb) {
        if (ensure(c, c->iota, nb * 4)) return MSMB200_ECUDA;
        iota_kernel<<<blocks_for(nb, 256), 256, 0, st>>>((uint32_t *)c->iota.p, nb);
        c->iota_n = nb;
    }
    const unsigned warp_blocks = blocks_for(ntiles * 32, 256);
    ba_scan_tiles_kernel<<<warp_blocks, 256, 0, st>>>(count, nb, Rs, (uint32_t *)c->ba_tile_sums.p, ntiles);
    ba_scan_sums_kernel<<<(unsigned)NR, 256, 0, st>>>((uint32_t *)c->ba_tile_sums.p, ntiles, (uint32_t)NR);
    ba_scan_apply_kernel<<<warp_blocks, 256, 0, st>>>(count, nb, Rs, (const uint32_t *)c->ba_tile_sums.p, ntiles, (uint32_t *)c->ba_bases.p, nbs);
    const uint32_t *bases = (const uint32_t *)c->ba_bases.p;
    scatter_kernel<<<blocks_for(m, 256), 256, 0, st>>>((const uint32_t *)c->keys.p, (const uint32_t *)c->vals.p, m, bases + 2 * nbs,
                                                       (const uint32_t *)c->ranks.p, (uint32_t *)c->sorted.p);
    MSM_CUDA(c, cudaMemsetAsync(c->ba_heavy.p, 0, 4, st));
    ba_emit_kernel<<<blocks_for(nb, 256), 256, 0, st>>>(count, nb, bases, nbs, rd, (const uint32_t *)c->sorted.p, (uint4 *)c->ba_adesc.p, (uint2 *)c->ba_cdesc.p,
                                                        (uint32_t *)c->ba_heavy.p, (uint4 *)c->bucket_sum.p, (uint32_t)(sizeof(aff_t<F>) / 16));
    if (T[3 * BA_RMAX] > BA_HEAVY)
        ba_emit_heavy_kernel<<<(unsigned)std::min<size_t>(m / BA_HEAVY + 1, (size_t)c->sms * 2), 256, 0, st>>>(count, bases, nbs, rd, (const uint32_t *)c->sorted.p,
                                                                                                          (uint4 *)c->ba_adesc.p, (uint2 *)c->ba_cdesc.p, (const uint32_t *)c->ba_heavy.p);
    c->launches += 8;
    MSM_CUDA(c, cudaEventRecord(c->ev[2], st));
    // ---- arithmetic rounds ----
    if (c->ba_resident <= 0) {
        int nbk = 0;
        MSM_CUDA(c, cudaFuncSetAttribute(ba_round_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ba_smem<F>::BYTES));
        MSM_CUDA(c, cudaFuncSetAttribute(ba_round_kernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ba_smem<F>::BYTES));
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nbk, ba_round_kernel<F, false>, BA_THREADS, ba_smem<F>::BYTES) != cudaSuccess || nbk < 1) nbk = 2;
        c->ba_resident = nbk;
    }
    const size_t wave = (size_t)c->sms * c->ba_resident;   // co-resident blocks: the round kernel is persistent
    const size_t nwarps = wave * (BA_THREADS / 32);
    if (ensure(c, c->ba_counters, (BA_RMAX + 1) * 4)) return MSMB200_ECUDA;
    if (!c->ba_sm_arrivals.p) {
        if (ensure(c, c->ba_sm_arrivals, 1024 * 4)) return MSMB200_ECUDA;
        MSM_CUDA(c, cudaMemsetAsync(c->ba_sm_arrivals.p, 0, 1024 * 4, st));
    }
    MSM_CUDA(c, cudaMemsetAsync(c->ba_counters.p, 0, (BA_RMAX + 1) * 4, st));
    // ping-pong point buffers between rounds, x[] and y[] separate: round r reads buffer r & 1 and writes buffer (r + 1) & 1
    F *bx[2] = {(F *)c->pts_b.p, (F *)c->pts_a.p};
    F *by[2] = {(F *)c->pts_b.p + e_even, (F *)c->pts_a.p + e_odd};
    for (uint32_t r = 0; r < Rs; r++) {
        const uint32_t A = T[3 * r], Cn = T[3 * r + 1];
        BaSched sched;
        sched.rows = (A + 31) / 32;
        const size_t share = (sched.rows + nwarps - 1) / nwarps;   // rows per warp if the round were split evenly
        if (share > (size_t)c->ba_batch_max) {   // several batches per warp: equal full batches, staggered start, shrinking tail
            const size_t nb_ = (share + c->ba_batch_max - 1) / c->ba_batch_max;
            sched.batch = (uint32_t)((share + nb_ - 1) / nb_);
            sched.stagger = c->ba_stagger ? 1u : 0u;
        } else {
            sched.batch = (uint32_t)std::max<size_t>(1, share);
            sched.stagger = 0;
        }
        if (c->ba_batch_fixed > 0) sched.batch = (uint32_t)c->ba_batch_fixed;
        sched.counter = (uint32_t *)c->ba_counters.p + r;
        sched.sm_arrivals = (uint32_t *)c->ba_sm_arrivals.p;
        const size_t batches = (sched.rows + sched.batch - 1) / sched.batch;
        const size_t add_blocks = (batches + BA_THREADS / 32 - 1) / (BA_THREADS / 32);
        const size_t copy_blocks = blocks_for(Cn, BA_THREADS);
        const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>(wave, std::max(add_blocks, copy_blocks)));
        const uint4 *ad = (const uint4 *)c->ba_adesc.p + rd.aoff[r];
        const uint2 *cd = (const uint2 *)c->ba_cdesc.p + rd.coff[r];
        ba_io<F> io;
        io.table = d_table;
        io.tstride16 = tstride16;
        io.in_x = bx[r & 1]; io.in_y = by[r & 1];
        io.out_x = bx[(r + 1) & 1]; io.out_y = by[(r + 1) & 1];
        io.bucket_sum = (aff_t<F> *)c->bucket_sum.p;
        if (r == 0)
            ba_round_kernel<F, true><<<grid, BA_THREADS, ba_smem<F>::BYTES, st>>>(io, ad, A, cd, Cn, (uint4 *)c->ba_scratch.p, scratch_stride, sched, r);
        else
            ba_round_kernel<F, false><<<grid, BA_THREADS, ba_smem<F>::BYTES, st>>>(io, ad, A, cd, Cn, (uint4 *)c->ba_scratch.p, scratch_stride, sched, r);
        c->launches += 1;
    }
    MSM_CUDA(c, cudaGetLastError());
    return MSMB200_OK;
}

// sort + accumulate + reduce + finalize over entries already in c->keys / c->vals with histogram in c->count
// F: field type of the hot kernels (multiplier inlined); FC: same layout, out-of-line multiplier, for the rest
template <class F, class FC>
static int run_buckets(Ctx *c, const Layout &L, const aff_t<F> *d_table, void *d_out_jac, bool want_affine) {
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)L.nbw * L.nwindows;
    const size_t m = L.m;
    const uint32_t tstride16 = L.tstride16 ? L.tstride16 : (uint32_t)(sizeof(aff_t<F>) / 16);
    // work-item length: long enough that a typical bucket is one item, short enough that there are at least
    // ~8 items per resident thread (SMs x 384 threads) even when buckets are few and heavy
    const size_t sms = (size_t)c->sms, subparts = 4 * sms;   // 148 / 592 on B200
    size_t avg = std::max<size_t>(1, m / std::max<size_t>(1, nb));
    uint32_t item_len = (uint32_t)std::min<size_t>(1024, std::max<size_t>(128, 8 * avg));
    if (nb < sms * 384 * 2) item_len = (uint32_t)std::max<size_t>(32, std::min<size_t>(item_len, m / (sms * 384 * 8)));
    // Small problems use short work items (>= ~3 warps per SM sub-partition of equal-length chains); the partials of a
    // split bucket are folded by one quad (combine_light_kernel) or, beyond heavy_items partials, by one block.
    const bool split_reduce = use_split_reduce(c, L);
    // buckets with 2..8 work items: one lane group each; 9..64: one warp of lane groups each (combine_light_kernel);
    // more: one block each (combine_heavy_kernel)
    const uint32_t heavy_items = 8, medium_items = 64;
    if (split_reduce) {
        const size_t want_items = (size_t)3 * subparts * 32;
        if (nb < want_items) item_len = (uint32_t)std::max<size_t>(8, std::min<size_t>(item_len, m / want_items));
    }
    // The longest chain is the critical path: with ~3 warps sharing a sub-partition it advances at a third of the pipe
    // rate, so keep item_len * 3 below half of the ideal duration of the whole phase (m / (sub-partitions * 32) additions per lane).
    item_len = (uint32_t)std::max<size_t>(8, std::min<size_t>(item_len, m / (subparts * 32 * 6)));
    if (c->item_len_fixed > 0) item_len = (uint32_t)c->item_len_fixed;
    const size_t max_items = std::min(nb, m) + m / item_len + 1;
    const size_t ntiles = (nb + SCAN_TILE - 1) / SCAN_TILE;

    int mode = c->accum_mode;
    if (c->accum_env) mode = c->accum_env;
    if (mode == 0) mode = default_accumulator<F>(m);
    c->last_accum = mode == 2 ? 2 : 1;
    if (ensure(c, c->sorted, (m + 2) * 4) || ensure(c, c->result, sizeof(jac_t<F>) + sizeof(aff_t<F>))) return MSMB200_ECUDA;
    if (mode != 2 &&   // workspace of the XYZZ work-item path only
        (ensure(c, c->packed, nb * 8) || ensure(c, c->scanned, nb * 8) || ensure(c, c->tile_sums, (ntiles + 1) * 8) ||
         ensure(c, c->seg_start, nb * 4) || ensure(c, c->item_start, nb * 4) || ensure(c, c->cursor, nb * 4) ||
         ensure(c, c->maxcount, 16) || ensure(c, c->item_begin, max_items * 4) || ensure(c, c->item_cnt, max_items * 4) ||
         ensure(c, c->order, max_items * 4) || ensure(c, c->len_hist, (item_len + 1) * 4) ||
         ensure(c, c->len_start, (item_len + 1) * 4) || ensure(c, c->len_cursor, (item_len + 1) * 4) ||
         ensure(c, c->partial, max_items * sizeof(xyzz_t<F>)) || ensure(c, c->heavy, (m / item_len + 2) * 4)))
        return MSMB200_ECUDA;

    uint32_t *count = (uint32_t *)c->count.p;
    // ---- sort by bucket ----
    const bool batch_affine = mode == 2;
    if (!batch_affine) {
        MSM_CUDA(c, cudaMemsetAsync(c->maxcount.p, 0, 4, st));
        prep_counts_kernel<<<blocks_for(nb, 256), 256, 0, st>>>(count, (uint64_t *)c->packed.p, nb, item_len, (uint32_t *)c->maxcount.p);
        scan_tiles_kernel<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>((const uint64_t *)c->packed.p, (uint64_t *)c->scanned.p,
                                                                       (uint64_t *)c->tile_sums.p, nb);
        scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>((uint64_t *)c->tile_sums.p, ntiles);
        scan_finish_kernel<<<blocks_for(nb, 256), 256, 0, st>>>((const uint64_t *)c->scanned.p, (const uint64_t *)c->tile_sums.p,
                                                                (uint32_t *)c->seg_start.p, (uint32_t *)c->item_start.p,
                                                                (uint32_t *)c->cursor.p, nb);
        scatter_kernel<<<blocks_for(m, 256), 256, 0, st>>>((const uint32_t *)c->keys.p, (const uint32_t *)c->vals.p, m,
                                                           (const uint32_t *)c->seg_start.p, (const uint32_t *)c->ranks.p,
                                                           (uint32_t *)c->sorted.p);
    }
    const uint64_t *totals = (const uint64_t *)c->tile_sums.p + ntiles;
    const void *bucket_points = nullptr;      // what the reduction reads: XYZZ partials or affine points
    const uint32_t *bucket_point_index = nullptr;
    if (!batch_affine) {
        MSM_CUDA(c, cudaMemsetAsync(c->len_hist.p, 0, (item_len + 1) * 4, st));
        MSM_CUDA(c, cudaMemsetAsync(c->heavy.p, 0, 4, st));
        if (ensure(c, c->light, (std::min(nb, m / item_len + 1) + 2) * 4) || ensure(c, c->medium, (std::min(nb, m / item_len + 1) + 2) * 4)) return MSMB200_ECUDA;
        MSM_CUDA(c, cudaMemsetAsync(c->light.p, 0, 4, st));
        MSM_CUDA(c, cudaMemsetAsync(c->medium.p, 0, 4, st));
        itemize_kernel<<<blocks_for(nb, 256), 256, 0, st>>>(count, (const uint32_t *)c->seg_start.p, (const uint32_t *)c->item_start.p,
                                                            nb, item_len, (uint32_t *)c->item_begin.p, (uint32_t *)c->item_cnt.p,
                                                            (uint32_t *)c->len_hist.p, (uint32_t *)c->heavy.p, heavy_items, (uint32_t *)c->light.p, medium_items,
                                                            (uint32_t *)c->medium.p);
        len_scan_kernel<<<1, 32, 0, st>>>((const uint32_t *)c->len_hist.p, (uint32_t *)c->len_start.p, (uint32_t *)c->len_cursor.p, item_len);
        order_items_kernel<<<blocks_for(max_items, 256), 256, 0, st>>>((const uint32_t *)c->item_cnt.p, totals,
                                                                       (const uint32_t *)c->len_start.p, (uint32_t *)c->len_cursor.p,
                                                                       (uint32_t *)c->order.p);
        c->launches += 8;
        MSM_CUDA(c, cudaEventRecord(c->ev[2], st));
        // ---- accumulate: XYZZ mixed additions, one thread per work item ----
        accumulate_kernel<F><<<blocks_for(max_items, 128), 128, 0, st>>>(d_table, tstride16, (const uint32_t *)c->sorted.p,
                                                                             (const uint32_t *)c->item_begin.p, (const uint32_t *)c->item_cnt.p,
                                                                             (const uint32_t *)c->order.p, totals, (xyzz_t<F> *)c->partial.p);
        {
            unsigned max_heavy = (unsigned)std::min<size_t>(m / item_len + 1, subparts);
            combine_heavy_kernel<FC><<<max_heavy, 128, 0, st>>>(count, (const uint32_t *)c->item_start.p, (const uint32_t *)c->heavy.p, item_len,
                                                               (xyzz_t<FC> *)c->partial.p);
        }
        {
            // at most min(nb, m / item_len) buckets have more than one work item (grids sized for that bound, idle warps
            // exit); medium buckets have more than heavy_items work items each
            using CT = typename coop_of<FC>::type;
            constexpr uint32_t GL = coop_group_lanes<CT>(), NG = 32 / GL;
            const size_t max_light = std::min(nb, m / item_len + 1), max_medium = std::min(nb, m / ((size_t)item_len * heavy_items) + 1);
            combine_light_kernel<FC><<<blocks_for(max_medium * 32, 128), 128, 0, st>>>(count, (const uint32_t *)c->item_start.p, (const uint32_t *)c->medium.p,
                                                                                    item_len, NG, (xyzz_t<FC> *)c->partial.p);
            combine_light_kernel<FC><<<blocks_for(max_light * GL, 128), 128, 0, st>>>(count, (const uint32_t *)c->item_start.p, (const uint32_t *)c->light.p,
                                                                                   item_len, 1, (xyzz_t<FC> *)c->partial.p);
            c->launches += 2;
        }
        c->launches += 2;
        bucket_points = c->partial.p;
        bucket_point_index = (const uint32_t *)c->item_start.p;
    } else {
        int rc = ba_accumulate<F>(c, d_table, tstride16, count, nb, m);
        if (rc) return rc;
        bucket_points = c->bucket_sum.p;
        bucket_point_index = (const uint32_t *)c->iota.p;
    }
    MSM_CUDA(c, cudaEventRecord(c->ev[3], st));
    // ---- reduce ----
    if (split_reduce) {
        const ReducePlan &P = *L.plan;
        const uint32_t nw = L.nwindows;
        if (ensure(c, c->red_a, (size_t)nw * P.s1.nlists * sizeof(xyzz_t<F>) + 64) || ensure(c, c->red_b, (size_t)nw * P.s1b.nlists * sizeof(xyzz_t<F>) + 64) ||
            ensure(c, c->red_c, (size_t)nw * P.s2a.nlists * sizeof(xyzz_t<F>) + 64) || ensure(c, c->red_d, (size_t)nw * P.s2b.nlists * sizeof(xyzz_t<F>) + 64))
            return MSMB200_ECUDA;
        // stage 1: one lane per slice of a digit list, gathering the bucket sums
        if (P.s1.nlists) {
            const unsigned g1 = blocks_for((size_t)nw * P.s1.nlists, 128);
            if (batch_affine)
                list_sum_kernel<F, 1><<<g1, 128, 0, st>>>(bucket_points, count, bucket_point_index, L.nbw,
                                                          (const uint32_t *)P.s1.start.p, (const uint32_t *)P.s1.idx.p, P.s1.nlists, nw,
                                                          (xyzz_t<F> *)c->red_a.p);
            else if (P.s1_coop)
                list_sum_coop_kernel<FC, 0><<<blocks_for((size_t)nw * P.s1.nlists * coop_group_lanes<typename coop_of<FC>::type>(), 128), 128, 0, st>>>(
                    (const xyzz_t<FC> *)bucket_points, count, bucket_point_index, L.nbw, (const uint32_t *)P.s1.start.p,
                    (const uint32_t *)P.s1.idx.p, P.s1.nlists, nw, 1, (xyzz_t<FC> *)c->red_a.p);
            else
                list_sum_kernel<F, 0><<<g1, 128, 0, st>>>(bucket_points, count, bucket_point_index, L.nbw,
                                                          (const uint32_t *)P.s1.start.p, (const uint32_t *)P.s1.idx.p, P.s1.nlists, nw,
                                                          (xyzz_t<F> *)c->red_a.p);
        }
        // stages 1b, 2a, 2b: quad-cooperative list sums over dense arrays
        auto coop = [&](const ListPlan &lp, const void *src, uint32_t in_stride, void *dst) {
            list_sum_coop_kernel<FC, 2><<<blocks_for((size_t)nw * lp.nlists * lp.tl * coop_group_lanes<typename coop_of<FC>::type>(), 128), 128, 0, st>>>(
                (const xyzz_t<FC> *)src, nullptr, nullptr, in_stride, (const uint32_t *)lp.start.p, (const uint32_t *)lp.idx.p, lp.nlists, nw,
                lp.tl, (xyzz_t<FC> *)dst);
        };
        coop(P.s1b, c->red_a.p, P.s1.nlists, c->red_b.p);
        coop(P.s2a, c->red_b.p, P.s1b.nlists, c->red_c.p);
        coop(P.s2b, c->red_c.p, P.s2a.nlists, c->red_d.p);
        c->launches += 4;
        MSM_CUDA(c, cudaEventRecord(c->ev[4], st));
        if (c->bits_out) {   // multi-GPU leg: hand out the per-bit sums; Horner and to_affine run once, after the all-gather
            MSM_CUDA(c, cudaMemcpyAsync(c->bits_out, c->red_d.p, (size_t)nw * P.nbits_w * sizeof(xyzz_t<F>), cudaMemcpyDeviceToDevice, st));
            MSM_CUDA(c, cudaEventRecord(c->ev[5], st));
            MSM_CUDA(c, cudaGetLastError());
            return MSMB200_OK;
        }
        jac_t<F> *d_jac = d_out_jac ? (jac_t<F> *)d_out_jac : (jac_t<F> *)c->result.p;
        aff_t<F> *d_aff = (aff_t<F> *)((char *)c->result.p + sizeof(jac_t<F>));
        bits_finalize_coop_kernel<FC><<<1, 32, 0, st>>>((const xyzz_t<FC> *)c->red_d.p, nw, P.nbits_w, L.wbits, (xyzz_t<FC> *)c->red_c.p,
                                                            (jac_t<FC> *)d_jac, want_affine ? (aff_t<FC> *)d_aff : nullptr);
        c->launches += 1;
        if (want_affine) MSM_CUDA(c, cudaMemcpyAsync(c->h_result, d_aff, sizeof(aff_t<F>), cudaMemcpyDeviceToHost, st));
        MSM_CUDA(c, cudaEventRecord(c->ev[5], st));
        MSM_CUDA(c, cudaGetLastError());
        if (want_affine) MSM_CUDA(c, cudaStreamSynchronize(st));
        return MSMB200_OK;
    }
    if (c->bits_out) return ctx_fail(c, MSMB200_ESTATE, "per-bit partial sums need the digit-splitting reducer");
    const uint32_t cpw = L.chunk_cnt ? L.chunk_cnt : L.nchunks;
    size_t nchunks = (size_t)cpw * L.nwindows;
    if (ensure(c, c->chunk_a, 2 * nchunks * sizeof(xyzz_t<F>)) || ensure(c, c->chunk_b, (2 * ((size_t)cpw / 4 + 2) * L.nwindows + 2) * sizeof(xyzz_t<F>)))
        return MSMB200_ECUDA;
#define MSM_REDUCE_LAUNCH(DENSE_, AFF_)                                                                                             \
    reduce_chunks_kernel<F, DENSE_, AFF_><<<blocks_for(nchunks, 64), 64, 0, st>>>(                                                  \
        bucket_points, count, bucket_point_index, L.bucket_vals, L.chunk_first, L.nbw, L.nwindows, L.vspan, cpw, L.chunk_lo, L.bucket_vals ? L.d_max : 1, \
        (xyzz_t<F> *)c->chunk_a.p)
    if (L.bucket_vals) { if (batch_affine) MSM_REDUCE_LAUNCH(false, true); else MSM_REDUCE_LAUNCH(false, false); }
    else { if (batch_affine) MSM_REDUCE_LAUNCH(true, true); else MSM_REDUCE_LAUNCH(true, false); }
#undef MSM_REDUCE_LAUNCH
    c->launches += 1;
    // tree of sums over the 2 rows of every window
    xyzz_t<F> *cur = (xyzz_t<F> *)c->chunk_a.p, *nxt = (xyzz_t<F> *)c->chunk_b.p;
    uint32_t per = cpw;
    const uint32_t nrows = 2 * L.nwindows;
    while (per > 1024) {
        uint32_t r = 8;
        uint32_t groups = (per + r - 1) / r;
        sum_groups_kernel<FC><<<blocks_for((size_t)groups * nrows, 128), 128, 0, st>>>((const xyzz_t<FC> *)cur, per, nrows, r, groups, (xyzz_t<FC> *)nxt);
        c->launches += 1;
        std::swap(cur, nxt);
        per = groups;
    }
    if (per > 1) {
        tree_tail_kernel<FC><<<nrows, 256, 0, st>>>((xyzz_t<FC> *)cur, per, (xyzz_t<FC> *)nxt);
        c->launches += 1;
        std::swap(cur, nxt);
        per = 1;
    }
    MSM_CUDA(c, cudaEventRecord(c->ev[4], st));
    // ---- finalize ----
    jac_t<F> *d_jac = d_out_jac ? (jac_t<F> *)d_out_jac : (jac_t<F> *)c->result.p;
    aff_t<F> *d_aff = (aff_t<F> *)((char *)c->result.p + sizeof(jac_t<F>));
    uint32_t vshift = 0;
    while ((1u << vshift) < L.vspan) vshift++;
    finalize_kernel<FC><<<1, 32, 0, st>>>((const xyzz_t<FC> *)cur, L.nwindows, L.wbits, vshift, (jac_t<FC> *)d_jac, want_affine ? (aff_t<FC> *)d_aff : nullptr);
    c->launches += 1;
    if (want_affine) MSM_CUDA(c, cudaMemcpyAsync(c->h_result, d_aff, sizeof(aff_t<F>), cudaMemcpyDeviceToHost, st));
    MSM_CUDA(c, cudaEventRecord(c->ev[5], st));
    MSM_CUDA(c, cudaGetLastError());
    if (want_affine) MSM_CUDA(c, cudaStreamSynchronize(st));
    return MSMB200_OK;
}

static int prepare_entries(Ctx *c, size_t m, size_t nb) {
    if (ensure(c, c->keys, m * 4) || ensure(c, c->vals, m * 4) || ensure(c, c->ranks, m * 4) || ensure(c, c->count, nb * 4)) return MSMB200_ECUDA;
    MSM_CUDA(c, cudaMemsetAsync(c->count.p, 0, nb * 4, c->stream));
    return MSMB200_OK;
}

static inline bool bgmw_trick(const msmb200_config &cfg) {
    return cfg.n_exp == 13 || cfg.n_exp == 14 || cfg.n_exp == 16 || cfg.n_exp == 17;  // main_p1.cpp:311
}

template <class F, class FC>
static int pippenger_impl(Ctx *c, const void *d_points, size_t npoints, const void *d_scalars, int nbits, void *d_out_jac,
                          bool want_affine, int wbits_table, int tile_bit0, int tile_window) {
    // wbits_table == 0: blst's Pippenger over the points themselves; > 0: blst_p1s_mult_wbits over a precomputed table of
    // 2^(wbits-1) multiples per point (d_points is that table), one bucket per window.
    // tile_bit0 >= 0: only the tile of tile_window bits starting at that bit (blst_p1s_tile_pippenger), not shifted.
    cudaStream_t st = c->stream;
    c->launches = 0;
    MSM_CUDA(c, cudaEventRecord(c->ev[0], st));
    int w = tile_bit0 >= 0 ? tile_window : wbits_table ? wbits_table : (int)pippenger_window_size(npoints);
    int tiles = tile_bit0 >= 0 ? 1 : nbits / w + 1;
    uint32_t nbw = wbits_table ? 2u : (1u << (w - 1)) + 1u;
    size_t m = npoints * (size_t)tiles, nb = (size_t)nbw * tiles;
    int rc = prepare_entries(c, m, nb);
    if (rc) return rc;
    const uint32_t vs = pick_vspan(c, nbw - 1, (uint32_t)tiles), nch = (nbw - 1 + vs - 1) / vs;
    uint32_t clo, ccnt;
    shard_slice(c, nch, &clo, &ccnt);
    const uint32_t blo = 1 + clo * vs, bhi = std::min<uint32_t>(nbw, 1 + (clo + ccnt) * vs);
    digits_booth_kernel<<<blocks_for(npoints, 256), 256, 0, st>>>((const uint32_t *)d_scalars, npoints, nbits, w, tiles,
                                                                  (uint32_t *)c->keys.p, (uint32_t *)c->vals.p, (uint32_t *)c->count.p, (uint32_t *)c->ranks.p, 1,
                                                                  wbits_table ? 1u : blo, wbits_table ? 0xffffffffu : bhi, wbits_table ? 1u << (w - 1) : 0u, tile_bit0);
    c->launches += 1;
    MSM_CUDA(c, cudaEventRecord(c->ev[1], st));
    Layout L{m, nbw, (uint32_t)tiles, (uint32_t)w, nullptr, 1, nullptr, vs, nch};
    L.chunk_lo = clo; L.chunk_cnt = ccnt;
    L.plan = dense_plan(c, c->plan_pip, nbw, (uint32_t)tiles, &c->plan_pip_windows);
    return run_buckets<F, FC>(c, L, (const aff_t<F> *)d_points, d_out_jac, want_affine);
}

template <class F, class FC> static int msm_impl(Ctx *c, int method, const void *d_scalars, void *d_out_jac, bool want_affine) {
    cudaStream_t st = c->stream;
    const msmb200_config &cfg = c->cfg;
    const size_t n = c->n;
    if (method == MSMB200_PIPPENGER) {
        if (!c->have_points) return ctx_fail(c, MSMB200_ESTATE, "fixed points not set");
        return pippenger_impl<F, FC>(c, c->d_points, n, d_scalars, 255, d_out_jac, want_affine, 0, -1, 0);
    }
    c->launches = 0;
    MSM_CUDA(c, cudaEventRecord(c->ev[0], st));
    if (method == MSMB200_CHES || method == MSMB200_CHES_INTEGRAL) {
        if (!c->have_ches) return ctx_fail(c, MSMB200_ESTATE, "CHES table not built");
        size_t m = n * (size_t)cfg.h, nb = c->bucket_set.size();
        int rc = prepare_entries(c, m, nb);
        if (rc) return rc;
        uint32_t clo, ccnt;
        shard_slice(c, c->red_nchunks, &clo, &ccnt);
        const uint32_t blo = (uint32_t)c->h_chunk_first[clo], bhi = (uint32_t)c->h_chunk_first[clo + ccnt];
        if (method == MSMB200_CHES) {
            // host-to-host call: upload the scalars in chunks on the copy stream and decompose each chunk as it lands
            const int K = c->h_scalars_pending ? (int)std::min<size_t>(8, std::max<size_t>(1, (n * 32) >> 22)) : 1;  // >= 4 MB per chunk
            for (int k = 0; k < K; k++) {
                const size_t i0 = n * (size_t)k / K, i1 = n * (size_t)(k + 1) / K;
                if (i1 == i0) continue;
                if (c->h_scalars_pending) {
                    MSM_CUDA(c, cudaMemcpyAsync((char *)const_cast<void *>(d_scalars) + i0 * 32, (const char *)c->h_scalars_pending + i0 * 32, (i1 - i0) * 32,
                                                cudaMemcpyHostToDevice, c->copy_stream));
                    MSM_CUDA(c, cudaEventRecord(c->ev_chunk[k], c->copy_stream));
                    MSM_CUDA(c, cudaStreamWaitEvent(st, c->ev_chunk[k], 0));
                }
                digits_ches_kernel<<<blocks_for(i1 - i0, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, cfg.h, cfg.e, c->d_dtab, (uint32_t *)c->keys.p,
                                                                           (uint32_t *)c->vals.p, (uint32_t *)c->count.p, (uint32_t *)c->ranks.p, 1, blo, bhi, i0, i1 - i0);
                c->launches += 1;
            }
        } else {
            if (ensure(c, c->flat, (m + 2) * 4) || ensure(c, c->signs, m) || ensure(c, c->pidx, m * 4)) return MSMB200_ECUDA;
            digits_std_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, cfg.h, cfg.e, (int *)c->flat.p);
            construct_nh_kernel<<<blocks_for(n, 256), 256, 0, st>>>((int *)c->flat.p, (unsigned char *)c->signs.p, (uint32_t *)c->pidx.p,
                                                                    n, cfg.h, c->d_dtab, c->d_bucket_vals);
            tile_lookup_kernel<<<blocks_for(m, 256), 256, 0, st>>>((const int *)c->flat.p, (const unsigned char *)c->signs.p,
                                                                   (const uint32_t *)c->pidx.p, m, c->d_v2i, (uint32_t *)c->keys.p,
                                                                   (uint32_t *)c->vals.p, (uint32_t *)c->count.p, (uint32_t *)c->ranks.p, blo, bhi);
            c->launches += 3;
        }
        MSM_CUDA(c, cudaEventRecord(c->ev[1], st));
        Layout L{m, (uint32_t)nb, 1, 0, c->d_bucket_vals, cfg.d, c->d_chunk_first, c->red_vspan, c->red_nchunks};
        L.chunk_lo = clo; L.chunk_cnt = ccnt;
        L.plan = &c->plan_ches;
        L.tstride16 = c->table_stride / 16;
        return run_buckets<F, FC>(c, L, (const aff_t<F> *)c->d_table_ches, d_out_jac, want_affine);
    }
    if (method == MSMB200_BGMW95) {
        if (!c->have_bgmw) return ctx_fail(c, MSMB200_ESTATE, "BGMW95 table not built");
        size_t m = n * (size_t)cfg.h_bgmw;
        uint32_t nbw = (1u << (cfg.e_bgmw - 1)) + 1u;
        int rc = prepare_entries(c, m, nbw);
        if (rc) return rc;
        const uint32_t vs = pick_vspan(c, nbw - 1, 1), nch = (nbw - 1 + vs - 1) / vs;
        uint32_t clo, ccnt;
        shard_slice(c, nch, &clo, &ccnt);
        const uint32_t blo = 1 + clo * vs, bhi = std::min<uint32_t>(nbw, 1 + (clo + ccnt) * vs);
        digits_bgmw_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, cfg.h_bgmw, cfg.e_bgmw,
                                                               bgmw_trick(cfg) ? 1 : 0, (uint32_t *)c->keys.p, (uint32_t *)c->vals.p,
                                                               (uint32_t *)c->count.p, (uint32_t *)c->ranks.p, 1, blo, bhi);
        c->launches += 1;
        MSM_CUDA(c, cudaEventRecord(c->ev[1], st));
        Layout L{m, nbw, 1, 0, nullptr, 1, nullptr, vs, nch};
        L.chunk_lo = clo; L.chunk_cnt = ccnt;
        L.plan = dense_plan(c, c->plan_bgmw, nbw, 1, &c->plan_bgmw_windows);
        L.tstride16 = c->table_stride / 16;
        return run_buckets<F, FC>(c, L, (const aff_t<F> *)c->d_table_bgmw, d_out_jac, want_affine);
    }
    return ctx_fail(c, MSMB200_EINVAL, "unknown method");
}

// generic tile over caller-provided (bucket value|index, sign, point index) arrays: the device half of the
// blst_p1_tile_pippenger_d_CHES / _BGMW95 shims
template <class F, class FC>
static int tile_impl(Ctx *c, const void *d_table, const int *d_bvals, const unsigned char *d_signs, const uint32_t *d_pidx, size_t m,
                     const int *d_v2i, const int *d_bucket_vals, size_t nbuckets, int d_max, const int *d_chunk_first, uint32_t vspan,
                     uint32_t nchunks, const ReducePlan *plan, void *d_out_jac) {
    cudaStream_t st = c->stream;
    c->launches = 0;
    MSM_CUDA(c, cudaEventRecord(c->ev[0], st));
    int rc = prepare_entries(c, m, nbuckets);
    if (rc) return rc;
    tile_lookup_kernel<<<blocks_for(m, 256), 256, 0, st>>>(d_bvals, d_signs, d_pidx, m, d_v2i, (uint32_t *)c->keys.p, (uint32_t *)c->vals.p,
                                                           (uint32_t *)c->count.p, (uint32_t *)c->ranks.p, 0u, 0xffffffffu);
    c->launches += 1;
    MSM_CUDA(c, cudaEventRecord(c->ev[1], st));
    Layout L{m, (uint32_t)nbuckets, 1, 0, d_bucket_vals, d_max, d_chunk_first, vspan, nchunks};
    L.plan = plan;
    return run_buckets<F, FC>(c, L, (const aff_t<F> *)d_table, d_out_jac, false);
}

// device halves of the literal shims that work on the caller's host arrays (api.cu)
static int shim_aux_impl(Ctx *c, int what, const ShimAux &x) {
    cudaStream_t st = c->stream;
    if (what == 0) construct_nh_triples_kernel<<<blocks_for(x.m, 256), 256, 0, st>>>(x.in, x.out_b, x.signs, x.ptrs, x.m, x.triples, x.base, x.entry_bytes);
    else if (what == 1) ptrs_to_index_kernel<<<blocks_for(x.m, 256), 256, 0, st>>>(x.ptrs, x.m, x.base, x.entry_bytes, x.entries, x.pidx, x.bad);
    else if (what == 2) max_int_kernel<<<blocks_for(x.m, 256), 256, 0, st>>>(x.in, x.m, x.out_b);
    else minmax_ptr_kernel<<<blocks_for(x.m, 256), 256, 0, st>>>(x.ptrs, x.m, (unsigned long long *)x.pidx);
    MSM_CUDA(c, cudaGetLastError());
    return MSMB200_OK;
}

template <class F> static int sum_partials_impl(Ctx *c, const void *d_partials, int count) {
    if (ensure(c, c->result, sizeof(jac_t<F>) + sizeof(aff_t<F>)) || ensure(c, c->red_c, sizeof(xyzz_t<F>) + 64)) return MSMB200_ECUDA;
    aff_t<F> *d_aff = (aff_t<F> *)((char *)c->result.p + sizeof(jac_t<F>));
    sum_partials_coop_kernel<F><<<1, 32, 0, c->stream>>>((const jac_t<F> *)d_partials, count, (xyzz_t<F> *)c->red_c.p, d_aff);
    MSM_CUDA(c, cudaMemcpyAsync(c->h_result, d_aff, sizeof(aff_t<F>), cudaMemcpyDeviceToHost, c->stream));
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    return MSMB200_OK;
}

// second half of the multi-GPU combine: per-entry sum over the ranks, one Horner pass, one inversion
template <class F> static int combine_bits_impl(Ctx *c, const void *d_gathered, int world, uint32_t nwindows, uint32_t nbits_w, uint32_t wbits) {
    using CT = typename coop_of<F>::type;
    const uint32_t entries = nwindows * nbits_w;
    if (ensure(c, c->result, sizeof(jac_t<F>) + sizeof(aff_t<F>)) || ensure(c, c->red_c, sizeof(xyzz_t<F>) + 64) ||
        ensure(c, c->red_d, (size_t)entries * sizeof(xyzz_t<F>) + 64))
        return MSMB200_ECUDA;
    aff_t<F> *d_aff = (aff_t<F> *)((char *)c->result.p + sizeof(jac_t<F>));
    sum_ranks_coop_kernel<F><<<blocks_for((size_t)entries * coop_group_lanes<CT>(), 128), 128, 0, c->stream>>>((const xyzz_t<F> *)d_gathered, entries, (uint32_t)world,
                                                                                                         (xyzz_t<F> *)c->red_d.p);
    bits_finalize_coop_kernel<F><<<1, 32, 0, c->stream>>>((const xyzz_t<F> *)c->red_d.p, nwindows, nbits_w, wbits, (xyzz_t<F> *)c->red_c.p, nullptr, d_aff);
    MSM_CUDA(c, cudaMemcpyAsync(c->h_result, d_aff, sizeof(aff_t<F>), cudaMemcpyDeviceToHost, c->stream));
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    return MSMB200_OK;
}

template <class F> static int table_build_impl(Ctx *c, int which) {
    if (!c->have_points) return ctx_fail(c, MSMB200_ESTATE, "fixed points not set");
    const msmb200_config &cfg = c->cfg;
    int h = which == 0 ? cfg.h : cfg.h_bgmw, e = which == 0 ? cfg.e : cfg.e_bgmw, nmult = which == 0 ? 3 : 1;
    size_t entries = c->n * (size_t)h * nmult;
    void **slot = which == 0 ? &c->d_table_ches : &c->d_table_bgmw;
    if (!*slot) {
        MSM_CUDA(c, cudaMalloc(slot, entries * (size_t)c->table_stride));
        if (c->table_stride != sizeof(aff_t<F>)) MSM_CUDA(c, cudaMemsetAsync(*slot, 0, entries * (size_t)c->table_stride, c->stream));   // the padding is never read as data
    }
    table_build_kernel<F><<<blocks_for(c->n, 128), 128, 0, c->stream>>>((const aff_t<F> *)c->d_points, c->n, h, e, nmult, (aff_t<F> *)*slot, c->table_stride / 16);
    MSM_CUDA(c, cudaGetLastError());
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    (which == 0 ? c->have_ches : c->have_bgmw) = true;
    return MSMB200_OK;
}

// 2^k mod r on the host (k up to ~2^22): plain double-and-reduce on 4 x u64
static void pow2_mod_r(uint64_t out[4], size_t k) {
    static const uint64_t R[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    uint64_t v[4] = {1, 0, 0, 0};
    for (size_t i = 0; i < k; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) { uint64_t nc = v[j] >> 63; v[j] = (v[j] << 1) | carry; carry = nc; }
        bool ge = carry != 0;
        if (!ge) { ge = true; for (int j = 3; j >= 0; j--) { if (v[j] != R[j]) { ge = v[j] > R[j]; break; } } }
        if (ge) { unsigned __int128 b = 0; for (int j = 0; j < 4; j++) { unsigned __int128 t = (unsigned __int128)v[j] - R[j] - (uint64_t)b; v[j] = (uint64_t)t; b = (t >> 64) & 1; } }
    }
    memcpy(out, v, 32);
}

template <class F> static const uint32_t *generator_words();
template <class F> static int generate_fix_points_impl(Ctx *c, size_t first) {
    const uint32_t chunk = 48;  // doublings per thread (multiple of 3)
    size_t nthreads = (c->n + chunk - 1) / chunk;
    // seeds: 2^(first + t*chunk) mod r as scalars, multiplied by G on the device
    std::vector<uint64_t> sc(nthreads * 4);
    {
        // incremental: s_{t+1} = s_t * 2^chunk
        uint64_t cur[4];
        pow2_mod_r(cur, first);
        static const uint64_t R[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
        for (size_t t = 0; t < nthreads; t++) {
            memcpy(&sc[4 * t], cur, 32);
            for (uint32_t k = 0; k < chunk; k++) {
                uint64_t carry = 0;
                for (int j = 0; j < 4; j++) { uint64_t nc = cur[j] >> 63; cur[j] = (cur[j] << 1) | carry; carry = nc; }
                bool ge = carry != 0;
                if (!ge) { ge = true; for (int j = 3; j >= 0; j--) { if (cur[j] != R[j]) { ge = cur[j] > R[j]; break; } } }
                if (ge) { unsigned __int128 b = 0; for (int j = 0; j < 4; j++) { unsigned __int128 x = (unsigned __int128)cur[j] - R[j] - (uint64_t)b; cur[j] = (uint64_t)x; b = (x >> 64) & 1; } }
            }
        }
    }
    void *d_sc = nullptr, *d_gen = nullptr, *d_seeds = nullptr;
    MSM_CUDA(c, cudaMalloc(&d_sc, nthreads * 32));
    MSM_CUDA(c, cudaMalloc(&d_gen, sizeof(aff_t<F>)));
    MSM_CUDA(c, cudaMalloc(&d_seeds, nthreads * sizeof(jac_t<F>)));
    MSM_CUDA(c, cudaMemcpyAsync(d_sc, sc.data(), nthreads * 32, cudaMemcpyHostToDevice, c->stream));
    MSM_CUDA(c, cudaMemcpyAsync(d_gen, generator_words<F>(), sizeof(aff_t<F>), cudaMemcpyHostToDevice, c->stream));
    scalar_mul_kernel<F><<<blocks_for(nthreads, 128), 128, 0, c->stream>>>((const aff_t<F> *)d_gen, 0, (const uint32_t *)d_sc, nthreads,
                                                                           (jac_t<F> *)d_seeds);
    fix_points_kernel<F><<<blocks_for(nthreads, 128), 128, 0, c->stream>>>((const jac_t<F> *)d_seeds, nthreads, chunk, c->n,
                                                                           (aff_t<F> *)c->d_points);
    MSM_CUDA(c, cudaGetLastError());
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(d_sc); cudaFree(d_gen); cudaFree(d_seeds);
    c->have_points = true;
    return MSMB200_OK;
}

template <class F> static int digits_impl(Ctx *c, int kind, const void *d_scalars, size_t n, uint32_t *d_keys, uint32_t *d_vals) {
    cudaStream_t st = c->stream;
    const msmb200_config &cfg = c->cfg;
    size_t nb = kind == 0 ? c->bucket_set.size() : kind == 1 ? ((size_t)1 << (cfg.e_bgmw - 1)) + 1 : (size_t)c->pip_tiles * (((size_t)1 << (c->pip_window - 1)) + 1);
    if (ensure(c, c->count, nb * 4)) return MSMB200_ECUDA;
    MSM_CUDA(c, cudaMemsetAsync(c->count.p, 0, nb * 4, st));
    if (kind == 0)
        digits_ches_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, cfg.h, cfg.e, c->d_dtab, d_keys, d_vals, (uint32_t *)c->count.p, nullptr, 0, 0u, 0xffffffffu, 0, n);
    else if (kind == 1)
        digits_bgmw_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, cfg.h_bgmw, cfg.e_bgmw, bgmw_trick(cfg) ? 1 : 0, d_keys, d_vals, (uint32_t *)c->count.p, nullptr, 0, 0u, 0xffffffffu);
    else
        digits_booth_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t *)d_scalars, n, 255, c->pip_window, c->pip_tiles, d_keys, d_vals, (uint32_t *)c->count.p, nullptr, 0, 0u, 0xffffffffu, 0u, -1);
    MSM_CUDA(c, cudaGetLastError());
    MSM_CUDA(c, cudaStreamSynchronize(st));
    return MSMB200_OK;
}

template <class F> static int wbits_precompute_impl(Ctx *c, const void *d_points, size_t npoints, int wbits, void *d_table) {
    const size_t total = npoints << (wbits - 1);
    wbits_precompute_kernel<F><<<blocks_for(total, 128), 128, 0, c->stream>>>((const aff_t<F> *)d_points, npoints, wbits, (aff_t<F> *)d_table);
    MSM_CUDA(c, cudaGetLastError());
    return MSMB200_OK;
}

template <class F> static int table_io_impl(Ctx *c, int dir, int serialized, const void *d_src, void *d_dst, size_t n, uint32_t *d_bad) {
    if (dir == 0) table_serialize_kernel<F><<<blocks_for(n, 128), 128, 0, c->stream>>>((const aff_t<F> *)d_src, n, (uint32_t *)d_dst);
    else table_deserialize_kernel<F><<<blocks_for(n, 128), 128, 0, c->stream>>>((const uint32_t *)d_src, n, serialized, (aff_t<F> *)d_dst, d_bad);
    MSM_CUDA(c, cudaGetLastError());
    return MSMB200_OK;
}

static int checksum_impl(Ctx *c, const void *d_src, size_t bytes, unsigned long long *d_out) {
    checksum_kernel<<<c->sms * 4, 256, 0, c->stream>>>((const unsigned long long *)d_src, bytes / 8, d_out);
    MSM_CUDA(c, cudaGetLastError());
    return MSMB200_OK;
}

template <class F, class FC> static int resident_blocks_impl(int which) {
    if (which == 2) return 32 / coop_group_lanes<typename coop_of<FC>::type>();  // groups per warp of the cooperative kernels
    int nb = 0;
    cudaError_t e = which == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, list_sum_kernel<F, 0>, 128, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, list_sum_coop_kernel<FC, 2>, 128, 0);
    return e == cudaSuccess && nb > 0 ? nb : 2;
}

template <class F, class FC> static int point_op_impl(int op, const void *a, const void *b, const unsigned char *flags, void *out, size_t n) {
    if (op == 9) batch_to_affine_kernel<FC><<<blocks_for((n + 2) / 3, 128), 128>>>((const jac_t<FC> *)a, n, (aff_t<FC> *)out);
    else if (op == 2 || op == 3) point_op_xyzz_kernel<F><<<blocks_for(n, 128), 128>>>(op, a, b, flags, out, n);
    else if (op == 6 || op == 7)
        point_op_coop_kernel<FC><<<blocks_for(coop_group_lanes<typename coop_of<FC>::type>() * n, 128), 128>>>(op, a, b, out, n);
    else point_op_misc_kernel<FC><<<blocks_for(n, 128), 128>>>(op, a, b, out, n);
    return cudaGetLastError() == cudaSuccess ? 0 : MSMB200_ECUDA;
}

}  // namespace msmb200

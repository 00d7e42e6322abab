// C ABI of libmsm_b200.so (include/msm_b200.h): context management, table builds, the four MSM methods,
// multi-GPU partial/sum entry points, blst-named shims and parity-test hooks. Host code only orchestrates;
// every computation is a CUDA kernel in msm_kernels.cuh. No CPU fallback exists: CUDA errors are returned.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <chrono>
#include <mutex>
#include "engine.hpp"

using namespace msmb200;

namespace msmb200 {

static thread_local std::string g_create_err;

int ctx_fail(Ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_err = msg;
    return code;
}
int ensure(Ctx *c, DevBuf &b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { ctx_fail(c, MSMB200_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); b.p = nullptr; return MSMB200_ECUDA; }
    b.bytes = want;
    return 0;
}

std::vector<int> build_chunk_first(const int *values, size_t count, uint32_t vspan, uint32_t *nchunks_out) {
    uint32_t maxv = count > 1 ? (uint32_t)values[count - 1] : 0;
    uint32_t nchunks = (maxv + vspan - 1) / vspan;
    if (nchunks == 0) nchunks = 1;
    std::vector<int> cf(nchunks + 1);
    size_t l = 1;
    for (uint32_t c = 0; c <= nchunks; c++) {
        while (l < count && (uint64_t)values[l] <= (uint64_t)c * vspan) l++;
        cf[c] = (int)l;
    }
    cf[nchunks] = (int)count;
    *nchunks_out = nchunks;
    return cf;
}
uint32_t pick_vspan_host(size_t max_value, uint32_t nwindows) {
    // ~32 K chunks in total (measured optimum on B200: G1 n=2^21 -> 64, G2 n=2^18 -> 16), 8 <= v <= 64, power of two
    uint32_t v = 8;
    while (v < 64 && ((max_value + 1) * nwindows + v - 1) / v > 32768 + 1024) v <<= 1;
    return v;
}

// ---- digit-splitting reduction plan (see list_sum_kernel) ----
// Quads per list for the cooperative list sums. `resident` 128-thread blocks fit on an SM, i.e. `resident` warps per
// sub-partition per wave; a sub-partition with one warp runs at about half the pipe rate, so up to 2 warps are free;
// a partially filled last wave costs a whole chain again.
static uint32_t pick_quads(size_t total_lists, double avg_len, int subparts, int resident, int gpw) {
    uint32_t best = (uint32_t)gpw;
    double best_cost = 1e300;
    for (uint32_t tq = (uint32_t)gpw; tq >= 1; tq >>= 1) {
        const double warps = std::ceil((double)total_lists * tq / (double)gpw), per_wave = (double)subparts * resident;
        const double full = std::floor(warps / per_wave), rem = warps - full * per_wave;
        double load = full * std::max(2.0, (double)resident);
        if (rem > 0) load += std::max(2.0, std::ceil(rem / subparts));
        const double chain = std::max(1.0, std::ceil(avg_len / tq) - 1.0 + std::log2((double)tq));
        const double cost = load * chain;
        if (cost < best_cost) { best_cost = cost; best = tq; }
    }
    return best;
}
static int upload_list_plan(Ctx *c, ListPlan &lp, const std::vector<uint32_t> &start, const std::vector<uint32_t> &idx) {
    lp.nlists = (uint32_t)start.size() - 1;
    lp.members = idx.size();
    if (ensure(c, lp.start, start.size() * 4) || ensure(c, lp.idx, (idx.size() + 1) * 4)) return MSMB200_ECUDA;
    MSM_CUDA(c, cudaMemcpy(lp.start.p, start.data(), start.size() * 4, cudaMemcpyHostToDevice));
    if (!idx.empty()) MSM_CUDA(c, cudaMemcpy(lp.idx.p, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
    return MSMB200_OK;
}
void free_reduce_plan(ReducePlan &plan) {
    for (ListPlan *lp : {&plan.s1, &plan.s1b, &plan.s2a, &plan.s2b})
        for (DevBuf *b : {&lp->start, &lp->idx})
            if (b->p) { cudaFree(b->p); b->p = nullptr; b->bytes = 0; }
    plan.valid = false;
}
// Cuts every list of (start, idx) into slices of at most `slice` members. Returns the sliced lists (same idx order) in
// (start_s) and, per original list, the contiguous run of its slices as a second-level plan (start_g, idx_g).
static void slice_lists(const std::vector<uint32_t> &start, uint32_t slice, std::vector<uint32_t> &start_s, std::vector<uint32_t> &start_g,
                        std::vector<uint32_t> &idx_g) {
    start_s.assign(1, 0);
    start_g.assign(1, 0);
    idx_g.clear();
    for (size_t i = 0; i + 1 < start.size(); i++) {
        for (uint32_t e = start[i]; e < start[i + 1]; e += slice) {
            idx_g.push_back((uint32_t)start_s.size() - 1);
            start_s.push_back(std::min(start[i + 1], e + slice));
        }
        start_g.push_back((uint32_t)idx_g.size());
    }
}
// Host form of the plan: the four list stages as (start, idx) arrays plus the team sizes. Pure host code (no CUDA), so
// the CPU test-suite can evaluate it on integers (msmb200_host_reduce_plan_eval).
struct HostListPlan { std::vector<uint32_t> start, idx; uint32_t tl = 1; };
struct HostReducePlan { HostListPlan s1, s1b, s2a, s2b; uint32_t c_lo = 0, nbits_w = 0, slice1 = 0; };
static void compute_reduce_plan(HostReducePlan &H, const int *values, size_t nbw, uint32_t nwindows, int subparts, int res1, int resq, int gpw,
                                bool s1_coop) {
    const uint32_t maxv = nbw > 1 ? (values ? (uint32_t)values[nbw - 1] : (uint32_t)(nbw - 1)) : 1u;
    uint32_t V = 0;
    while (V < 32 && (maxv >> V) != 0) V++;
    const uint32_t c_lo = std::max<uint32_t>(1, (V + 1) / 2);
    const uint32_t nlo = (1u << c_lo) - 1, nhi = maxv >> c_lo;
    uint32_t chi = 0;
    while ((nhi >> chi) != 0) chi++;
    // digit lists: list (lo digit d) = d - 1, list (hi digit d) = nlo + d - 1; members are local bucket indices
    const uint32_t nl1 = nlo + nhi;
    std::vector<uint32_t> start(nl1 + 1, 0);
    std::vector<uint32_t> &idx = H.s1.idx;
    for (size_t l = 1; l < nbw; l++) {
        uint32_t v = values ? (uint32_t)values[l] : (uint32_t)l;
        if (v & nlo) start[(v & nlo) - 1 + 1]++;
        if (v >> c_lo) start[nlo + (v >> c_lo) - 1 + 1]++;
    }
    for (uint32_t i = 0; i < nl1; i++) start[i + 1] += start[i];
    idx.assign(start[nl1], 0);
    {
        std::vector<uint32_t> cur(start.begin(), start.end() - 1);
        for (size_t l = 1; l < nbw; l++) {
            uint32_t v = values ? (uint32_t)values[l] : (uint32_t)l;
            if (v & nlo) idx[cur[(v & nlo) - 1]++] = (uint32_t)l;
            if (v >> c_lo) idx[cur[nlo + (v >> c_lo) - 1]++] = (uint32_t)l;
        }
    }
    // stage 1: one LANE per slice of a digit list (perfect balance whatever the list lengths), slices sized so that the
    // whole stage is exactly ONE wave of resident blocks (a partially filled second wave would cost a full chain again);
    // stage 1b: the slices of each digit list, summed by lane groups
    const double lanes1 = (s1_coop ? (double)gpw * resq : 32.0 * res1) * subparts;
    uint32_t slice1 = (uint32_t)std::max(2.0, std::ceil((double)idx.size() * nwindows / lanes1));
    auto count_slices = [&](uint32_t sl) {
        size_t n = 0;
        for (uint32_t i = 0; i < nl1; i++) n += (start[i + 1] - start[i] + sl - 1) / sl;
        return n * nwindows;
    };
    while (slice1 < (1u << 20) && (double)count_slices(slice1) > lanes1) slice1 += std::max(1u, slice1 / 16);  // ragged last slices
    // Fp2 (groups of 8 lanes: gpw == 4): a cooperative addition of stage 1b costs about half of a whole stage-1 step, so
    // short slices are made half as long again — fewer slices for stage 1b to fold. Measured (gpurun_out/r2as_slice_scale.log):
    // G2 reduce 0.85 -> 0.71 ms at n=2^16, 0.97 -> 0.88 ms at 2^18, 0.38 -> 0.34 ms at 2^15; at n=2^21 (slices of 31) and for
    // Fp the one-wave length is the optimum.
    if (gpw == 4 && slice1 <= 16) slice1 += slice1 / 2;
    if (const char *e = getenv("MSMB200_SLICE1")) slice1 = (uint32_t)std::max(1, atoi(e));
    H.slice1 = slice1;
    slice_lists(start, slice1, H.s1.start, H.s1b.start, H.s1b.idx);
    H.s1.tl = 1;
    H.s1b.tl = pick_quads((size_t)nl1 * nwindows, nl1 ? (double)H.s1b.idx.size() / nl1 : 1.0, subparts, resq, gpw);
    // stage 2a: bit k of the lo (k < c_lo) or hi (k >= c_lo) digit value, cut into slices of SLICE members; 2b: bit lists
    const uint32_t SLICE = 32, nbits = c_lo + chi;
    std::vector<uint32_t> start_a(1, 0);
    for (uint32_t k = 0; k < nbits; k++) {
        const bool hi = k >= c_lo;
        const uint32_t bit = hi ? k - c_lo : k, vmax = hi ? nhi : nlo, base = hi ? nlo : 0;
        for (uint32_t v = 1; v <= vmax; v++)
            if ((v >> bit) & 1) H.s2a.idx.push_back(base + v - 1);
        start_a.push_back((uint32_t)H.s2a.idx.size());
    }
    slice_lists(start_a, SLICE, H.s2a.start, H.s2b.start, H.s2b.idx);
    const size_t nl2a = H.s2a.start.size() - 1;
    H.s2a.tl = pick_quads(nl2a * nwindows, nl2a ? (double)H.s2a.idx.size() / nl2a : 1.0, subparts, resq, gpw);
    H.s2b.tl = pick_quads((size_t)nbits * nwindows, nbits ? (double)H.s2b.idx.size() / nbits : 1.0, subparts, resq, gpw);
    H.c_lo = c_lo;
    H.nbits_w = nbits;
}
int build_reduce_plan(Ctx *c, ReducePlan &plan, const int *values, size_t nbw, uint32_t nwindows) {
    const int sms = c->sms;
    plan.s1_coop = false;
    HostReducePlan H;
    compute_reduce_plan(H, values, nbw, nwindows, 4 * sms, c->ops->resident_blocks(0), c->ops->resident_blocks(1), c->ops->resident_blocks(2),
                        plan.s1_coop);
    const HostListPlan *hs[4] = {&H.s1, &H.s1b, &H.s2a, &H.s2b};
    ListPlan *ds[4] = {&plan.s1, &plan.s1b, &plan.s2a, &plan.s2b};
    for (int k = 0; k < 4; k++) {
        int rc = upload_list_plan(c, *ds[k], hs[k]->start, hs[k]->idx);
        if (rc) return rc;
        ds[k]->tl = hs[k]->tl;
    }
    plan.c_lo = H.c_lo;
    plan.nbits_w = H.nbits_w;
    plan.key_nbw = nbw;
    plan.valid = true;
    return MSMB200_OK;
}

static int ctx_init_common(Ctx *c, int group, int device) {
    c->group = group;
    c->device = device;
    c->ops = group == 1 ? group_ops_g1() : group_ops_g2();
    MSM_CUDA(c, cudaSetDevice(device));
    MSM_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    for (auto &e : c->ev) MSM_CUDA(c, cudaEventCreate(&e));
    MSM_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (auto &e : c->ev_chunk) MSM_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    MSM_CUDA(c, cudaMallocHost(&c->h_result, 512));
    MSM_CUDA(c, cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device));
    if (const char *e = getenv("MSMB200_BA_BATCH")) c->ba_batch_fixed = atoi(e);
    if (const char *e = getenv("MSMB200_ITEM_LEN")) c->item_len_fixed = std::max(0, atoi(e));
    if (const char *e = getenv("MSMB200_ACCUM")) c->accum_env = atoi(e);
    if (const char *e = getenv("MSMB200_REDUCE")) c->reduce_env = atoi(e);
    if (const char *e = getenv("MSMB200_BA_BATCH_MAX")) c->ba_batch_max = std::max(1, atoi(e));
    if (const char *e = getenv("MSMB200_VSPAN")) c->vspan_env = std::max(0, atoi(e));
    c->no_overlap = getenv("MSMB200_NO_OVERLAP") != nullptr;
    c->table_stride = (uint32_t)(getenv("MSMB200_PACKED_TABLES") ? c->ops->aff_bytes : (c->ops->aff_bytes + 127) & ~(size_t)127);
    return MSMB200_OK;
}

// releases everything the context owns; the memory of *c itself belongs to whoever allocated it
static void ctx_release(Ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf *bufs[] = {&c->scalars, &c->keys, &c->vals, &c->ranks, &c->sorted, &c->count, &c->packed, &c->scanned, &c->tile_sums, &c->seg_start,
                      &c->item_start, &c->cursor, &c->item_begin, &c->item_cnt, &c->order, &c->len_hist, &c->len_start, &c->len_cursor,
                      &c->partial, &c->chunk_a, &c->chunk_b, &c->result, &c->flat, &c->signs, &c->pidx, &c->heavy, &c->light, &c->medium,
                      &c->ba_totals, &c->ba_tile_sums, &c->ba_bases, &c->ba_adesc, &c->ba_cdesc, &c->ba_heavy, &c->pts_a, &c->pts_b, &c->ba_scratch,
                      &c->bucket_sum, &c->iota, &c->ba_counters, &c->ba_sm_arrivals, &c->maxcount, &c->red_a, &c->red_b, &c->red_c, &c->red_d};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    free_reduce_plan(c->plan_ches); free_reduce_plan(c->plan_bgmw); free_reduce_plan(c->plan_pip);
    void *ptrs[] = {c->d_bucket_vals, c->d_v2i, c->d_dtab, c->d_chunk_first, c->d_points, c->d_table_ches, c->d_table_bgmw};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->h_totals) cudaFreeHost(c->h_totals);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_chunk) if (e) cudaEventDestroy(e);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = nullptr;
}

// lazily created context used by the context-free blst-named shims
static std::mutex g_shim_mu;
// the blst functions are re-entrant; the shims share one lazily created context per group, so calls of the
// same group are serialised (the GPU runs one MSM at a time anyway)
static std::mutex g_shim_call_mu[3];
// a shim call = the lock plus a wall-clock stamp (msmb200_blst_last_call_ms: what one drop-in call cost, host to host)
static double g_shim_last_ms[3] = {0, 0, 0};
struct ShimCall {
    std::lock_guard<std::mutex> lock;
    int group;
    std::chrono::steady_clock::time_point t0;
    explicit ShimCall(int g) : lock(g_shim_call_mu[g]), group(g), t0(std::chrono::steady_clock::now()) {}
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
    void lap(const char *what) const {   // MSMB200_SHIM_TRACE=1: where a drop-in call spends its time (stderr)
        static const bool on = getenv("MSMB200_SHIM_TRACE") != nullptr;
        if (on) fprintf(stderr, "msm_b200 shim[g%d] %-28s %9.3f ms\n", group, what, ms());
    }
    ~ShimCall() { g_shim_last_ms[group] = ms(); lap("return"); }
};
static Ctx *g_shim[3] = {nullptr, nullptr, nullptr};
// may_fail: return nullptr instead of aborting (entry points that have an error channel)
static Ctx *shim_ctx(int group, bool may_fail = false) {
    std::lock_guard<std::mutex> lk(g_shim_mu);
    if (!g_shim[group]) {
        Ctx *c = new Ctx();
        const char *dev = getenv("MSMB200_DEVICE");
        if (ctx_init_common(c, group, dev ? atoi(dev) : 0) != MSMB200_OK) {
            if (may_fail) { g_create_err = c->err; delete c; return nullptr; }
            fprintf(stderr, "msm_b200: cannot initialise CUDA for blst shim: %s\n", c->err.c_str());
            abort();  // the blst signatures have no error channel; never fall back to the CPU
        }
        g_shim[group] = c;
    }
    return g_shim[group];
}

}  // namespace msmb200

struct msmb200_ctx { Ctx c; };
static inline Ctx *C(msmb200_ctx *x) { return &x->c; }

extern "C" {

int msmb200_config_lookup(const char *name, msmb200_config *out) {
    const msmb200_config *c = name ? find_config(name) : nullptr;
    if (!c || !out) return MSMB200_EINVAL;
    *out = *c;
    return MSMB200_OK;
}

long msmb200_host_bucket_set(int e, int a, int *out, long cap) {
    if (e < 4 || e > 24 || a < 0) return MSMB200_EINVAL;
    std::vector<int> B = build_bucket_set(1 << e, a);
    if (out && cap >= (long)B.size()) memcpy(out, B.data(), B.size() * sizeof(int));
    return (long)B.size();
}
int msmb200_host_bucket_set_check(long q, long a, long out[5]) {
    if (q < 16 || q > ((long)1 << 26) || (q & 1) || a < 0 || a + 1 > q / 2 || !out) return MSMB200_EINVAL;
    const std::vector<int> B = build_bucket_set((int)q, (int)a);
    const BucketSetCheck r = check_bucket_set(B, q, a);
    out[0] = r.leading_ok && r.cover_ok ? 1 : 0;
    out[1] = r.size; out[2] = r.max_gap; out[3] = r.leading_ok ? 1 : 0; out[4] = r.first_uncovered;
    return MSMB200_OK;
}
int msmb200_host_digit_table(int e, int a, int *out_triples) {
    if (e < 4 || e > 24 || a < 0 || !out_triples) return MSMB200_EINVAL;
    int q = 1 << e;
    std::vector<int> B = build_bucket_set(q, a);
    std::vector<uint32_t> T = build_digit_table(q, B);
    for (int d = 0; d <= q; d++) {
        if (T[d] == DT_INVALID) return MSMB200_EINVAL;
        out_triples[3 * d + 0] = (int)((T[d] >> DT_M_SHIFT) & 3u) + 1;
        out_triples[3 * d + 1] = B[T[d] & DT_IDX_MASK];
        out_triples[3 * d + 2] = (int)((T[d] >> DT_A_SHIFT) & 1u);
    }
    return MSMB200_OK;
}
size_t msmb200_pippenger_window_size(size_t npoints) { return pippenger_window_size(npoints); }

// Evaluates the digit-splitting reduction plan on 64-bit integers instead of points (host only): x[l] plays the bucket
// sum of local bucket l, additions are mod 2^64, doubling is a shift; the result must equal sum_l value(l) * x[l].
// Exercises exactly the arrays the GPU kernels walk (stage 1 slices, 1b, 2a, 2b, Horner over bit positions).
int msmb200_host_reduce_plan_eval(const int *values, size_t nbw, int resident_stage1, int resident_coop, int groups_per_warp, const uint64_t *x,
                                  uint64_t *out, uint32_t *out_info) {
    if (nbw < 2 || !x || !out || resident_stage1 < 1 || resident_coop < 1 || groups_per_warp < 1) return MSMB200_EINVAL;
    HostReducePlan H;
    compute_reduce_plan(H, values, nbw, 1, 4 * 148 /* the B200 shape the CPU tests evaluate plans for */, resident_stage1, resident_coop, groups_per_warp, false);
    auto run = [](const HostListPlan &lp, const std::vector<uint64_t> &in) {
        std::vector<uint64_t> o(lp.start.size() - 1, 0);
        for (size_t i = 0; i + 1 < lp.start.size(); i++)
            for (uint32_t e = lp.start[i]; e < lp.start[i + 1]; e++) o[i] += in[lp.idx[e]];
        return o;
    };
    std::vector<uint64_t> in(x, x + nbw);
    in[0] = 0;
    std::vector<uint64_t> a = run(H.s1, in), b = run(H.s1b, a), c2 = run(H.s2a, b), d = run(H.s2b, c2);
    uint64_t acc = 0;
    for (int k = (int)H.nbits_w - 1; k >= 0; k--) acc = (acc << 1) + d[k];
    *out = acc;
    if (out_info) {
        out_info[0] = H.c_lo; out_info[1] = H.nbits_w; out_info[2] = H.slice1; out_info[3] = (uint32_t)(H.s1.start.size() - 1);
        out_info[4] = (uint32_t)(H.s1b.start.size() - 1); out_info[5] = H.s1b.tl; out_info[6] = (uint32_t)(H.s2a.start.size() - 1); out_info[7] = H.s2a.tl;
    }
    return MSMB200_OK;
}

const char *msmb200_last_error(const msmb200_ctx *ctx) { return ctx ? ctx->c.err.c_str() : g_create_err.c_str(); }

int msmb200_ctx_create(msmb200_ctx **out, int group, const msmb200_config *cfg, size_t npoints, int device) {
    if (!out || !cfg || (group != 1 && group != 2) || npoints == 0) return ctx_fail(nullptr, MSMB200_EINVAL, "bad arguments");
    if (cfg->e < 4 || cfg->e > 22 || cfg->h < 1 || cfg->h * cfg->e < 255 || cfg->h > 64 || cfg->e_bgmw < 2 || cfg->e_bgmw > 22 ||
        cfg->h_bgmw * cfg->e_bgmw < 255 || cfg->h_bgmw > 128 || cfg->d < 1 || cfg->d > 8)
        return ctx_fail(nullptr, MSMB200_EINVAL, "configuration out of range");
    if ((double)npoints * cfg->h * 3 >= 2147483648.0 || (double)npoints * cfg->h_bgmw >= 2147483648.0)
        return ctx_fail(nullptr, MSMB200_EINVAL, "table index would exceed 31 bits; shard the points over more contexts");
    msmb200_ctx *x = new msmb200_ctx();
    Ctx *c = &x->c;
    c->cfg = *cfg;
    c->n = npoints;
    int rc = ctx_init_common(c, group, device);
    if (rc) { g_create_err = c->err; ctx_release(c); delete x; return rc; }
    // CHES parameters
    c->q = 1 << cfg->e;
    c->bucket_set = build_bucket_set(c->q, cfg->a);
    if (cfg->bsize && (size_t)cfg->bsize != c->bucket_set.size()) {
        g_create_err = "bucket set size differs from configured B_SIZE";
        ctx_release(c);
        delete x;
        return MSMB200_EINVAL;
    }
    int gap = 0;
    for (size_t i = 1; i < c->bucket_set.size(); i++) gap = std::max(gap, c->bucket_set[i] - c->bucket_set[i - 1]);
    std::vector<uint32_t> dtab = build_digit_table(c->q, c->bucket_set);
    bool covered = gap <= cfg->d;
    for (uint32_t v : dtab) covered = covered && v != DT_INVALID;
    if (!covered) { g_create_err = "bucket set does not cover all digits / max gap exceeds d"; ctx_release(c); delete x; return MSMB200_EINVAL; }
    std::vector<int> v2i(c->q / 2 + 1, 0);
    for (size_t i = 0; i < c->bucket_set.size(); i++) v2i[c->bucket_set[i]] = (int)i;
    c->pip_window = (int)pippenger_window_size(npoints);
    c->pip_tiles = 255 / c->pip_window + 1;
#define CREATE_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { g_create_err = std::string(#call) + ": " + cudaGetErrorString(e__); ctx_release(c); delete x; return MSMB200_ECUDA; } } while (0)
    CREATE_CUDA(cudaMalloc(&c->d_bucket_vals, c->bucket_set.size() * sizeof(int)));
    CREATE_CUDA(cudaMalloc(&c->d_v2i, v2i.size() * sizeof(int)));
    CREATE_CUDA(cudaMalloc(&c->d_dtab, dtab.size() * sizeof(uint32_t)));
    c->red_vspan = c->vspan_env ? (uint32_t)c->vspan_env : pick_vspan_host((size_t)c->bucket_set.back(), 1);
    std::vector<int> cf = build_chunk_first(c->bucket_set.data(), c->bucket_set.size(), c->red_vspan, &c->red_nchunks);
    c->h_chunk_first = cf;
    CREATE_CUDA(cudaMalloc(&c->d_chunk_first, cf.size() * sizeof(int)));
    CREATE_CUDA(cudaMemcpy(c->d_chunk_first, cf.data(), cf.size() * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMalloc(&c->d_points, npoints * c->ops->aff_bytes));
    CREATE_CUDA(cudaMemcpy(c->d_bucket_vals, c->bucket_set.data(), c->bucket_set.size() * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(c->d_v2i, v2i.data(), v2i.size() * sizeof(int), cudaMemcpyHostToDevice));
    CREATE_CUDA(cudaMemcpy(c->d_dtab, dtab.data(), dtab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
#undef CREATE_CUDA
    rc = build_reduce_plan(c, c->plan_ches, c->bucket_set.data(), c->bucket_set.size(), 1);
    if (rc) { g_create_err = c->err; ctx_release(c); delete x; return rc; }
    *out = x;
    return MSMB200_OK;
}

void msmb200_ctx_destroy(msmb200_ctx *ctx) {
    if (!ctx) return;
    ctx_release(&ctx->c);
    delete ctx;
}

int msmb200_set_stream(msmb200_ctx *ctx, void *cuda_stream) {
    if (!ctx) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    cudaSetDevice(c->device);
    if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return MSMB200_OK;
}

int msmb200_set_bucket_shard(msmb200_ctx *ctx, int rank, int world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return MSMB200_EINVAL;
    C(ctx)->shard_rank = rank;
    C(ctx)->shard_world = world;
    return MSMB200_OK;
}
int msmb200_set_accumulator(msmb200_ctx *ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return MSMB200_EINVAL;
    C(ctx)->accum_mode = mode;
    return MSMB200_OK;
}
int msmb200_set_reducer(msmb200_ctx *ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return MSMB200_EINVAL;
    C(ctx)->reduce_mode = mode;
    return MSMB200_OK;
}
int msmb200_set_tuning(msmb200_ctx *ctx, const char *key, int value) {
    if (!ctx || !key || value < 0) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    const std::string k(key);
    if (k == "ba_batch_max") c->ba_batch_max = std::max(1, value);
    else if (k == "ba_batch") c->ba_batch_fixed = value;
    else if (k == "ba_stagger") c->ba_stagger = value;
    else if (k == "item_len") c->item_len_fixed = value;
    else return ctx_fail(c, MSMB200_EINVAL, "unknown tuning key " + k);
    return MSMB200_OK;
}
int msmb200_set_points(msmb200_ctx *ctx, const void *points_affine_host) {
    if (!ctx || !points_affine_host) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    MSM_CUDA(c, cudaMemcpyAsync(c->d_points, points_affine_host, c->n * c->ops->aff_bytes, cudaMemcpyHostToDevice, c->stream));
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_points = true;
    c->have_ches = c->have_bgmw = false;
    return MSMB200_OK;
}
int msmb200_generate_fix_points(msmb200_ctx *ctx, size_t first) {
    if (!ctx) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    c->have_ches = c->have_bgmw = false;
    return c->ops->generate_fix_points(c, first);
}
int msmb200_table_build_ches(msmb200_ctx *ctx) {
    if (!ctx) return MSMB200_EINVAL;
    MSM_CUDA(C(ctx), cudaSetDevice(C(ctx)->device));
    return C(ctx)->ops->table_build(C(ctx), 0);
}
int msmb200_table_build_bgmw95(msmb200_ctx *ctx) {
    if (!ctx) return MSMB200_EINVAL;
    MSM_CUDA(C(ctx), cudaSetDevice(C(ctx)->device));
    return C(ctx)->ops->table_build(C(ctx), 1);
}
int msmb200_download(msmb200_ctx *ctx, int which, size_t first, size_t count, void *out_host) {
    if (!ctx || !out_host) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    const void *src = nullptr;
    size_t total = 0;
    if (which == 0) { src = c->have_points ? c->d_points : nullptr; total = c->n; }
    else if (which == 1) { src = c->have_ches ? c->d_table_ches : nullptr; total = c->n * (size_t)c->cfg.h * 3; }
    else if (which == 2) { src = c->have_bgmw ? c->d_table_bgmw : nullptr; total = c->n * (size_t)c->cfg.h_bgmw; }
    else return ctx_fail(c, MSMB200_EINVAL, "which must be 0, 1 or 2");
    if (!src) return ctx_fail(c, MSMB200_ESTATE, "requested array not built");
    if (first + count > total) return ctx_fail(c, MSMB200_EINVAL, "range out of bounds");
    MSM_CUDA(c, cudaSetDevice(c->device));
    const size_t ab = c->ops->aff_bytes, stride = which == 0 ? ab : c->table_stride;   // out: always the reference's packed layout
    if (count) MSM_CUDA(c, cudaMemcpy2DAsync(out_host, ab, (const char *)src + first * stride, stride, ab, count, cudaMemcpyDeviceToHost, c->stream));
    MSM_CUDA(c, cudaStreamSynchronize(c->stream));
    return MSMB200_OK;
}
// ---- table persistence ----
struct TableFileHeader {
    char magic[8];       // "MSMB200T"
    uint32_t version, group, format, which;
    int32_t cfg[8];      // n_exp, e, h, a, d, bsize, e_bgmw, h_bgmw
    uint64_t npoints, entries;
    uint64_t points_digest;  // version 2: checksum of the fixed points the array belongs to (a table of OTHER points must not load)
};
// checksum of the context's fixed points (device computation, 8 bytes read back)
static int points_digest(Ctx *c, uint64_t *out) {
    unsigned long long *d = nullptr;
    MSM_CUDA(c, cudaMalloc((void **)&d, 8));
    cudaMemsetAsync(d, 0, 8, c->stream);
    int rc = c->ops->checksum(c, c->d_points, c->n * c->ops->aff_bytes, d);
    if (!rc && (cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess))
        rc = ctx_fail(c, MSMB200_ECUDA, "points digest");
    cudaFree(d);
    return rc;
}
static int table_slot(Ctx *c, int which, void ***slot, size_t *entries, bool **have) {
    if (which == 0) { *slot = &c->d_points; *entries = c->n; *have = &c->have_points; }
    else if (which == 1) { *slot = &c->d_table_ches; *entries = c->n * (size_t)c->cfg.h * 3; *have = &c->have_ches; }
    else if (which == 2) { *slot = &c->d_table_bgmw; *entries = c->n * (size_t)c->cfg.h_bgmw; *have = &c->have_bgmw; }
    else return ctx_fail(c, MSMB200_EINVAL, "which must be 0 (points), 1 (CHES table) or 2 (BGMW95 table)");
    return MSMB200_OK;
}
int msmb200_table_save(msmb200_ctx *ctx, int which, const char *path, int format) {
    if (!ctx || !path || (format != 0 && format != 1)) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    void **slot; size_t entries; bool *have;
    int rc = table_slot(c, which, &slot, &entries, &have);
    if (rc) return rc;
    if (!*have || !*slot) return ctx_fail(c, MSMB200_ESTATE, "requested array not built");
    MSM_CUDA(c, cudaSetDevice(c->device));
    FILE *f = fopen(path, "wb");
    if (!f) return ctx_fail(c, MSMB200_EINVAL, std::string("cannot open ") + path);
    TableFileHeader h{};
    memcpy(h.magic, "MSMB200T", 8);
    h.version = 2; h.group = (uint32_t)c->group; h.format = (uint32_t)format; h.which = (uint32_t)which;
    if (c->have_points && points_digest(c, &h.points_digest)) { fclose(f); return MSMB200_ECUDA; }
    const int32_t cf[8] = {c->cfg.n_exp, c->cfg.e, c->cfg.h, c->cfg.a, c->cfg.d, c->cfg.bsize, c->cfg.e_bgmw, c->cfg.h_bgmw};
    memcpy(h.cfg, cf, sizeof(cf));
    h.npoints = c->n; h.entries = entries;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const size_t ab = c->ops->aff_bytes, chunk = (size_t)1 << 20;
    std::vector<unsigned char> host(std::min(entries, chunk) * ab);
    const size_t stride = which == 0 ? ab : c->table_stride;   // the context's own tables are padded to whole lines; files are packed
    void *d_stage = nullptr, *d_packed = nullptr;
    if ((format == 1 && cudaMalloc(&d_stage, std::min(entries, chunk) * ab) != cudaSuccess) ||
        (stride != ab && cudaMalloc(&d_packed, std::min(entries, chunk) * ab) != cudaSuccess)) {
        if (d_stage) cudaFree(d_stage);
        fclose(f);
        return ctx_fail(c, MSMB200_ECUDA, "cudaMalloc");
    }
    for (size_t off = 0; ok && off < entries; off += chunk) {
        const size_t cnt = std::min(chunk, entries - off);
        const void *src = (const char *)*slot + off * stride;
        if (stride != ab) {
            if (cudaMemcpy2DAsync(d_packed, ab, src, stride, ab, cnt, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { ok = false; break; }
            src = d_packed;
        }
        if (format == 1) {
            if (c->ops->table_io(c, 0, 1, src, d_stage, cnt, nullptr)) { ok = false; break; }
            src = d_stage;
        }
        ok = cudaMemcpyAsync(host.data(), src, cnt * ab, cudaMemcpyDeviceToHost, c->stream) == cudaSuccess && cudaStreamSynchronize(c->stream) == cudaSuccess &&
             fwrite(host.data(), ab, cnt, f) == cnt;
    }
    if (d_stage) cudaFree(d_stage);
    if (d_packed) cudaFree(d_packed);
    ok = (fclose(f) == 0) && ok;
    return ok ? MSMB200_OK : ctx_fail(c, MSMB200_ECUDA, std::string("writing ") + path + " failed");
}
int msmb200_table_load(msmb200_ctx *ctx, int which, const char *path) {
    if (!ctx || !path) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    void **slot; size_t entries; bool *have;
    int rc = table_slot(c, which, &slot, &entries, &have);
    if (rc) return rc;
    MSM_CUDA(c, cudaSetDevice(c->device));
    FILE *f = fopen(path, "rb");
    if (!f) return ctx_fail(c, MSMB200_EINVAL, std::string("cannot open ") + path);
    TableFileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "MSMB200T", 8) != 0 || h.version != 2 || h.format > 1) { fclose(f); return ctx_fail(c, MSMB200_EINVAL, "not a table file (or written by an older version)"); }
    // the table depends on the group, the points (n) and the radix / length of its method; the points on group and n only
    bool match = h.group == (uint32_t)c->group && h.which == (uint32_t)which && h.npoints == c->n && h.entries == entries;
    if (which == 1) match = match && h.cfg[1] == c->cfg.e && h.cfg[2] == c->cfg.h;
    if (which == 2) match = match && h.cfg[6] == c->cfg.e_bgmw && h.cfg[7] == c->cfg.h_bgmw;
    if (!match) { fclose(f); return ctx_fail(c, MSMB200_EINVAL, "table file was written for another group / size / configuration"); }
    if (which != 0) {  // a precomputation table belongs to the points it was built from: they must be here and be the same
        uint64_t dig = 0;
        if (!c->have_points) { fclose(f); return ctx_fail(c, MSMB200_ESTATE, "load or generate the fixed points before their table"); }
        if (points_digest(c, &dig)) { fclose(f); return MSMB200_ECUDA; }
        if (dig != h.points_digest) { fclose(f); return ctx_fail(c, MSMB200_EINVAL, "table file was built from other fixed points"); }
    }
    const size_t ab = c->ops->aff_bytes, chunk = (size_t)1 << 20, stride = which == 0 ? ab : c->table_stride;
    if (!*slot) {
        if (cudaMalloc(slot, entries * stride) != cudaSuccess) { fclose(f); *slot = nullptr; return ctx_fail(c, MSMB200_ECUDA, "cudaMalloc"); }
        if (stride != ab) cudaMemsetAsync(*slot, 0, entries * stride, c->stream);
    }
    *have = false;
    if (which == 0) c->have_ches = c->have_bgmw = false;
    std::vector<unsigned char> host(std::min(entries, chunk) * ab);
    void *d_stage = nullptr, *d_packed = nullptr;
    uint32_t *d_bad = nullptr, bad = 0;
    if (cudaMalloc(&d_stage, std::min(entries, chunk) * ab) != cudaSuccess || cudaMalloc((void **)&d_bad, 4) != cudaSuccess ||
        (stride != ab && cudaMalloc(&d_packed, std::min(entries, chunk) * ab) != cudaSuccess)) {
        if (d_stage) cudaFree(d_stage);
        if (d_bad) cudaFree(d_bad);
        fclose(f);
        return ctx_fail(c, MSMB200_ECUDA, "cudaMalloc");
    }
    cudaMemsetAsync(d_bad, 0, 4, c->stream);
    bool ok = true;
    for (size_t off = 0; ok && off < entries; off += chunk) {
        const size_t cnt = std::min(chunk, entries - off);
        ok = fread(host.data(), ab, cnt, f) == cnt &&
             cudaMemcpyAsync(d_stage, host.data(), cnt * ab, cudaMemcpyHostToDevice, c->stream) == cudaSuccess &&
             c->ops->table_io(c, 1, (int)h.format, d_stage, stride != ab ? d_packed : (void *)((char *)*slot + off * ab), cnt, d_bad) == MSMB200_OK &&
             (stride == ab || cudaMemcpy2DAsync((char *)*slot + off * stride, stride, d_packed, ab, ab, cnt, cudaMemcpyDeviceToDevice, c->stream) == cudaSuccess) &&
             cudaStreamSynchronize(c->stream) == cudaSuccess;  // the host buffer is reused
    }
    fclose(f);
    if (ok) ok = cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_stage); cudaFree(d_bad);
    if (d_packed) cudaFree(d_packed);
    if (!ok) return ctx_fail(c, MSMB200_ECUDA, std::string("reading ") + path + " failed");
    if (bad) return ctx_fail(c, MSMB200_EINVAL, std::to_string(bad) + " entries are not valid curve points (range, flags or y^2 = x^3 + B)");
    *have = true;
    return MSMB200_OK;
}

long msmb200_bucket_set(msmb200_ctx *ctx, int *out, long cap) {
    if (!ctx) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    if (out && cap >= (long)c->bucket_set.size()) memcpy(out, c->bucket_set.data(), c->bucket_set.size() * sizeof(int));
    return (long)c->bucket_set.size();
}

int msmb200_msm_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_affine_host) {
    if (!ctx || !scalars_dev || !out_affine_host) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    int rc = c->ops->msm(c, method, scalars_dev, nullptr, true);
    if (rc) return rc;
    memcpy(out_affine_host, c->h_result, c->ops->aff_bytes);
    return MSMB200_OK;
}
int msmb200_msm(msmb200_ctx *ctx, int method, const void *scalars_host, void *out_affine_host) {
    if (!ctx || !scalars_host || !out_affine_host) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    if (ensure(c, c->scalars, c->n * 32)) return MSMB200_ECUDA;
    if (method == MSMB200_CHES && !c->no_overlap) {
        // chunked upload overlapped with the digit decomposition (msm_impl)
        c->h_scalars_pending = scalars_host;
        int rc = c->ops->msm(c, method, c->scalars.p, nullptr, true);
        c->h_scalars_pending = nullptr;
        if (rc) return rc;
        memcpy(out_affine_host, c->h_result, c->ops->aff_bytes);
        return MSMB200_OK;
    }
    MSM_CUDA(c, cudaMemcpyAsync(c->scalars.p, scalars_host, c->n * 32, cudaMemcpyHostToDevice, c->stream));
    return msmb200_msm_device(ctx, method, c->scalars.p, out_affine_host);
}
int msmb200_msm_partial_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_jacobian_dev) {
    if (!ctx || !scalars_dev || !out_jacobian_dev) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    return c->ops->msm(c, method, scalars_dev, out_jacobian_dev, false);
}
// layout of the per-bit sums of a method under this context: out[0] = windows, out[1] = bit positions per window,
// out[2] = doublings between windows; builds the (cached) dense reduction plan when the method needs one
int msmb200_msm_bits_layout(msmb200_ctx *ctx, int method, uint32_t out[3]) {
    if (!ctx || !out) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    if (c->reduce_mode == 1 || c->reduce_env == 1 || c->shard_world > 1) return ctx_fail(c, MSMB200_ESTATE, "per-bit partial sums need the digit-splitting reducer");
    const ReducePlan *plan = nullptr;
    uint32_t nw = 1, wbits = 0;
    if (method == MSMB200_CHES || method == MSMB200_CHES_INTEGRAL) plan = &c->plan_ches;
    else if (method == MSMB200_BGMW95) {
        const size_t nbw = ((size_t)1 << (c->cfg.e_bgmw - 1)) + 1;
        if (!c->plan_bgmw.valid || c->plan_bgmw.key_nbw != nbw || c->plan_bgmw_windows != 1) {
            if (build_reduce_plan(c, c->plan_bgmw, nullptr, nbw, 1) != MSMB200_OK) return MSMB200_ECUDA;
            c->plan_bgmw_windows = 1;
        }
        plan = &c->plan_bgmw;
    } else if (method == MSMB200_PIPPENGER) {
        nw = (uint32_t)c->pip_tiles; wbits = (uint32_t)c->pip_window;
        const size_t nbw = ((size_t)1 << (c->pip_window - 1)) + 1;
        if (!c->plan_pip.valid || c->plan_pip.key_nbw != nbw || c->plan_pip_windows != nw) {
            if (build_reduce_plan(c, c->plan_pip, nullptr, nbw, nw) != MSMB200_OK) return MSMB200_ECUDA;
            c->plan_pip_windows = nw;
        }
        plan = &c->plan_pip;
    } else return ctx_fail(c, MSMB200_EINVAL, "unknown method");
    if (!plan->valid || (nw > 1 && plan->nbits_w > wbits)) return ctx_fail(c, MSMB200_ESTATE, "per-bit partial sums unavailable for this layout");
    out[0] = nw; out[1] = plan->nbits_w; out[2] = wbits;
    return MSMB200_OK;
}
int msmb200_msm_bits_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_bits_dev) {
    if (!ctx || !scalars_dev || !out_bits_dev) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    c->bits_out = out_bits_dev;
    int rc = c->ops->msm(c, method, scalars_dev, nullptr, false);
    c->bits_out = nullptr;
    return rc;
}
int msmb200_combine_bits_device(msmb200_ctx *ctx, const void *gathered_dev, int world, const uint32_t layout[3], void *out_affine_host) {
    if (!ctx || !gathered_dev || world < 1 || !layout || !out_affine_host || layout[0] == 0 || layout[1] == 0) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    int rc = c->ops->combine_bits(c, gathered_dev, world, layout[0], layout[1], layout[2]);
    if (rc) return rc;
    memcpy(out_affine_host, c->h_result, c->ops->aff_bytes);
    return MSMB200_OK;
}
int msmb200_sum_partials_device(msmb200_ctx *ctx, const void *partials_dev, int count, void *out_affine_host) {
    if (!ctx || !partials_dev || count < 1 || !out_affine_host) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    int rc = c->ops->sum_partials(c, partials_dev, count);
    if (rc) return rc;
    memcpy(out_affine_host, c->h_result, c->ops->aff_bytes);
    return MSMB200_OK;
}

int msmb200_last_timings(msmb200_ctx *ctx, float out_ms[6]) {
    if (!ctx || !out_ms) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    MSM_CUDA(c, cudaEventSynchronize(c->ev[5]));
    for (int k = 0; k < 5; k++) MSM_CUDA(c, cudaEventElapsedTime(&c->last_ms[k], c->ev[k], c->ev[k + 1]));
    MSM_CUDA(c, cudaEventElapsedTime(&c->last_ms[5], c->ev[0], c->ev[5]));
    memcpy(out_ms, c->last_ms, sizeof(c->last_ms));
    return MSMB200_OK;
}
int msmb200_last_launches(msmb200_ctx *ctx) { return ctx ? C(ctx)->launches : MSMB200_EINVAL; }
int msmb200_last_accumulator(msmb200_ctx *ctx) { return ctx ? C(ctx)->last_accum : MSMB200_EINVAL; }

// Serialisation is a pure byte shuffle of one 96/192-byte result plus a from-Montgomery multiplication; it is
// done on the device like everything else (point_op 6 would be overkill) — here via the field-op kernel.
int msmb200_affine_serialize(int group, const void *affine_host, unsigned char *out) {
    if ((group != 1 && group != 2) || !affine_host || !out) return MSMB200_EINVAL;
    size_t nfp = group == 1 ? 2 : 4, nbytes = nfp * 48;
    bool inf = true;
    for (size_t i = 0; i < nbytes; i++) inf = inf && ((const unsigned char *)affine_host)[i] == 0;
    if (inf) { memset(out, 0, nbytes); out[0] = 0x40; return MSMB200_OK; }
    // from_fp: multiply by 1 (non-Montgomery) on the device
    std::vector<uint64_t> one(nfp * 6, 0), res(nfp * 6);
    for (size_t i = 0; i < nfp; i++) one[6 * i] = 1;
    int rc = msmb200_test_field_op(-1, 1, 0, affine_host, one.data(), res.data(), nfp);
    if (rc) return rc;
    // G1: X | Y.  G2: X.im | X.re | Y.im | Y.re   (src/e2.c:176-192), each 48 bytes big-endian
    static const int order1[2] = {0, 1}, order2[4] = {1, 0, 3, 2};
    const int *ord = group == 1 ? order1 : order2;
    for (size_t k = 0; k < nfp; k++) {
        const uint64_t *l = &res[6 * ord[k]];
        for (int b = 0; b < 48; b++) out[48 * k + b] = (unsigned char)(l[(47 - b) / 8] >> (8 * ((47 - b) % 8)));
    }
    return MSMB200_OK;
}

int msmb200_measure_peaks(int device, double *imad_macs_per_s, double *fp_mul_per_s) {
    if (!imad_macs_per_s || !fp_mul_per_s) return MSMB200_EINVAL;
    double out[4];
    int rc = msmb200_measure_peaks_ex(device, out);
    if (rc) return rc;
    *imad_macs_per_s = out[0];
    *fp_mul_per_s = out[1];
    return MSMB200_OK;
}
int msmb200_measure_peaks_ex(int device, double out[4]) {
    if (!out) return MSMB200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess) return MSMB200_ECUDA;
    return measure_peaks(out);
}

// ---- parity-test hooks ------------------------------------------------------------------------------
static int with_device_buffers(int device, const void *a, size_t abytes, const void *b, size_t bbytes, const unsigned char *flags,
                               size_t fbytes, void *out, size_t obytes, int (*fn)(const void *, const void *, const unsigned char *, void *, void *),
                               void *arg) {
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return MSMB200_ECUDA;
    void *da = nullptr, *db = nullptr, *df = nullptr, *dout = nullptr;
    int rc = MSMB200_ECUDA;
    do {
        if (cudaMalloc(&da, abytes) != cudaSuccess) break;
        if (b && cudaMalloc(&db, bbytes) != cudaSuccess) break;
        if (flags && cudaMalloc(&df, fbytes) != cudaSuccess) break;
        if (cudaMalloc(&dout, obytes) != cudaSuccess) break;
        if (cudaMemcpy(da, a, abytes, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (b && cudaMemcpy(db, b, bbytes, cudaMemcpyHostToDevice) != cudaSuccess) break;
        if (flags && cudaMemcpy(df, flags, fbytes, cudaMemcpyHostToDevice) != cudaSuccess) break;
        rc = fn(da, db, (const unsigned char *)df, dout, arg);
        if (rc) break;
        rc = MSMB200_ECUDA;
        if (cudaDeviceSynchronize() != cudaSuccess) break;
        if (cudaMemcpy(out, dout, obytes, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        rc = MSMB200_OK;
    } while (0);
    cudaFree(da); cudaFree(db); cudaFree(df); cudaFree(dout);
    return rc;
}
struct FieldArg { int field, op; size_t n; };
static int field_fn(const void *a, const void *b, const unsigned char *, void *out, void *arg) {
    FieldArg *f = (FieldArg *)arg;
    return group_ops_g1()->field_op(f->field, f->op, a, b, out, f->n);
}
int msmb200_test_field_op(int device, int field, int op, const void *a, const void *b, void *out, size_t n) {
    if ((field != 1 && field != 2) || op < 0 || op > 7 || !a || !out || n == 0) return MSMB200_EINVAL;
    size_t eb = field == 1 ? 48 : 96;
    FieldArg arg{field, op, n};
    return with_device_buffers(device, a, n * eb, b, n * eb, nullptr, 0, out, n * eb, field_fn, &arg);
}
struct PointArg { int group, op; size_t n; };
static int point_fn(const void *a, const void *b, const unsigned char *flags, void *out, void *arg) {
    PointArg *p = (PointArg *)arg;
    const GroupOps *ops = p->group == 1 ? group_ops_g1() : group_ops_g2();
    return ops->point_op(p->op, a, b, flags, out, p->n);
}
int msmb200_test_point_op(int device, int group, int op, const void *a, const void *b, const unsigned char *flags, void *out, size_t n) {
    if ((group != 1 && group != 2) || op < 0 || op > 9 || !a || !out || n == 0) return MSMB200_EINVAL;
    const GroupOps *ops = group == 1 ? group_ops_g1() : group_ops_g2();
    size_t A = ops->aff_bytes, J = ops->jac_bytes, X = ops->xyzz_bytes;
    size_t ab[10] = {J, J, X, X, X, J, X, X, X, J}, bb[10] = {J, 0, A, X, 0, 0, X, 0, 0, 0}, ob[10] = {J, J, X, X, J, A, X, X, X, A};
    if (bb[op] && !b) return MSMB200_EINVAL;
    PointArg arg{group, op, n};
    return with_device_buffers(device, a, n * ab[op], bb[op] ? b : nullptr, n * bb[op], flags, n, out, n * ob[op], point_fn, &arg);
}
int msmb200_test_digits(msmb200_ctx *ctx, int kind, const void *scalars_host, size_t n, uint32_t *out_key, uint32_t *out_val) {
    if (!ctx || kind < 0 || kind > 2 || !scalars_host || !out_key || !out_val || n == 0 || n > C(ctx)->n) return MSMB200_EINVAL;
    Ctx *c = C(ctx);
    MSM_CUDA(c, cudaSetDevice(c->device));
    size_t per = kind == 0 ? c->cfg.h : kind == 1 ? c->cfg.h_bgmw : c->pip_tiles;
    size_t m = n * per;
    if (ensure(c, c->scalars, n * 32) || ensure(c, c->keys, m * 4) || ensure(c, c->vals, m * 4)) return MSMB200_ECUDA;
    MSM_CUDA(c, cudaMemcpyAsync(c->scalars.p, scalars_host, n * 32, cudaMemcpyHostToDevice, c->stream));
    int rc = c->ops->digits(c, kind, c->scalars.p, n, (uint32_t *)c->keys.p, (uint32_t *)c->vals.p);
    if (rc) return rc;
    MSM_CUDA(c, cudaMemcpy(out_key, c->keys.p, m * 4, cudaMemcpyDeviceToHost));
    MSM_CUDA(c, cudaMemcpy(out_val, c->vals.p, m * 4, cudaMemcpyDeviceToHost));
    return MSMB200_OK;
}

// ---- blst-named shims ---------------------------------------------------------------------------------
// blst pointer-array convention (src/multi_scalar.c:401,:413): NULL means "next contiguous element".
static void gather_ptr_array(std::vector<unsigned char> &dst, const void *const ptrs[], size_t count, size_t elem, size_t out_elem) {
    dst.assign(count * out_elem, 0);
    const unsigned char *cur = nullptr;
    size_t pi = 0;
    for (size_t i = 0; i < count; i++) {
        if (i == 0) cur = (const unsigned char *)ptrs[pi++];
        else if (ptrs[pi] != nullptr) cur = (const unsigned char *)ptrs[pi++];
        else cur += elem;
        memcpy(&dst[i * out_elem], cur, elem);
    }
}
static void shim_fail(Ctx *c, const char *what) {
    fprintf(stderr, "msm_b200: %s failed: %s\n", what, c->err.c_str());
    abort();  // void blst signature: no error channel, and never a CPU fallback
}
// blst_pNs_mult_wbits_precompute / blst_pNs_mult_wbits (bindings/blst.h:228-236,:368-376; src/multi_scalar.c:81-261): the
// library's fixed-window table MSM. The table lives in HOST memory in the reference's layout (the caller owns it), so
// the shim uploads it per call; it is gathered by the same accumulate kernel as the CHES / BGMW95 tables.
static void shim_wbits_precompute(int group, void *table, size_t wbits, const void *const points[], size_t npoints) {
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    size_t ab = c->ops->aff_bytes;
    if (npoints == 0 || wbits == 0) return;
    if (wbits > 16) { c->err = "wbits > 16 is not supported"; shim_fail(c, "blst_pNs_mult_wbits_precompute"); }
    cudaSetDevice(c->device);
    std::vector<unsigned char> hp;
    gather_ptr_array(hp, points, npoints, ab, ab);
    const size_t total = npoints << (wbits - 1);
    void *dp = nullptr, *dt = nullptr;
    if (cudaMalloc(&dp, hp.size()) != cudaSuccess || cudaMalloc(&dt, total * ab) != cudaSuccess) { c->err = "cudaMalloc"; shim_fail(c, "blst_pNs_mult_wbits_precompute"); }
    cudaMemcpyAsync(dp, hp.data(), hp.size(), cudaMemcpyHostToDevice, c->stream);
    if (c->ops->wbits_precompute(c, dp, npoints, (int)wbits, dt)) shim_fail(c, "blst_pNs_mult_wbits_precompute");
    cudaMemcpyAsync(table, dt, total * ab, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, "blst_pNs_mult_wbits_precompute"); }
    cudaFree(dp); cudaFree(dt);
}
static void shim_mult_wbits(int group, void *ret, const void *table, size_t wbits, size_t npoints, const unsigned char *const scalars[], size_t nbits) {
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    size_t ab = c->ops->aff_bytes, jb = c->ops->jac_bytes;
    if (npoints == 0 || nbits == 0 || wbits == 0) { memset(ret, 0, jb); return; }
    if (nbits > 256 || wbits > 16 || ((double)npoints * (double)((size_t)1 << (wbits - 1))) >= 2147483648.0) {
        c->err = "nbits > 256, wbits > 16 or npoints * 2^(wbits-1) >= 2^31 is not supported";
        shim_fail(c, "blst_pNs_mult_wbits");
    }
    cudaSetDevice(c->device);
    std::vector<unsigned char> hs;
    gather_ptr_array(hs, (const void *const *)scalars, npoints, (nbits + 7) / 8, 32);
    const size_t total = npoints << (wbits - 1);
    void *dt = nullptr, *ds = nullptr, *dj = nullptr;
    if (cudaMalloc(&dt, total * ab) != cudaSuccess || cudaMalloc(&ds, hs.size()) != cudaSuccess || cudaMalloc(&dj, jb) != cudaSuccess) {
        c->err = "cudaMalloc"; shim_fail(c, "blst_pNs_mult_wbits");
    }
    cudaMemcpyAsync(dt, table, total * ab, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(ds, hs.data(), hs.size(), cudaMemcpyHostToDevice, c->stream);
    if (c->ops->pippenger(c, dt, npoints, ds, (int)nbits, dj, false, (int)wbits, -1, 0)) shim_fail(c, "blst_pNs_mult_wbits");
    cudaMemcpyAsync(ret, dj, jb, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, "blst_pNs_mult_wbits"); }
    cudaFree(dt); cudaFree(ds); cudaFree(dj);
}
// State of the literal shims that work on the CALLER's host arrays (one per group, guarded by g_shim_call_mu):
//  * the host precomputation table the drivers' pointer arrays point into (main_p1.cpp:219,:224,:330,:345,:365), mirrored in
//    HBM: registered explicitly (msmb200_blst_register_table) or learned from the pointer range of the first call, so
//    that later calls translate pointers to indices on the device instead of gathering and uploading n*h points;
//  * the bucket set / index map / reduction plan of the last CHES tile call and the dense plan of the last BGMW95 call;
//  * the {m, b, alpha} digit table of blst_pN_construct_nh_scalars_nh_points; grow-only device staging buffers.
// A host table mirrored in HBM. The allocation keeps MIRROR_SLACK entries of room on both sides: a table learned from the
// pointers of one call starts at the lowest pointer seen, and a later call may reach an entry or two further out.
constexpr size_t MIRROR_SLACK = 8;
struct Mirror {
    const unsigned char *base = nullptr; size_t entries = 0; unsigned char *d = nullptr; uint64_t probe = 0, used = 0;
    unsigned char *d_alloc = nullptr; size_t cap = 0, before = 0;   // allocation, its capacity in entries, entries of room before `d`
};
struct ShimState {
    Mirror mir[4]; uint64_t tick = 0;   // CHES table, BGMW95 table, the fixed points, one spare; least recently used is replaced
    const int *bs_key = nullptr; size_t bs_n = 0; int bs_last = -1, bs_mid = -1, bs_d = 0;
    DevBuf d_bs, d_v2i, d_cf; uint32_t vspan = 0, nchunks = 0; ReducePlan plan; size_t dense_n = 0; ReducePlan dense_plan; uint32_t dense_vspan = 0, dense_chunks = 0;
    const void *tri_key = nullptr; size_t tri_n = 0; DevBuf d_tri;
    DevBuf sc, sg, pi, ptrs, jac, aux, pts, in;
    unsigned char *h_stage = nullptr; size_t h_stage_bytes = 0;   // pinned staging for gathered host data
};
static ShimState g_state[3];
static uint64_t probe_table(const unsigned char *base, size_t entries, size_t ab) {  // a few entries, to notice a table rebuilt in place
    uint64_t h = 0;
    for (size_t k : {(size_t)0, entries / 3, entries / 2, entries - 1}) {
        uint64_t w[4];
        memcpy(w, base + k * ab + ab - 32, 32);
        h = (h ^ w[0] ^ (w[3] << 1)) * 0x9E3779B97F4A7C15ull + w[1] + w[2];
    }
    return h;
}
// mirror [base, base + entries): grow a mirror that overlaps the range and has the room, refresh the one with this base,
// else take the least recently used slot
static Mirror *shim_upload_table(Ctx *c, ShimState &S, const unsigned char *base, size_t entries) {
    const size_t ab = c->ops->aff_bytes;
    const unsigned char *end = base + entries * ab;
    for (Mirror &k : S.mir) {
        if (!k.base || !k.d || (k.base == base && k.entries == entries)) continue;
        const unsigned char *kend = k.base + k.entries * ab;
        if (end <= k.base || base >= kend || (size_t)(k.base > base ? k.base - base : base - k.base) % ab) continue;
        const size_t add_lo = base < k.base ? (size_t)(k.base - base) / ab : 0, add_hi = end > kend ? (size_t)(end - kend) / ab : 0;
        if (add_lo > k.before || k.before + k.entries + add_hi > k.cap) continue;
        if (probe_table(k.base, k.entries, ab) != k.probe) continue;   // stale: let the caller replace it as a whole
        if (add_lo && cudaMemcpyAsync(k.d - add_lo * ab, base, add_lo * ab, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return nullptr;
        if (add_hi && cudaMemcpyAsync(k.d + k.entries * ab, kend, add_hi * ab, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return nullptr;
        k.d -= add_lo * ab; k.before -= add_lo; k.base -= add_lo * ab; k.entries += add_lo + add_hi;
        k.probe = probe_table(k.base, k.entries, ab); k.used = ++S.tick;
        return &k;
    }
    Mirror *m = nullptr;
    for (Mirror &k : S.mir) if (k.base == base) m = &k;
    if (!m) { m = &S.mir[0]; for (Mirror &k : S.mir) if (k.used < m->used) m = &k; }
    if (m->d_alloc && m->cap < entries + 2 * MIRROR_SLACK) { cudaFree(m->d_alloc); m->d_alloc = nullptr; }
    m->base = nullptr; m->d = nullptr;
    if (!m->d_alloc) {
        if (cudaMalloc((void **)&m->d_alloc, (entries + 2 * MIRROR_SLACK) * ab) != cudaSuccess) { m->d_alloc = nullptr; cudaGetLastError(); return nullptr; }
        m->cap = entries + 2 * MIRROR_SLACK;
    }
    m->before = MIRROR_SLACK;
    m->d = m->d_alloc + MIRROR_SLACK * ab;
    if (cudaMemcpyAsync(m->d, base, entries * ab, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { m->d = nullptr; return nullptr; }
    m->base = base; m->entries = entries; m->probe = probe_table(base, entries, ab); m->used = ++S.tick;
    return m;
}
static bool pinned_stage(ShimState &S, size_t bytes) {
    if (S.h_stage_bytes >= bytes) return true;
    if (S.h_stage) cudaFreeHost(S.h_stage);
    S.h_stage = nullptr; S.h_stage_bytes = 0;
    if (cudaMallocHost((void **)&S.h_stage, bytes + bytes / 4) != cudaSuccess) { cudaGetLastError(); return false; }
    S.h_stage_bytes = bytes + bytes / 4;
    return true;
}
// points[0 .. npoints) are host pointers (already uploaded to S.ptrs). Returns the device array the indices written to S.pi
// refer to: a mirrored host table when every pointer lies inside one (registered, or learned from the range the pointers
// span - the drivers only ever pass pointers into one table per call), else the points gathered on the host and uploaded.
static const void *shim_resolve_points(Ctx *c, ShimState &S, const void *const points[], size_t npoints, const char *what) {
    const size_t ab = c->ops->aff_bytes;
    ShimAux x;
    x.ptrs = (unsigned long long *)S.ptrs.p; x.m = npoints; x.entry_bytes = (unsigned)ab; x.pidx = (uint32_t *)S.pi.p; x.bad = (uint32_t *)S.aux.p;
    const unsigned char *p0 = (const unsigned char *)points[0], *p1 = (const unsigned char *)points[npoints - 1];
    for (int attempt = 0; attempt < 2; attempt++) {
        for (Mirror &m : S.mir) {
            if (!m.base || !m.d || p0 < m.base || p0 >= m.base + m.entries * ab || p1 < m.base || p1 >= m.base + m.entries * ab) continue;
            if (probe_table(m.base, m.entries, ab) != m.probe) {   // rebuilt in place: upload again
                if (!shim_upload_table(c, S, m.base, m.entries)) break;
            }
            uint32_t bad = 0;
            cudaMemsetAsync(S.aux.p, 0, 4, c->stream);
            x.base = (unsigned long long)(uintptr_t)m.base; x.entries = m.entries;
            if (c->ops->shim_aux(c, 1, x)) shim_fail(c, what);
            cudaMemcpyAsync(&bad, S.aux.p, 4, cudaMemcpyDeviceToHost, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
            if (bad == 0) { m.used = ++S.tick; return m.d; }
        }
        if (attempt == 0) {  // learn the table from the range the pointers span
            unsigned long long mm[2] = {~0ull, 0ull};
            cudaMemcpyAsync((char *)S.aux.p + 16, mm, 16, cudaMemcpyHostToDevice, c->stream);
            ShimAux y = x;
            y.pidx = (uint32_t *)((char *)S.aux.p + 16);
            if (c->ops->shim_aux(c, 3, y)) shim_fail(c, what);
            cudaMemcpyAsync(mm, (char *)S.aux.p + 16, 16, cudaMemcpyDeviceToHost, c->stream);
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
            const unsigned long long span = mm[1] - mm[0];
            // a sparse set of pointers (far fewer points than the range holds) is not a table worth mirroring
            if (span % ab != 0 || span / ab + 1 > ((size_t)1 << 31) || span / ab + 1 > 64 * npoints + 4096 ||
                !shim_upload_table(c, S, (const unsigned char *)(uintptr_t)mm[0], span / ab + 1))
                break;
        }
    }
    // pointers into unrelated buffers: the general (slow) path of the blst convention
    if (!pinned_stage(S, npoints * (ab + 4)) || ensure(c, S.pts, npoints * ab)) { c->err = "staging memory"; shim_fail(c, what); }
    uint32_t *pidx = (uint32_t *)(S.h_stage + npoints * ab);
    for (size_t k = 0; k < npoints; k++) { memcpy(S.h_stage + k * ab, points[k], ab); pidx[k] = (uint32_t)k; }
    cudaMemcpyAsync(S.pts.p, S.h_stage, npoints * ab, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(S.pi.p, pidx, npoints * 4, cudaMemcpyHostToDevice, c->stream);
    cudaStreamSynchronize(c->stream);  // the staging buffer is reused by the next call
    return S.pts.p;
}
static void shim_mult_pippenger(int group, void *ret, const void *const points[], size_t npoints, const unsigned char *const scalars[],
                                size_t nbits, int tile_bit0 = -1, int tile_window = 0) {
    // tile_bit0 >= 0: blst_pNs_tile_pippenger — one window of tile_window bits starting at bit tile_bit0 (not shifted)
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    ShimState &S = g_state[group];
    const char *what = "blst_pNs_mult_pippenger";
    size_t ab = c->ops->aff_bytes, jb = c->ops->jac_bytes;
    if (npoints == 0 || nbits == 0) { memset(ret, 0, jb); return; }
    if (nbits > 256) { c->err = "nbits > 256 is not supported"; shim_fail(c, what); }
    if (tile_bit0 >= 0 && (tile_window < 1 || tile_window > 24 || (size_t)tile_bit0 >= nbits)) { memset(ret, 0, jb); return; }
    cudaSetDevice(c->device);
    const size_t sb = (nbits + 7) / 8;
    if (!pinned_stage(S, npoints * (ab + 32)) || ensure(c, S.sc, npoints * 32) || ensure(c, S.jac, jb)) { c->err = "staging memory"; shim_fail(c, what); }
    // scalars: blst's convention (an array of pointers, or one pointer followed by NULL for a contiguous array) -> 32-byte rows
    unsigned char *hs = S.h_stage;
    {
        const unsigned char *cur = nullptr;
        size_t pi = 0;
        for (size_t i = 0; i < npoints; i++) {
            if (i == 0) cur = scalars[pi++];
            else if (scalars[pi] != nullptr) cur = scalars[pi++];
            else cur += sb;
            memcpy(hs + i * 32, cur, sb);
            if (sb < 32) memset(hs + i * 32 + sb, 0, 32 - sb);
        }
    }
    cudaMemcpyAsync(S.sc.p, hs, npoints * 32, cudaMemcpyHostToDevice, c->stream);
    // points: one contiguous host array (the fixed-base case: FIX_POINTS_LIST, main_p1.cpp:416-418) is mirrored in HBM once
    // and found again by address; anything else is gathered and uploaded per call
    bool contiguous = true;
    const unsigned char *p0 = (const unsigned char *)points[0];
    if (npoints > 1 && points[1] != nullptr)
        for (size_t k = 1; k < npoints && contiguous; k++) contiguous = (const unsigned char *)points[k] == p0 + k * ab;
    const void *d_points = nullptr;
    if (contiguous && npoints >= 64) {
        for (Mirror &m : S.mir)
            if (m.base && m.d && p0 >= m.base && p0 + npoints * ab <= m.base + m.entries * ab && (size_t)(p0 - m.base) % ab == 0 &&
                probe_table(m.base, m.entries, ab) == m.probe) {
                m.used = ++S.tick;
                d_points = m.d + (p0 - m.base);
            }
        if (!d_points) {
            Mirror *m = shim_upload_table(c, S, p0, npoints);
            if (m) d_points = m->d + (p0 - m->base);
        }
    }
    if (!d_points) {
        unsigned char *hp = S.h_stage + npoints * 32;
        if (ensure(c, S.pts, npoints * ab)) shim_fail(c, what);
        const unsigned char *cur = nullptr;
        size_t pi = 0;
        for (size_t i = 0; i < npoints; i++) {
            if (i == 0) cur = (const unsigned char *)points[pi++];
            else if (points[pi] != nullptr) cur = (const unsigned char *)points[pi++];
            else cur += ab;
            memcpy(hp + i * ab, cur, ab);
        }
        cudaMemcpyAsync(S.pts.p, hp, npoints * ab, cudaMemcpyHostToDevice, c->stream);
        d_points = S.pts.p;
    }
    call_lock.lap("pippenger: inputs staged");
    if (c->ops->pippenger(c, d_points, npoints, S.sc.p, (int)nbits, S.jac.p, false, 0, tile_bit0, tile_window)) shim_fail(c, what);
    call_lock.lap("pippenger: enqueued");
    cudaMemcpyAsync(ret, S.jac.p, jb, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
}
static void shim_tile(int group, void *ret, const void *const points[], size_t npoints, const int scalars[], const unsigned char signs[],
                      const int *bucket_set_ascend, const int *v2i, size_t nbuckets, int d_max) {
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    ShimState &S = g_state[group];
    const size_t jb = c->ops->jac_bytes;
    if (npoints == 0) { memset(ret, 0, jb); return; }
    cudaSetDevice(c->device);
    const char *what = "blst_pN_tile_pippenger";
    if (ensure(c, S.sc, npoints * 4) || ensure(c, S.sg, npoints) || ensure(c, S.pi, npoints * 4) || ensure(c, S.ptrs, npoints * 8) || ensure(c, S.jac, jb) ||
        ensure(c, S.aux, 64))
        shim_fail(c, what);
    cudaMemcpyAsync(S.sc.p, scalars, npoints * 4, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(S.sg.p, signs, npoints, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(S.ptrs.p, points, npoints * 8, cudaMemcpyHostToDevice, c->stream);
    const void *d_table = shim_resolve_points(c, S, points, npoints, what);
    call_lock.lap("tile: inputs on the device");
    // ---- bucket set, index map and reduction plan: cached across calls ----
    const ReducePlan *plan = nullptr;
    const int *d_bs = nullptr, *d_v = nullptr, *d_cf = nullptr;
    uint32_t vspan = 0, nchunks = 0;
    if (bucket_set_ascend) {
        const int last = bucket_set_ascend[nbuckets - 1], mid = bucket_set_ascend[nbuckets / 2];
        if (!(S.bs_key == bucket_set_ascend && S.bs_n == nbuckets && S.bs_last == last && S.bs_mid == mid && S.plan.valid)) {
            S.vspan = pick_vspan_host((size_t)last, 1);
            std::vector<int> cf = build_chunk_first(bucket_set_ascend, nbuckets, S.vspan, &S.nchunks);
            if (ensure(c, S.d_bs, nbuckets * 4) || ensure(c, S.d_v2i, ((size_t)last + 1) * 4) || ensure(c, S.d_cf, cf.size() * 4)) shim_fail(c, what);
            cudaMemcpyAsync(S.d_bs.p, bucket_set_ascend, nbuckets * 4, cudaMemcpyHostToDevice, c->stream);
            cudaMemcpyAsync(S.d_v2i.p, v2i, ((size_t)last + 1) * 4, cudaMemcpyHostToDevice, c->stream);
            cudaMemcpyAsync(S.d_cf.p, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice, c->stream);
            cudaStreamSynchronize(c->stream);
            if (build_reduce_plan(c, S.plan, bucket_set_ascend, nbuckets, 1)) shim_fail(c, what);
            S.bs_key = bucket_set_ascend; S.bs_n = nbuckets; S.bs_last = last; S.bs_mid = mid;
        }
        plan = &S.plan; d_bs = (const int *)S.d_bs.p; d_v = (const int *)S.d_v2i.p; d_cf = (const int *)S.d_cf.p; vspan = S.vspan; nchunks = S.nchunks;
    } else {
        if (S.dense_n != nbuckets || !S.dense_plan.valid) {
            S.dense_vspan = pick_vspan_host(nbuckets - 1, 1);
            S.dense_chunks = (uint32_t)((nbuckets - 1 + S.dense_vspan - 1) / S.dense_vspan);
            if (build_reduce_plan(c, S.dense_plan, nullptr, nbuckets, 1)) shim_fail(c, what);
            S.dense_n = nbuckets;
        }
        plan = &S.dense_plan; vspan = S.dense_vspan; nchunks = S.dense_chunks;
    }
    if (c->ops->tile(c, d_table, (const int *)S.sc.p, (const unsigned char *)S.sg.p, (const uint32_t *)S.pi.p, npoints, d_v, d_bs, nbuckets, d_max, d_cf, vspan,
                     nchunks, plan, S.jac.p))
        shim_fail(c, what);
    cudaMemcpyAsync(ret, S.jac.p, jb, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
}
// blst_pN_construct_nh_scalars_nh_points (bindings/blst.h:274-276, src/multi_scalar.c:748-775), in place on the caller's
// arrays like the reference: digits in / bucket values out, signs out, host pointers into the caller's 3nh table out.
static void shim_construct_nh(int group, int nh_scalars[], unsigned char booth_signs[], void *nh_points_ptr[], size_t npoints, const void *table,
                              const void *triples) {
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    ShimState &S = g_state[group];
    if (npoints == 0) return;
    cudaSetDevice(c->device);
    const char *what = "blst_pN_construct_nh_scalars_nh_points";
    if (ensure(c, S.in, npoints * 4) || ensure(c, S.sc, npoints * 4) || ensure(c, S.sg, npoints) || ensure(c, S.ptrs, npoints * 8) || ensure(c, S.aux, 64))
        shim_fail(c, what);
    cudaMemcpyAsync(S.in.p, nh_scalars, npoints * 4, cudaMemcpyHostToDevice, c->stream);
    // the digit table has q + 1 entries; its length is not an argument: size the upload by the largest digit (+1 for a carry)
    ShimAux x;
    x.in = (const int *)S.in.p; x.m = npoints; x.out_b = (int *)S.aux.p;
    int maxd = 0;
    cudaMemsetAsync(S.aux.p, 0, 4, c->stream);
    if (c->ops->shim_aux(c, 2, x)) shim_fail(c, what);
    cudaMemcpyAsync(&maxd, S.aux.p, 4, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
    const size_t need = (size_t)maxd + 2;
    if (S.tri_key != triples || S.tri_n < need) {
        if (ensure(c, S.d_tri, need * 12)) shim_fail(c, what);
        cudaMemcpyAsync(S.d_tri.p, triples, need * 12, cudaMemcpyHostToDevice, c->stream);
        S.tri_key = triples; S.tri_n = need;
    }
    x.out_b = (int *)S.sc.p; x.signs = (unsigned char *)S.sg.p; x.ptrs = (unsigned long long *)S.ptrs.p; x.triples = (const int *)S.d_tri.p;
    x.base = (unsigned long long)(uintptr_t)table; x.entry_bytes = (unsigned)c->ops->aff_bytes;
    if (c->ops->shim_aux(c, 0, x)) shim_fail(c, what);
    cudaMemcpyAsync(nh_scalars, S.sc.p, npoints * 4, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(booth_signs, S.sg.p, npoints, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(nh_points_ptr, S.ptrs.p, npoints * 8, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, what); }
}

// blst_pNs_add (src/bulk_addition.c:145-164): sum of npoints affine points. One bucket holding every point, so the
// work is exactly the bucket-accumulation stage (batch-affine rounds or XYZZ work items + block combine).
static void shim_points_add(int group, void *ret, const void *const points[], size_t npoints) {
    Ctx *c = shim_ctx(group);
    ShimCall call_lock(group);
    size_t ab = c->ops->aff_bytes, jb = c->ops->jac_bytes;
    if (npoints == 0) { memset(ret, 0, jb); return; }
    cudaSetDevice(c->device);
    std::vector<unsigned char> hp;
    gather_ptr_array(hp, points, npoints, ab, ab);
    std::vector<int> ones(npoints, 1);
    std::vector<uint32_t> pidx(npoints);
    for (size_t k = 0; k < npoints; k++) pidx[k] = (uint32_t)k;
    void *dp = nullptr, *dsc = nullptr, *dsg = nullptr, *dpi = nullptr, *dj = nullptr;
    if (cudaMalloc(&dp, hp.size()) != cudaSuccess || cudaMalloc(&dsc, npoints * 4) != cudaSuccess || cudaMalloc(&dsg, npoints) != cudaSuccess ||
        cudaMalloc(&dpi, npoints * 4) != cudaSuccess || cudaMalloc(&dj, jb) != cudaSuccess) { c->err = "cudaMalloc"; shim_fail(c, "blst_pNs_add"); }
    cudaMemcpyAsync(dp, hp.data(), hp.size(), cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dsc, ones.data(), npoints * 4, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(dsg, 0, npoints, c->stream);
    cudaMemcpyAsync(dpi, pidx.data(), npoints * 4, cudaMemcpyHostToDevice, c->stream);
    if (c->ops->tile(c, dp, (const int *)dsc, (const unsigned char *)dsg, (const uint32_t *)dpi, npoints, nullptr, nullptr, 2, 1, nullptr, 8, 1, nullptr, dj))
        shim_fail(c, "blst_pNs_add");
    cudaMemcpyAsync(ret, dj, jb, cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = cudaGetErrorString(cudaGetLastError()); shim_fail(c, "blst_pNs_add"); }
    cudaFree(dp); cudaFree(dsc); cudaFree(dsg); cudaFree(dpi); cudaFree(dj);
}
// blst_pNs_to_affine (bindings/blst.h:222,:362; src/multi_scalar.c:17-59): batched Jacobian -> affine
static void shim_points_to_affine(int group, void *dst, const void *const points[], size_t npoints) {
    if (npoints == 0) return;
    const GroupOps *ops = group == 1 ? group_ops_g1() : group_ops_g2();
    std::vector<unsigned char> hp;
    gather_ptr_array(hp, points, npoints, ops->jac_bytes, ops->jac_bytes);
    const char *dev = getenv("MSMB200_DEVICE");
    if (msmb200_test_point_op(dev ? atoi(dev) : 0, group, 9, hp.data(), nullptr, nullptr, dst, npoints)) {
        fprintf(stderr, "msm_b200: blst_pNs_to_affine failed (no usable CUDA device?)\n");
        abort();  // void blst signature: no error channel, and never a CPU fallback
    }
}
void msmb200_blst_p1s_to_affine(void *dst, const void *const points[], size_t npoints) { shim_points_to_affine(1, dst, points, npoints); }
void msmb200_blst_p2s_to_affine(void *dst, const void *const points[], size_t npoints) { shim_points_to_affine(2, dst, points, npoints); }
void msmb200_blst_p1s_add(void *ret, const void *const points[], size_t npoints) { shim_points_add(1, ret, points, npoints); }
void msmb200_blst_p2s_add(void *ret, const void *const points[], size_t npoints) { shim_points_add(2, ret, points, npoints); }

size_t msmb200_blst_p1s_mult_wbits_precompute_sizeof(size_t wbits, size_t npoints) { return ((size_t)96 * npoints) << (wbits - 1); }
size_t msmb200_blst_p2s_mult_wbits_precompute_sizeof(size_t wbits, size_t npoints) { return ((size_t)192 * npoints) << (wbits - 1); }
void msmb200_blst_p1s_mult_wbits_precompute(void *table, size_t wbits, const void *const points[], size_t npoints) { shim_wbits_precompute(1, table, wbits, points, npoints); }
void msmb200_blst_p2s_mult_wbits_precompute(void *table, size_t wbits, const void *const points[], size_t npoints) { shim_wbits_precompute(2, table, wbits, points, npoints); }
size_t msmb200_blst_p1s_mult_wbits_scratch_sizeof(size_t) { return 0; }  /* the device owns its scratch */
size_t msmb200_blst_p2s_mult_wbits_scratch_sizeof(size_t) { return 0; }
void msmb200_blst_p1s_mult_wbits(void *ret, const void *table, size_t wbits, size_t npoints, const unsigned char *const scalars[], size_t nbits, void *) {
    shim_mult_wbits(1, ret, table, wbits, npoints, scalars, nbits);
}
void msmb200_blst_p2s_mult_wbits(void *ret, const void *table, size_t wbits, size_t npoints, const unsigned char *const scalars[], size_t nbits, void *) {
    shim_mult_wbits(2, ret, table, wbits, npoints, scalars, nbits);
}
size_t msmb200_blst_p1s_mult_pippenger_scratch_sizeof(size_t npoints) { return (size_t)192 << (pippenger_window_size(npoints) - 1); }
size_t msmb200_blst_p2s_mult_pippenger_scratch_sizeof(size_t npoints) { return (size_t)384 << (pippenger_window_size(npoints) - 1); }
void msmb200_blst_p1s_mult_pippenger(void *ret, const void *const points[], size_t npoints, const unsigned char *const scalars[], size_t nbits, void *) {
    shim_mult_pippenger(1, ret, points, npoints, scalars, nbits);
}
void msmb200_blst_p2s_mult_pippenger(void *ret, const void *const points[], size_t npoints, const unsigned char *const scalars[], size_t nbits, void *) {
    shim_mult_pippenger(2, ret, points, npoints, scalars, nbits);
}
// blst_pNs_tile_pippenger (bindings/blst.h:242-246,:382-386; src/multi_scalar.c:587-600): the tile-grid entry point the
// upstream Rust / Go bindings use to spread windows over threads
void msmb200_blst_p1s_tile_pippenger(void *ret, const void *const points[], size_t npoints, const unsigned char *const scalars[], size_t nbits, void *,
                                     size_t bit0, size_t window) {
    shim_mult_pippenger(1, ret, points, npoints, scalars, nbits, (int)bit0, (int)window);
}
void msmb200_blst_p2s_tile_pippenger(void *ret, const void *const points[], size_t npoints, const unsigned char *const scalars[], size_t nbits, void *,
                                     size_t bit0, size_t window) {
    shim_mult_pippenger(2, ret, points, npoints, scalars, nbits, (int)bit0, (int)window);
}
void msmb200_blst_p1_tile_pippenger_d_CHES(void *ret, const void *const points[], size_t npoints, const int scalars[], const unsigned char booth_signs[],
                                           void *, int bucket_set_ascend[], int bucket_value_to_its_index[], size_t bucket_set_size, int d_max) {
    shim_tile(1, ret, points, npoints, scalars, booth_signs, bucket_set_ascend, bucket_value_to_its_index, bucket_set_size, d_max);
}
void msmb200_blst_p2_tile_pippenger_d_CHES(void *ret, const void *const points[], size_t npoints, const int scalars[], const unsigned char booth_signs[],
                                           void *, int bucket_set_ascend[], int bucket_value_to_its_index[], size_t bucket_set_size, int d_max) {
    shim_tile(2, ret, points, npoints, scalars, booth_signs, bucket_set_ascend, bucket_value_to_its_index, bucket_set_size, d_max);
}
void msmb200_blst_p1_construct_nh_scalars_nh_points(int nh_scalars[], unsigned char booth_signs[], void *nh_points_ptr[], size_t npoints,
                                                    void *precomputation_points_list_3nh, const void *digit_conversion_hash_table) {
    shim_construct_nh(1, nh_scalars, booth_signs, nh_points_ptr, npoints, precomputation_points_list_3nh, digit_conversion_hash_table);
}
void msmb200_blst_p2_construct_nh_scalars_nh_points(int nh_scalars[], unsigned char booth_signs[], void *nh_points_ptr[], size_t npoints,
                                                    void *precomputation_points_list_3nh, const void *digit_conversion_hash_table) {
    shim_construct_nh(2, nh_scalars, booth_signs, nh_points_ptr, npoints, precomputation_points_list_3nh, digit_conversion_hash_table);
}
double msmb200_blst_last_call_ms(int group) { return group == 1 || group == 2 ? g_shim_last_ms[group] : -1.0; }
int msmb200_blst_register_table(int group, const void *host_table, size_t entries) {
    if ((group != 1 && group != 2) || !host_table || entries == 0 || entries > ((size_t)1 << 31)) return MSMB200_EINVAL;
    Ctx *c = shim_ctx(group, true);
    if (!c) return MSMB200_ECUDA;   // no CUDA device: msmb200_last_error(NULL) says why
    ShimCall call_lock(group);
    cudaSetDevice(c->device);
    if (!shim_upload_table(c, g_state[group], (const unsigned char *)host_table, entries) || cudaStreamSynchronize(c->stream) != cudaSuccess)
        return ctx_fail(c, MSMB200_ECUDA, "uploading the host table failed");
    return MSMB200_OK;
}
void msmb200_blst_p1_tile_pippenger_BGMW95(void *ret, const void *const points[], size_t npoints, const int scalars[], const unsigned char booth_signs[],
                                           void *, size_t q_exponent) {
    shim_tile(1, ret, points, npoints, scalars, booth_signs, nullptr, nullptr, ((size_t)1 << (q_exponent - 1)) + 1, 1);
}
void msmb200_blst_p2_tile_pippenger_BGMW95(void *ret, const void *const points[], size_t npoints, const int scalars[], const unsigned char booth_signs[],
                                           void *, size_t q_exponent) {
    shim_tile(2, ret, points, npoints, scalars, booth_signs, nullptr, nullptr, ((size_t)1 << (q_exponent - 1)) + 1, 1);
}

}  // extern "C"

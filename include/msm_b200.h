/*
 * msm_b200 — C ABI of the B200-native fixed-base MSM for BLS12-381 G1/G2.
 *
 * Drop-in boundary for the hot path of LuoGuiwen/MSM_blst. Every entry point names the
 * reference interface it replaces (paths relative to the reference root). Plain pointers and sizes only.
 *
 * Data encodings are the reference's own (bindings/blst.h:53-62,:164-165,:191-192,:251-252):
 *   field element  = 6 x u64 little-endian limbs, Montgomery form, fully reduced  (48 B; Fp2 = re,im 96 B)
 *   affine point   = {x, y}            96 B (G1) / 192 B (G2); infinity = all zero
 *   Jacobian point = {x, y, z}         144 B / 288 B;          infinity: z = 0
 *   scalar         = 32 bytes little-endian (blst_scalar / uint256_t::data[4]), < r, nbits = 255
 *
 * All msmb200_* functions return 0 on success, a negative MSMB200_E* code otherwise; the message is
 * available from msmb200_last_error(). There is no CPU fallback: without a usable CUDA device every
 * compute entry point fails with MSMB200_ECUDA.
 */
#ifndef MSM_B200_H
#define MSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSMB200_OK 0
#define MSMB200_EINVAL (-1)  /* bad argument / unknown config / call order */
#define MSMB200_ECUDA (-2)   /* CUDA runtime error (no device, OOM, launch failure) */
#define MSMB200_ESTATE (-3)  /* required table / points not built yet */

typedef struct msmb200_ctx msmb200_ctx;

/* MSM methods = the four methods of main_p1.cpp / main_p2.cpp (README.md:80-84) */
enum {
    MSMB200_CHES = 1,          /* pippenger_variant_q_over_5_CHES                               main_p1.cpp:192-246 */
    MSMB200_CHES_INTEGRAL = 2, /* pippenger_variant_q_over_5_CHES_integral_scalar_conversion    main_p1.cpp:249-291 */
    MSMB200_BGMW95 = 3,        /* pippenger_variant_BGMW95                                      main_p1.cpp:294-398 */
    MSMB200_PIPPENGER = 4      /* pippenger_blst_built_in -> blst_p1s_mult_pippenger            main_p1.cpp:400-436 */
};

/* One row of ches_config_files/config_file_n_exp_*.h:5-17 (compile-time constants there, run-time here). */
typedef struct {
    int n_exp;    /* N_EXP */
    int e;        /* EXPONENT_OF_q       (q = 2^e) */
    int h;        /* h_LEN_SCALAR */
    int a;        /* a_LEADING_TERM */
    int d;        /* d_MAX_DIFF */
    int bsize;    /* B_SIZE (0 = compute) */
    int e_bgmw;   /* EXPONENT_OF_q_BGMW95 */
    int h_bgmw;   /* h_BGMW95 */
} msmb200_config;

/* Look up a reference configuration by the name used on its command line (`./run.sh config=NAME`,
 * makefile:15-18): "8".."21", "16_beta", "17_beta", "20_beta". */
int msmb200_config_lookup(const char *name, msmb200_config *out);

/* Parameter search (main_bucket_set_construction.cpp: construct_bucket_set :39-72, check_bucket_set_validity :74-113,
 * max_gap_in_bucket_set :115-122) for ANY even radix q (not only the 17 shipped configurations) and leading term a:
 * out[0] = valid (both checks), out[1] = |B|, out[2] = largest gap d, out[3] = leading-digit check (every top digit
 * 0..a+1 is m*b), out[4] = first digit in [0, q] that is neither m*b nor q - m*b (-1 when all are covered). Host only. */
int msmb200_host_bucket_set_check(long q, long a, long out[5]);
/* Host-only parameter construction (no CUDA needed), the run-time form of construct_bucket_set
 * (auxiliaryfunc.h:257-288) and of the DIGIT_CONVERSION_HASH_TABLE fill (main_p1.cpp:140-152).
 * msmb200_host_bucket_set returns |B| (fills out[] when cap >= |B|). msmb200_host_digit_table writes q+1
 * triples (m, b, alpha) in the reference's digit_decomposition layout (bindings/blst.h:253). */
long msmb200_host_bucket_set(int e, int a, int *out, long cap);
int msmb200_host_digit_table(int e, int a, int *out_triples);
/* pippenger_window_size (src/multi_scalar.c:268-275) */
size_t msmb200_pippenger_window_size(size_t npoints);
/* Host-only check of the digit-splitting bucket-reduction plan (the static lists list_sum_kernel / list_sum_coop_kernel
 * walk, replacing integrate_buckets_accumulation_d_CHES, src/multi_scalar.c:301-321): evaluates the plan on 64-bit
 * integers x[0..nbw) instead of points; *out must equal sum_l value(l) * x[l] mod 2^64. values: ascending bucket values
 * (values[0] = 0) or NULL for dense (value == index). out_info (8 words, may be NULL): c_lo, bit positions, stage-1 slice
 * length, stage-1 slices, digit lists, groups per digit list, stage-2a lists, groups per 2a list. */
int msmb200_host_reduce_plan_eval(const int *values, size_t nbw, int resident_stage1, int resident_coop, int groups_per_warp,
                                  const uint64_t *x, uint64_t *out, uint32_t *out_info);

/* ---- context ------------------------------------------------------------------------------------ */

/* group: 1 = G1 (main_p1.cpp), 2 = G2 (main_p2.cpp). npoints: number of fixed points owned by this
 * context (N_POINTS, or the shard n/G of one GPU). device: CUDA ordinal. Builds BUCKET_SET,
 * BUCKET_VALUE_TO_ITS_INDEX and DIGIT_CONVERSION_HASH_TABLE (main_p1.cpp:128-153) and uploads them. */
int msmb200_ctx_create(msmb200_ctx **out, int group, const msmb200_config *cfg, size_t npoints, int device);
void msmb200_ctx_destroy(msmb200_ctx *ctx);
const char *msmb200_last_error(const msmb200_ctx *ctx); /* ctx may be NULL: last create() error */
/* Use the caller's CUDA stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). */
int msmb200_set_stream(msmb200_ctx *ctx, void *cuda_stream);

/* Multi-GPU by BUCKET RANGE (the alternative of SURVEY §8e): every context holds all points and tables; context
 * `rank` of `world` accumulates and reduces only its 1/world slice of the bucket-reduction chunks, so both hot
 * loops scale 1/world; the per-context results are Jacobian partials that sum to the full MSM (same gather as the
 * point-sharded mode). A sharded context (world > 1) always reduces with the chunked reducer (mode 1 below), the one that
 * works on a range of chunks. Default rank 0 of 1. */
int msmb200_set_bucket_shard(msmb200_ctx *ctx, int rank, int world);

/* Bucket-accumulation algorithm: 0 = library default, 1 = XYZZ mixed additions, one thread per work item (the
 * reference's xyzz_dadd_affine loop, src/multi_scalar.c:437-461), 2 = batch-affine pairwise rounds sharing one
 * inversion per batch (the reference's bulk_addition.c:51-143 analogue). Results are identical. */
int msmb200_set_accumulator(msmb200_ctx *ctx, int mode);

/* Bucket-reduction algorithm (replaces integrate_buckets_accumulation_d_CHES / integrate_buckets,
 * src/multi_scalar.c:281-321): 0 = library default, 1 = chunked running sums with the reference's gap
 * accumulators tmp_d[], 2 = digit splitting (bucket value = lo + 2^c * hi: two additions per bucket into
 * per-digit lists, then per-bit lists and one Horner pass; no long dependent chain). Results are identical. */
int msmb200_set_reducer(msmb200_ctx *ctx, int mode);

/* Performance knobs of one context (never change results). Keys: "ba_batch_max" (upper bound of the batch-affine slots
 * per lane per round), "ba_batch" (forced value, 0 = automatic), "ba_stagger" (staggered first batches of the round kernel,
 * default 1), "item_len" (XYZZ work-item length, 0 = automatic).
 * The same keys are read once from the environment at context creation as MSMB200_<KEY IN CAPITALS>. */
int msmb200_set_tuning(msmb200_ctx *ctx, const char *key, int value);

/* FIX_POINTS_LIST (main_p1.cpp:47): upload caller's affine points (host memory, npoints entries). */
int msmb200_set_points(msmb200_ctx *ctx, const void *points_affine_host);
/* init_fix_point_list (main_p1.cpp:52-66): P_i = 2^(first+i+1) * G computed on the device. */
int msmb200_generate_fix_points(msmb200_ctx *ctx, size_t first);
/* init_pippenger_CHES_q_over_5 table loop (main_p1.cpp:156-172): T3nh[3(i*h+j)+m-1] = m q^j P_i. In HBM every entry of the
 * context's own tables sits on its own 128-byte line (G2: two lines); that layout is internal (MSMB200_PACKED_TABLES at
 * context creation keeps the packed one). */
int msmb200_table_build_ches(msmb200_ctx *ctx);
/* init_pippenger_BGMW95 (main_p1.cpp:94-122): TBGMW[i*h'+j] = q'^j P_i. */
int msmb200_table_build_bgmw95(msmb200_ctx *ctx);
/* Copy `count` affine entries starting at `first` back to host, ALWAYS in the reference's packed array layout
 * (blst_pN_affine[count], byte-identical to PRECOMPUTATION_POINTS_LIST_3nh / _BGMW95). which: 0 = fixed points, 1 = T3nh,
 * 2 = TBGMW. */
int msmb200_download(msmb200_ctx *ctx, int which, size_t first, size_t count, void *out_host);
/* Table persistence (the reference rebuilds its tables on every run, main_p1.cpp:615-617). which: 0 fixed
 * points, 1 CHES 3nh table, 2 BGMW95 table. format 0: the in-memory blst_pN_affine layout (Montgomery limbs); format 1:
 * blst_pN_affine_serialize of every entry (src/e1.c:153-162, src/e2.c:194-203: 96 / 192 bytes big-endian, infinity
 * 0x40), readable by blst_pN_deserialize. An 80-byte header records group, configuration, npoints, entry count and a checksum of the
 * fixed points the array belongs to (a table is only accepted next to the points it was built from);
 * load checks it against the context and validates EVERY entry on the device like blst_pN_deserialize does
 * (flags, coordinates < p, y^2 = x^3 + B) before marking the array usable. */
int msmb200_table_save(msmb200_ctx *ctx, int which, const char *path, int format);
int msmb200_table_load(msmb200_ctx *ctx, int which, const char *path);
/* bucket set (BUCKET_SET, ascending, [0] = 0); returns |B| (or negative error); fills out if cap >= |B| */
long msmb200_bucket_set(msmb200_ctx *ctx, int *out, long cap);

/* ---- the MSM: one call = one method body of main_p1.cpp:192-436 incl. blst_p1_to_affine ----------- */

/* scalars_host: npoints x 32 bytes LE in host memory. out_affine_host: 96 B / 192 B affine struct
 * (Montgomery limbs, canonical) — the value the reference methods return. */
int msmb200_msm(msmb200_ctx *ctx, int method, const void *scalars_host, void *out_affine_host);
/* Same with scalars already resident in device memory (device pointer); result still lands on the host. */
int msmb200_msm_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_affine_host);
/* Multi-GPU leg: this context's partial sum as a Jacobian point written to DEVICE memory (144 B / 288 B),
 * asynchronously on the context's stream, ready for one NCCL all-gather. */
int msmb200_msm_partial_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_jacobian_dev);
/* Sum `count` Jacobian partials (device memory, contiguous) and normalise: the G-1 dadds + blst_p1_to_affine. */
int msmb200_sum_partials_device(msmb200_ctx *ctx, const void *partials_dev, int count, void *out_affine_host);
/* Multi-GPU leg with ONE Horner pass and ONE inversion for the whole job: every rank stops after the bucket reduction and
 * hands out its per-bit sums (layout[0] windows x layout[1] bit positions, blst_pNxyzz each: 192 B / 384 B) instead of a
 * finalised point; after one all-gather, any rank sums entry by entry over the ranks, runs the Horner pass over the bit
 * positions (bit k of window w weighs 2^(w * layout[2] + k)) and normalises. Replaces G x (xyzz_to_Jacobian + shift
 * loop) + (G - 1) dadd + to_affine (src/multi_scalar.c:565-575). All ranks must use the same configuration.
 * Precondition of the three *_device entry points: work is enqueued on the context's stream (msmb200_set_stream); the
 * caller's collective must be ordered after it on the same stream or by an event. */
int msmb200_msm_bits_layout(msmb200_ctx *ctx, int method, uint32_t layout[3]);
int msmb200_msm_bits_device(msmb200_ctx *ctx, int method, const void *scalars_dev, void *out_bits_dev);
int msmb200_combine_bits_device(msmb200_ctx *ctx, const void *gathered_dev, int world, const uint32_t layout[3], void *out_affine_host);
/* blst_p1_affine_serialize / blst_p2_affine_serialize (src/e1.c:153-162, src/e2.c:194-203): 96 / 192 bytes. */
int msmb200_affine_serialize(int group, const void *affine_host, unsigned char *out);

/* Per-phase device times (ms) of the last msm call, CUDA events on the context's stream:
 * [0] digits+histogram [1] sort (scan, scatter, items) [2] bucket accumulation [3] bucket reduction
 * [4] window combine + to_affine [5] total device time. */
int msmb200_last_timings(msmb200_ctx *ctx, float out_ms[6]);
/* Number of kernels launched by the last msm call. */
int msmb200_last_launches(msmb200_ctx *ctx);
/* Bucket accumulator the last MSM ran: 1 = XYZZ work items (xyzz_dadd_affine per entry), 2 = batch-affine rounds. */
int msmb200_last_accumulator(msmb200_ctx *ctx);

/* Live roofline denominators (register-only microbenchmarks): full 32x32+64-bit multiply-accumulates
 * per second with IMAD.WIDE.U32, and dependent-chain mul_mont_384 per second, on `device`. */
int msmb200_measure_peaks(int device, double *imad_macs_per_s, double *fp_mul_per_s);
/* Same run, all figures: out[0] = multiply-accumulates per second, out[1] = dependent mul_mont_384 per second,
 * out[2] = multiply-accumulates per clock per SM (32 on B200: one IMAD.WIDE per 4 cycles per sub-partition, whatever the
 * operands), out[3] = SM clock in MHz while the ~55 ms IMAD kernel ran. Both multiplicands of every IMAD.WIDE change
 * each iteration (ptxas folds loop-invariant products into additions). */
int msmb200_measure_peaks_ex(int device, double out[4]);

/* ---- blst-named drop-ins (signatures of bindings/blst.h:238-240,:274-283,:299-304) ------------------
 * Exported under an msmb200_ prefix so the library can be linked next to libblst.a; build the reference
 * drivers with -Dblst_p1s_mult_pippenger=msmb200_blst_p1s_mult_pippenger etc.. Pointer
 * arrays follow the blst convention: a NULL entry means "contiguous after the previous one". */
size_t msmb200_blst_p1s_mult_pippenger_scratch_sizeof(size_t npoints);
void msmb200_blst_p1s_mult_pippenger(void *ret_jacobian, const void *const points[], size_t npoints,
                                     const unsigned char *const scalars[], size_t nbits, void *scratch);
size_t msmb200_blst_p2s_mult_pippenger_scratch_sizeof(size_t npoints);
void msmb200_blst_p2s_mult_pippenger(void *ret_jacobian, const void *const points[], size_t npoints,
                                     const unsigned char *const scalars[], size_t nbits, void *scratch);
void msmb200_blst_p1_tile_pippenger_d_CHES(void *ret_jacobian, const void *const points[], size_t npoints,
                                           const int scalars[], const unsigned char booth_signs[], void *buckets,
                                           int bucket_set_ascend[], int bucket_value_to_its_index[],
                                           size_t bucket_set_size, int d_max);
void msmb200_blst_p2_tile_pippenger_d_CHES(void *ret_jacobian, const void *const points[], size_t npoints,
                                           const int scalars[], const unsigned char booth_signs[], void *buckets,
                                           int bucket_set_ascend[], int bucket_value_to_its_index[],
                                           size_t bucket_set_size, int d_max);
/* blst_p{1,2}_construct_nh_scalars_nh_points (bindings/blst.h:274-276, src/multi_scalar.c:748-775), literally: in place on
 * the caller's host arrays (standard q-ary digits in, bucket values out; booth signs out; host pointers into the caller's
 * 3nh table out; the alpha carry crosses slots exactly like the reference's sequential pass), computed on the device. */
void msmb200_blst_p1_construct_nh_scalars_nh_points(int nh_scalars[], unsigned char booth_signs[], void *nh_points_ptr[], size_t npoints,
                                                    void *precomputation_points_list_3nh, const void *digit_conversion_hash_table);
void msmb200_blst_p2_construct_nh_scalars_nh_points(int nh_scalars[], unsigned char booth_signs[], void *nh_points_ptr[], size_t npoints,
                                                    void *precomputation_points_list_3nh, const void *digit_conversion_hash_table);
/* Optional: mirror the caller's host precomputation table (PRECOMPUTATION_POINTS_LIST_3nh / _BGMW95, `entries` affine
 * points) in HBM now. The tile shims translate points[k] to (points[k] - host_table) / sizeof(affine) on the device and
 * gather from the mirror, so a call uploads 13 bytes per entry instead of n*h points. Without this call the shims learn
 * the table from the pointer range of the first call (the drivers only ever pass pointers into one table). A table that
 * is rebuilt in place is noticed (probe of sampled entries) and uploaded again. */
int msmb200_blst_register_table(int group, const void *host_table, size_t entries);
/* Wall time in ms (host to host, lock to return) of the group's most recent blst-named shim call. */
double msmb200_blst_last_call_ms(int group);
void msmb200_blst_p1_tile_pippenger_BGMW95(void *ret_jacobian, const void *const points[], size_t npoints,
                                           const int scalars[], const unsigned char booth_signs[], void *buckets,
                                           size_t q_exponent);
void msmb200_blst_p2_tile_pippenger_BGMW95(void *ret_jacobian, const void *const points[], size_t npoints,
                                           const int scalars[], const unsigned char booth_signs[], void *buckets,
                                           size_t q_exponent);

/* blst_p1s_to_affine / blst_p2s_to_affine (bindings/blst.h:222,:362; src/multi_scalar.c:17-59): batched normalisation of
 * npoints Jacobian points (pointer-array convention as above) into dst[npoints]; SURVEY §8f rank 3. */
void msmb200_blst_p1s_to_affine(void *dst_affine, const void *const points[], size_t npoints);
void msmb200_blst_p2s_to_affine(void *dst_affine, const void *const points[], size_t npoints);
/* blst_p1s_add / blst_p2s_add (bindings/blst.h:224,:364; src/bulk_addition.c:145-164): Jacobian sum of npoints affine
 * points (pointer-array convention as above). SURVEY §8f rank 1: the library's other consumer of bulk addition. */
void msmb200_blst_p1s_add(void *ret_jacobian, const void *const points[], size_t npoints);
void msmb200_blst_p2s_add(void *ret_jacobian, const void *const points[], size_t npoints);

/* blst_p{1,2}s_tile_pippenger (bindings/blst.h:242-246,:382-386; src/multi_scalar.c:587-600): ONE window of `window` bits
 * starting at bit `bit0` (the top one when bit0 + window > nbits), result not shifted — the tile-grid entry point of the
 * upstream Rust / Go bindings. scratch is ignored. */
void msmb200_blst_p1s_tile_pippenger(void *ret_jacobian, const void *const points[], size_t npoints,
                                     const unsigned char *const scalars[], size_t nbits, void *scratch, size_t bit0, size_t window);
void msmb200_blst_p2s_tile_pippenger(void *ret_jacobian, const void *const points[], size_t npoints,
                                     const unsigned char *const scalars[], size_t nbits, void *scratch, size_t bit0, size_t window);

/* blst_p{1,2}s_mult_wbits_precompute{,_sizeof} / blst_p{1,2}s_mult_wbits{,_scratch_sizeof} (bindings/blst.h:228-236,
 * :368-376; src/multi_scalar.c:81-261): the library's fixed-window table MSM (first half).
 * table[i * 2^(wbits-1) + k] = (k + 1) * P_i, affine, in HOST memory in the reference's layout (byte-identical: affine
 * points are canonical); mult_wbits recodes the scalars into signed wbits-windows like the reference, gathers the
 * selected row entries on the device (one bucket per window) and combines the windows by Horner. `scratch` is
 * ignored (scratch_sizeof returns 0). wbits in [1, 16], npoints * 2^(wbits-1) < 2^31. */
size_t msmb200_blst_p1s_mult_wbits_precompute_sizeof(size_t wbits, size_t npoints);
void msmb200_blst_p1s_mult_wbits_precompute(void *table_affine, size_t wbits, const void *const points[], size_t npoints);
size_t msmb200_blst_p1s_mult_wbits_scratch_sizeof(size_t npoints);
void msmb200_blst_p1s_mult_wbits(void *ret_jacobian, const void *table_affine, size_t wbits, size_t npoints,
                                 const unsigned char *const scalars[], size_t nbits, void *scratch);
size_t msmb200_blst_p2s_mult_wbits_precompute_sizeof(size_t wbits, size_t npoints);
void msmb200_blst_p2s_mult_wbits_precompute(void *table_affine, size_t wbits, const void *const points[], size_t npoints);
size_t msmb200_blst_p2s_mult_wbits_scratch_sizeof(size_t npoints);
void msmb200_blst_p2s_mult_wbits(void *ret_jacobian, const void *table_affine, size_t wbits, size_t npoints,
                                 const unsigned char *const scalars[], size_t nbits, void *scratch);

/* ---- building blocks exposed for parity tests (each is a batched CUDA kernel launch; host buffers) ---
 * field ops on n elements. field: 1 = Fp (48 B), 2 = Fp2 (96 B).
 * op: 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 5 mul_by_3, 6 inverse        (blst_fp_mul ... bindings/blst.h:108-137),
 *     7 inverse through the warp-level Montgomery trick of the batch-affine accumulator */
int msmb200_test_field_op(int device, int field, int op, const void *a, const void *b, void *out, size_t n);
/* point ops on n elements, group 1|2. op: 0 jac add-or-double (blst_p1_add_or_double), 1 jac double,
 * 2 xyzz += affine with sign flags (blst_p1xyzz_dadd_affine), 3 xyzz += xyzz (blst_p1xyzz_dadd),
 * 4 xyzz -> jacobian, 5 jacobian -> affine (blst_p1_to_affine), 6 / 7 quad-cooperative xyzz += xyzz / xyzz doubling
 * (csrc/coop.cuh: four lanes per point), 8 one-thread xyzz doubling, 9 batched jacobian -> affine (3 points per inversion) */
int msmb200_test_point_op(int device, int group, int op, const void *a, const void *b, const unsigned char *flags,
                          void *out, size_t n);
/* digit decomposition of the context's configuration for n scalars (host). kind 0: CHES -> out_key = bucket
 * index, out_val = table index | sign<<31 (h entries per scalar); kind 1: BGMW95 (h' entries);
 * kind 2: Pippenger windows (tiles entries, key = window*(2^(w-1)+1) + |digit|). */
int msmb200_test_digits(msmb200_ctx *ctx, int kind, const void *scalars_host, size_t n, uint32_t *out_key,
                        uint32_t *out_val);

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_H */

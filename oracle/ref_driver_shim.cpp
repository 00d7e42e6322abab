// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Wraps the UNMODIFIED reference driver (main_p1.cpp for GROUP=1, main_p2.cpp for GROUP=2) as a shared
// library so that tests can call the reference's own four MSM methods with seeded scalars:
//   pippenger_variant_q_over_5_CHES                              main_p1.cpp:192-246
//   pippenger_variant_q_over_5_CHES_integral_scalar_conversion   main_p1.cpp:249-291
//   pippenger_variant_BGMW95                                     main_p1.cpp:294-398
//   pippenger_blst_built_in                                      main_p1.cpp:400-436
// The driver source is #included by absolute path where it lies under /root/reference (never copied);
// its configuration is whatever ches_config_files/config_file.h holds there (config 10: n = 2^10).
// Built by oracle/Makefile into oracle/_ref/refdrv_p{1,2}.so (git-ignored).
// Variants built by oracle/Makefile from this same file:
//   refdrv_p{1,2}.so            the driver on the compiled reference (libblst_ref.so)                    [the oracle pin]
//   refdrv_p{1,2}_dropin.so     the driver with its five MSM imports re-pointed (-D...=msmb200_...) at libmsm_b200.so, so
//                               the reference's OWN glue runs on the CUDA shims with seeded scalars      [drop-in test]
//   refdrv_p1_dropin_c<N>.so    the same for configuration N: REF_MAIN names the driver inside a symlink farm
//                               (oracle/_ref/drv<N>/) whose ches_config_files/config_file.h links to the reference's
//                               config_file_n_exp_<N>.h - what the reference makefile does with cp (makefile:15-18)
#include <cstring>
#include <chrono>
#define main msmref_unused_main
#ifdef REF_MAIN
#include REF_MAIN
#if GROUP == 1
typedef blst_p1_affine ref_affine_t;
#define REF_SERIALIZE blst_p1_affine_serialize
#else
typedef blst_p2_affine ref_affine_t;
#define REF_SERIALIZE blst_p2_affine_serialize
#endif
#elif GROUP == 1
#include "/root/reference/main_p1.cpp"
typedef blst_p1_affine ref_affine_t;
#define REF_SERIALIZE blst_p1_affine_serialize
#else
#include "/root/reference/main_p2.cpp"
typedef blst_p2_affine ref_affine_t;
#define REF_SERIALIZE blst_p2_affine_serialize
#endif
#undef main

extern "C" {
int refdrv_n_exp() { return N_EXP; }
int refdrv_params(int *out8) {
    int v[8] = {N_EXP, EXPONENT_OF_q, h_LEN_SCALAR, a_LEADING_TERM, d_MAX_DIFF, B_SIZE, EXPONENT_OF_q_BGMW95, h_BGMW95};
    memcpy(out8, v, sizeof(v));
    return 0;
}
// main_p1.cpp:613-617
void refdrv_init() {
    init_fix_point_list();
    init_pippenger_CHES_q_over_5();
    init_pippenger_BGMW95();
}
// the fixed points only (the tables then come from refdrv_set_tables)
void refdrv_init_points() { init_fix_point_list(); }
size_t refdrv_table_entries(int which) { return which == 0 ? (size_t)3 * N_POINTS * h_LEN_SCALAR : (size_t)h_BGMW95 * N_POINTS; }
const void *refdrv_fix_points() { return FIX_POINTS_LIST; }
const void *refdrv_table(int which) { return which == 0 ? (const void *)PRECOMPUTATION_POINTS_LIST_3nh : (const void *)PRECOMPUTATION_POINTS_LIST_BGMW95; }
const int *refdrv_bucket_set() { return BUCKET_SET; }
const int *refdrv_hash() { return (const int *)DIGIT_CONVERSION_HASH_TABLE; }
static double g_last_ms = 0;
double refdrv_last_ms() { return g_last_ms; }   // wall time of the driver method of the last refdrv_msm call
// scalars: N_POINTS x 4 u64 LE limbs. out_struct: affine struct (Montgomery limbs); out_bytes: serialised
int refdrv_msm(int method, const uint64_t *scalars, void *out_struct, unsigned char *out_bytes) {
    uint256_t *arr = new uint256_t[N_POINTS];
    for (size_t i = 0; i < N_POINTS; i++)
        for (int k = 0; k < 4; k++) arr[i].data[k] = scalars[4 * i + k];
    ref_affine_t r;
    const auto t0 = std::chrono::steady_clock::now();
    switch (method) {
    case 1: r = pippenger_variant_q_over_5_CHES(arr); break;
    case 2: r = pippenger_variant_q_over_5_CHES_integral_scalar_conversion(arr); break;
    case 3: r = pippenger_variant_BGMW95(arr); break;
    case 4: r = pippenger_blst_built_in(arr); break;
    default: delete[] arr; return -1;
    }
    g_last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    delete[] arr;
    if (out_struct) memcpy(out_struct, &r, sizeof(r));
    if (out_bytes) REF_SERIALIZE(out_bytes, &r);
    return 0;
}
// digit conversions of the driver for one scalar (auxiliaryfunc.h:92-118 / :130-145)
void refdrv_digits(int kind, const uint64_t *scalar, int *out_m, int *out_b) {
    uint256_t s;
    for (int k = 0; k < 4; k++) s.data[k] = scalar[k];
    if (kind == 0) {
        scalar_MB_expr e;
        trans_uint256_t_to_MB_radixq_expr(e, s);
        for (int j = 0; j < h_LEN_SCALAR; j++) { out_m[j] = e[j][0]; out_b[j] = e[j][1]; }
    } else {
        std::array<int, h_BGMW95> e;
        trans_uint256_t_to_qhalf_expr(e, s);
        for (int j = 0; j < h_BGMW95; j++) out_m[j] = e[j];
    }
}
}

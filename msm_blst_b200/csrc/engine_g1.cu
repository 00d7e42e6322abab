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

static const GroupOps kOps = {
    sizeof(aff_t<fp_t>), sizeof(jac_t<fp_t>), sizeof(xyzz_t<fp_t>),
    msm_impl<fp_t, fpc_t>, generate_fix_points_impl<fpc_t>, table_build_impl<fpc_t>, sum_partials_impl<fpc_t>, tile_impl<fp_t, fpc_t>,
    pippenger_impl<fp_t, fpc_t>, field_op_g, point_op_impl<fp_t, fpc_t>, digits_impl<fp_t>};
const GroupOps *group_ops_g1() { return &kOps; }

}  // namespace msmb200

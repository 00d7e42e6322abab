// G2 instantiation of the MSM engine (F = Fp2).
#include "engine_impl.cuh"

namespace msmb200 {

// BLS12_381_G2 generator, affine Montgomery (reference src/e2.c:23-47): x.re, x.im, y.re, y.im
static const uint64_t kG2[24] = {
    0xf5f28fa202940a10ULL, 0xb3f5fb2687b4961aULL, 0xa1a893b53e2ae580ULL, 0x9894999d1a3caee9ULL, 0x6f67b7631863366bULL, 0x058191924350bcd7ULL,
    0xa5a9c0759e23f606ULL, 0xaaa0c59dbccd60c3ULL, 0x3bb17e18e2867806ULL, 0x1b1ab6cc8541b367ULL, 0xc2b6ed0ef2158547ULL, 0x11922a097360edf3ULL,
    0x4c730af860494c4aULL, 0x597cfa1f5e369c5aULL, 0xe7e6856caa0a635aULL, 0xbbefb5e96e0d495fULL, 0x07d3a975f0ef25a2ULL, 0x0083fd8e7e80dae5ULL,
    0xadc0fc92df64b05dULL, 0x18aa270a2b1461dcULL, 0x86adac6a3be4eba0ULL, 0x79495c4ec93da33aULL, 0xe7175850a43ccaedULL, 0x0b2bc2a163de1bf2ULL};
template <> const uint32_t *generator_words<fp2_t>() { return (const uint32_t *)kG2; }

static const GroupOps kOps = {
    sizeof(aff_t<fp2_t>), sizeof(jac_t<fp2_t>), sizeof(xyzz_t<fp2_t>),
    msm_impl<fp2_t, fp2_t>, generate_fix_points_impl<fp2_t>, table_build_impl<fp2_t>, sum_partials_impl<fp2_t>, combine_bits_impl<fp2_t>, tile_impl<fp2_t, fp2_t>,
    pippenger_impl<fp2_t, fp2_t>, shim_aux_impl, nullptr, point_op_impl<fp2_t, fp2_t>, digits_impl<fp2_t>, resident_blocks_impl<fp2_t, fp2_t>, wbits_precompute_impl<fp2_t>, table_io_impl<fp2_t>, checksum_impl};
const GroupOps *group_ops_g2() { return &kOps; }

}  // namespace msmb200

"""GPU tests at BASELINE.json's full sizes (-m gpu): known answers of SURVEY App. C, agreement of the four
methods, and size-independent properties (linearity, shard additivity)."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M(product_lib):
    import torch

    assert torch.cuda.is_available()
    return product_lib


def add_mod_r(a, b):
    """(a + b) mod r on n x 4 u64 arrays via Python ints (vectorised enough for 2^16)."""
    ia, ib = O.scalars_to_ints(a), O.scalars_to_ints(b)
    out = np.zeros_like(a)
    for i, (x, y) in enumerate(zip(ia, ib)):
        s = (x + y) % O.R_ORDER
        out[i] = [(s >> (64 * k)) & (2**64 - 1) for k in range(4)]
    return out


def test_g1_n16_all_methods_known_answer_and_linearity(M, golden):
    ctx = M.MsmContext(1, "16")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    # spot-check table rows against the oracle's single_scalar_multiplication chain
    h = ctx.cfg.h
    for i in (0, 1, 65535):
        oc = O.OracleCtx(1, "16", n=1, first=i)
        oc.init_fix_points()
        oc.build_table(0)
        assert (ctx.download(1, 3 * i * h, 3 * h) == oc.table(0)).all()
    sc = O.gen_scalars(1, ctx.n)
    res = [ctx.msm(m, sc) for m in (1, 2, 3, 4)]
    for r in res:
        assert M.affine_serialize(1, r).hex() == golden["kat_appc"]["g1_n16"]
    # linearity: MSM(s) + MSM(t) == MSM(s + t mod r)
    t = O.gen_scalars(2, ctx.n)
    st = add_mod_r(sc, t)
    a, b, c = ctx.msm(1, sc), ctx.msm(1, t), ctx.msm(1, st)
    one = np.frombuffer(bytes.fromhex("fdff02000000097602000cc40b00f4ebba58c7535798485f455752705358ce776dec56a2971a075c93e480fac35ef615"), dtype=np.uint8)
    parts = np.concatenate([a, one, b, one])
    chk = np.zeros(96, dtype=np.uint8)
    O.oracle().oracle_sum_partials(1, O.ptr(parts), 2, O.ptr(chk))
    assert (chk == c).all()
    for seed in (3, 4, 5):
        s2 = O.gen_scalars(seed, ctx.n)
        exp, _ = O.closed_form(1, s2)
        for m in (1, 3, 4):
            assert (ctx.msm(m, s2) == exp).all()
    ctx.close()


@pytest.mark.parametrize("cfgname", ["16_beta", "13"])
def test_g1_other_configs(M, cfgname):
    """_beta configuration and an r-a trick configuration (e'*h' == 255)."""
    ctx = M.MsmContext(1, cfgname)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    sc = O.gen_scalars(6, ctx.n)
    exp, _ = O.closed_form(1, sc)
    for m in (1, 2, 3, 4):
        assert (ctx.msm(m, sc) == exp).all(), m
    ctx.close()


def test_g1_n21_known_answer(M, golden):
    ctx = M.MsmContext(1, "21")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(1, ctx.n)
    r = ctx.msm(1, sc)
    assert M.affine_serialize(1, r).hex() == golden["kat_appc"]["g1_n21"]
    assert M.affine_serialize(1, ctx.msm(2, sc)).hex() == golden["kat_appc"]["g1_n21"]
    assert M.affine_serialize(1, ctx.msm(4, sc)).hex() == golden["kat_appc"]["g1_n21"]
    ctx.init_pippenger_BGMW95()
    assert M.affine_serialize(1, ctx.msm(3, sc)).hex() == golden["kat_appc"]["g1_n21"]
    sc2 = O.gen_scalars(2, ctx.n)
    exp, _ = O.closed_form(1, sc2)
    assert (ctx.msm(1, sc2) == exp).all()
    ctx.close()


def test_g2_n18_known_answer(M, golden):
    ctx = M.MsmContext(2, "18")
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    sc = O.gen_scalars(1, ctx.n)
    for m in (1, 2, 3, 4):
        assert M.affine_serialize(2, ctx.msm(m, sc)).hex() == golden["kat_appc"]["g2_n18"], m
    ctx.close()


def test_g1_n21_sharded_over_8_contexts_with_shard_configs(M, golden):
    """The 8-GPU decomposition emulated on one device: shard g uses the configuration tuned for n/8 = 2^18."""
    import torch

    from msm_blst_b200 import distributed as D

    n, world = 1 << 21, 8
    sc = O.gen_scalars(1, n)
    partials = torch.zeros((world, 144), dtype=torch.uint8, device="cuda")
    last = None
    for r in range(world):
        lo, hi = D.shard_range(n, r, world)
        ctx = M.MsmContext(1, D.shard_config_name(hi - lo), npoints=hi - lo, first=lo)
        ctx.init_fix_point_list()
        ctx.init_pippenger_CHES_q_over_5()
        d_sc = torch.from_numpy(sc[lo:hi].view(np.uint8).copy()).cuda()
        ctx.msm_partial_device(1, d_sc.data_ptr(), partials[r].data_ptr())
        torch.cuda.synchronize()
        if last is not None:
            last.close()
        last = ctx
    got = last.sum_partials_device(partials.data_ptr(), world)
    assert M.affine_serialize(1, got).hex() == golden["kat_appc"]["g1_n21"]
    last.close()


ALL_CONFIGS = ["8", "9", "10", "11", "12", "13", "14", "15", "16", "16_beta", "17", "17_beta", "18", "19", "20", "20_beta", "21"]


@pytest.mark.parametrize("cfgname", ALL_CONFIGS)
def test_g1_size_sweep_all_reference_configs(M, cfgname):
    """BASELINE configs[4]: every ches_config_files/config_file_n_exp_*.h row (n = 2^8 .. 2^21 incl. the _beta rows),
    G1, all four methods (CHES, its integral-scalar-conversion variant, BGMW95 with the r - a trick rows, blst's Pippenger);
    checked against the closed form of the synthetic points. At n = 2^16 the device-resident entry point is checked too."""
    ctx = M.MsmContext(1, cfgname)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    sc = O.gen_scalars(100 + len(cfgname) + ctx.cfg.n_exp, ctx.n)
    exp, _ = O.closed_form(1, sc)
    assert (ctx.msm(1, sc) == exp).all()
    assert (ctx.msm(2, sc) == exp).all()
    ctx.init_pippenger_BGMW95()
    assert (ctx.msm(3, sc) == exp).all()
    assert (ctx.msm(4, sc) == exp).all()
    if ctx.cfg.n_exp == 16:
        import torch

        d_sc = torch.from_numpy(sc.view(np.uint8).copy()).cuda()
        for m in (1, 2, 3, 4):
            assert (ctx.msm_device(m, d_sc.data_ptr()) == exp).all(), m
    ctx.close()


@pytest.mark.parametrize("cfgname", ALL_CONFIGS)
def test_g2_size_sweep_all_reference_configs(M, cfgname):
    ctx = M.MsmContext(2, cfgname)
    ctx.init_fix_point_list()
    ctx.init_pippenger_CHES_q_over_5()
    ctx.init_pippenger_BGMW95()
    sc = O.gen_scalars(200 + ctx.cfg.n_exp, ctx.n)
    exp, _ = O.closed_form(2, sc)
    for m in (1, 2, 3, 4):
        assert (ctx.msm(m, sc) == exp).all(), m
    ctx.close()

"""Pins the CPU oracle (oracle/lfbm5d_oracle.c) against the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile): every comparison here is bit-exact. Skipped where the reference is absent."""
import ctypes as C

import numpy as np
import pytest

import golden_inputs as gi
import lfdata

pytestmark = pytest.mark.filterwarnings("ignore")


def prep(oracle, H, W, sigma=25.0, n=24, aw=3, ah=3):
    clean = lfdata.synth_lf(aw, ah, H, W)
    noisy = oracle.add_noise(clean, sigma)
    y = noisy.copy()
    for st in range(aw * ah):
        oracle.lib().orc_color_space_transform(oracle.fp(y[st]), oracle.OPP, W, H, 3, 1)
    return clean, noisy, np.stack([oracle.symetrize(y[st], n) for st in range(aw * ah)])


def test_rng_and_noise(oracle, ref):
    ref.lib().ref_mt_seed(20171016)
    oracle.lib().orc_mt_seed(20171016)
    assert all(ref.lib().ref_mt_res53() == oracle.lib().orc_mt_res53() for _ in range(3000))


def test_small_helpers(oracle, ref):
    for (m, N, s) in [(289, 24, 4), (1057, 24, 4), (652, 21, 3), (100, 6, 1), (300, 16, 3)]:
        a, b = np.zeros(2000, np.uint32), np.zeros(2000, np.uint32)
        na = ref.lib().ref_ind_initialize(ref.up(a), m, N, s)
        nb = oracle.lib().orc_ind_initialize(oracle.up(b), m, N, s)
        assert na == nb and np.array_equal(a[:na], b[:nb])
    for k in (4, 8, 12, 16):
        r = [np.zeros(k * k, np.float32) for _ in range(3)]
        o = [np.zeros(k * k, np.float32) for _ in range(3)]
        ref.lib().ref_preProcess(*[ref.fp(x) for x in r], k)
        oracle.lib().orc_preProcess(*[oracle.fp(x) for x in o], k)
        assert all(np.array_equal(x, y) for x, y in zip(r, o))
    for a in (1, 3, 5):
        r = [np.zeros(a * a, np.float32) for _ in range(2)]
        o = [np.zeros(a * a, np.float32) for _ in range(2)]
        ref.lib().ref_preProcess_4d(ref.fp(r[0]), ref.fp(r[1]), a, a)
        oracle.lib().orc_preProcess_4d(oracle.fp(o[0]), oracle.fp(o[1]), a, a)
        assert all(np.array_equal(x, y) for x, y in zip(r, o))
        if a > 1:
            r = [np.zeros(a * a, np.float32) for _ in range(2)]
            o = [np.zeros(a * a, np.float32) for _ in range(2)]
            ref.lib().ref_preProcess_4d_sadct(ref.fp(r[0]), ref.fp(r[1]), a)
            oracle.lib().orc_preProcess_4d_sadct(oracle.fp(o[0]), oracle.fp(o[1]), a)
            assert all(np.array_equal(x, y) for x, y in zip(r, o))
    for cs in (oracle.YUV, oracle.YCBCR, oracle.OPP, oracle.RGB):
        r, o = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ref.lib().ref_estimate_sigma(C.c_float(10.0), ref.fp(r), 3, cs)
        oracle.lib().orc_estimate_sigma(C.c_float(10.0), oracle.fp(o), 3, cs)
        assert np.array_equal(r, o)
    rng = np.random.RandomState(1)
    img = (rng.rand(3, 20, 24) * 255).astype(np.float32)
    for cs in (oracle.YUV, oracle.YCBCR, oracle.OPP):
        for fwd in (1, 0):
            a, b = img.copy(), img.copy()
            ref.lib().ref_color_space_transform(ref.fp(a), cs, 24, 20, 3, fwd)
            oracle.lib().orc_color_space_transform(oracle.fp(b), cs, 24, 20, 3, fwd)
            assert np.array_equal(a, b)
    s1 = np.zeros((3, 20 + 14, 24 + 14), np.float32)
    ref.lib().ref_symetrize(ref.fp(img), ref.fp(s1), 24, 20, 3, 7)
    assert np.array_equal(s1, oracle.symetrize(img, 7))
    for (aidx, asize) in [(0, 17), (8, 17), (16, 17), (1, 3), (4, 5)]:
        ra, oa = [C.c_int() for _ in range(3)], [C.c_int() for _ in range(3)]
        ref.lib().ref_angular_search_window(*[C.byref(x) for x in ra], aidx, asize, 1)
        oracle.lib().orc_angular_search_window(*[C.byref(x) for x in oa], aidx, asize, 1)
        assert [x.value for x in ra] == [x.value for x in oa]


def test_transform_primitives(oracle, ref):
    rng = np.random.RandomState(2)
    for N in (2, 4, 8, 16, 32):
        v = rng.randn(N).astype(np.float32) * 50
        for fr, fo in (("ref_haar_forward", "orc_haar_forward"), ("ref_haar_inverse", "orc_haar_inverse"), ("ref_hadamard", "orc_hadamard")):
            a, b = v.copy(), v.copy()
            getattr(ref.lib(), fr)(ref.fp(a), N)
            getattr(oracle.lib(), fo)(oracle.fp(b), N)
            assert np.array_equal(a, b), (fr, N)
    for k in (8, 16):
        p = (rng.rand(k * k) * 255).astype(np.float32)
        a, b = np.zeros_like(p), np.zeros_like(p)
        ref.lib().ref_bior_2d_forward(ref.fp(p), ref.fp(a), k)
        oracle.lib().orc_bior_2d_forward(oracle.fp(p), oracle.fp(b), k)
        assert np.array_equal(a, b)
        ref.lib().ref_bior_2d_inverse(ref.fp(a), k)
        oracle.lib().orc_bior_2d_inverse(oracle.fp(b), k)
        assert np.array_equal(a, b)     # (the reference's bior pair is not a perfect-reconstruction pair on random data)
        for mode in (0, 1):
            ref.lib().ref_set_dct_mode(mode)
            oracle.set_dct_mode(mode)
            a, b = np.zeros_like(p), np.zeros_like(p)
            ref.lib().ref_dct_2d_patch(ref.fp(p), ref.fp(a), k)
            oracle.lib().orc_dct_2d_forward(oracle.fp(p), oracle.fp(b), k)
            assert np.array_equal(a, b)
            ref.lib().ref_dct_2d_inverse_patch(ref.fp(a), k)
            oracle.lib().orc_dct_2d_inverse(oracle.fp(b), k)
            assert np.array_equal(a, b) and np.abs(a - p).max() < 1e-2     # known answer: round trip
        ref.lib().ref_set_dct_mode(0)
        oracle.set_dct_mode(0)


def test_block_matching_lists_bit_exact(oracle, ref):
    clean, noisy, sym = prep(oracle, 128, 144)
    img = sym[4, 0]
    c1, i1 = ref.precompute_bm(img, 16, 8, 24, 18, 4, 3000.0)
    c2, i2 = oracle.bm_self(img, 16, 8, 24, 18, 4, 3000.0)
    m = np.arange(9)[None, :] < c1[:, None]
    assert np.array_equal(c1, c2) and np.array_equal(i1 * m, i2 * m) and (c1 > 0).sum() > 500
    # flat image: every distance ties -> libstdc++ heap/sort order decides
    flat = np.full_like(img, 7.0)
    c1, i1 = ref.precompute_bm(flat, 8, 16, 24, 18, 4, 2000.0)
    c2, i2 = oracle.bm_self(flat, 8, 16, 24, 18, 4, 2000.0)
    m = np.arange(17)[None, :] < c1[:, None]
    assert np.array_equal(c1, c2) and np.array_equal(i1 * m, i2 * m)
    nties = 0
    for st in (0, 8):
        f1, s1, _ = ref.precompute_bm_stereo(img, sym[st, 0], 16, 24, 6, 3000.0)
        f2, s2, ties = oracle.bm_stereo(img, sym[st, 0], 16, 24, 6, 3000.0)
        assert np.array_equal(f1, f2) and np.array_equal(s1, s2)
        nties += int(ties.sum())
    f1, s1, _ = ref.precompute_bm_stereo(flat, flat, 8, 24, 6, 2000.0)
    f2, s2, ties = oracle.bm_stereo(flat, flat, 8, 24, 6, 2000.0)
    assert np.array_equal(f1, f2) and np.array_equal(s1, s2) and ties.sum() > 1000


def test_std_sort_emulation(oracle, ref):
    # ties everywhere: quantised distances
    rng = np.random.RandomState(3)
    img1 = np.round(rng.rand(70, 80) * 3).astype(np.float32)
    img2 = np.round(rng.rand(70, 80) * 3).astype(np.float32)
    f1, s1, all1 = ref.precompute_bm_stereo(img1, img2, 8, 24, 6, 2000.0, want_all=True)
    f2, s2, ties = oracle.bm_stereo(img1, img2, 8, 24, 6, 2000.0)
    assert ties.sum() > 100 and np.array_equal(f1, f2) and np.array_equal(s1, s2)
    c1, i1 = ref.precompute_bm(img1, 8, 16, 24, 18, 4, 2000.0)
    c2, i2 = oracle.bm_self(img1, 8, 16, 24, 18, 4, 2000.0)
    m = np.arange(17)[None, :] < c1[:, None]
    assert np.array_equal(c1, c2) and np.array_equal(i1 * m, i2 * m)


@pytest.mark.parametrize("mode", [0, 1])
def test_single_pass_bit_exact(oracle, ref, mode):
    clean, noisy, sym = prep(oracle, 40, 48)
    z = np.zeros_like(sym)
    mask, proc = np.ones(9), np.zeros(9)
    ref.lib().ref_set_dct_mode(mode)
    oracle.set_dct_mode(mode)
    try:
        for (k, N, t2, t4, t5) in [(16, 8, ref.ID, ref.SADCT, ref.HAAR), (16, 1, ref.BIOR, ref.SADCT, ref.HAAR),
                                   (8, 16, ref.DCT, ref.DCT, ref.HADAMARD), (8, 8, ref.ID, ref.ID, ref.HAAR)]:
            rn, rd = ref.pass_step1(sym, z, z, mask, proc, 4, 4, 3, 25.0, 2.7, 18, 6, k, N, 4, t2, t4, t5)
            on, od = oracle.run_pass(1, sym, None, z, z, mask, proc, 4, 3, 25.0, 2.7, 18, 6, k, N, 4, t2, t4, t5)
            assert np.array_equal(rn, on) and np.array_equal(rd, od), (k, N, t2, t4, t5)
        est = np.where(rd != 0, rn / np.where(rd != 0, rd, 1), sym).astype(np.float32)
        for t5 in (ref.HAAR, ref.HADAMARD):
            rn2, rd2 = ref.pass_step2(sym, est, z, z, mask, proc, 4, 4, 3, 25.0, 18, 6, 8, 16, 4, ref.DCT, ref.SADCT, t5)
            on2, od2 = oracle.run_pass(2, sym, est, z, z, mask, proc, 4, 3, 25.0, 0.0, 18, 6, 8, 16, 4, ref.DCT, ref.SADCT, t5)
            assert np.array_equal(rn2, on2) and np.array_equal(rd2, od2)
        # an empty SAI in the window forces the shape-adaptive path for every group; a processed SAI is not aggregated
        mask2, proc2 = mask.copy(), proc.copy()
        mask2[2] = 0
        proc2[2] = 1
        proc2[7] = 1
        rn, rd = ref.pass_step1(sym, z, z, mask2, proc2, 4, 4, 3, 25.0, 2.7, 18, 6, 16, 8, 4, ref.ID, ref.SADCT, ref.HAAR)
        on, od = oracle.run_pass(1, sym, None, z, z, mask2, proc2, 4, 3, 25.0, 2.7, 18, 6, 16, 8, 4, ref.ID, ref.SADCT, ref.HAAR)
        assert np.array_equal(rn, on) and np.array_equal(rd, od)
        rn2, rd2 = ref.pass_step2(sym, est, rn, rd, mask2, proc2, 4, 4, 3, 25.0, 18, 6, 8, 16, 4, ref.DCT, ref.SADCT, ref.HAAR)
        on2, od2 = oracle.run_pass(2, sym, est, rn, rd, mask2, proc2, 4, 3, 25.0, 0.0, 18, 6, 8, 16, 4, ref.DCT, ref.SADCT, ref.HAAR)
        assert np.array_equal(rn2, on2) and np.array_equal(rd2, od2)
    finally:
        ref.lib().ref_set_dct_mode(0)
        oracle.set_dct_mode(0)


def test_full_run_multi_pass_bit_exact(oracle, ref):
    clean = lfdata.synth_lf(5, 5, 32, 40)
    noisy = oracle.add_noise(clean, 25.0)
    mask = np.ones(25)
    rb, rn = ref.run_step1(noisy, mask, 25.0, 2.7, 5, 5, 1, 8, 18, 6, 16, 4, ref.ID, ref.SADCT, ref.HAAR)
    ob, on, sch = oracle.run_step1(noisy, mask, 25.0, 2.7, 5, 5, 1, 8, 18, 6, 16, 4, oracle.ID, oracle.SADCT, oracle.HAAR)
    assert np.array_equal(rb, ob) and np.array_equal(rn, on) and len(sch) == 5
    rd, rb2, rn2 = ref.run_step2(rn, rb, mask, 25.0, 5, 5, 1, 16, 18, 6, 8, 4, ref.DCT, ref.SADCT, ref.HAAR)
    od, ob2, on2, sch = oracle.run_step2(on, ob, mask, 25.0, 5, 5, 1, 16, 18, 6, 8, 4, oracle.DCT, oracle.SADCT, oracle.HAAR)
    assert np.array_equal(rd, od) and np.array_equal(rb2, ob2) and np.array_equal(rn2, on2)
    assert oracle.psnr(od, clean)[0] > oracle.psnr(ob, clean)[0] > oracle.psnr(noisy, clean)[0] + 5


def test_lfbm3d_bit_exact(oracle, ref):
    """bm3d_LF path (config 4 parameters): A = 1 specialisation with BM3D's thresholds."""
    clean = lfdata.synth_lf(2, 1, 36, 44)
    noisy = oracle.add_noise(clean, 10.0)
    mask = np.ones(2)
    for (t2h, t2w) in ((ref.BIOR, ref.DCT), (ref.DCT, ref.BIOR)):
        rb, rd, rn = ref.run_bm3d_lf(noisy, mask, 10.0, 16, 16, 8, 8, 16, 32, 3, 3, t2h, t2w, 2.7)
        ob, od, on = oracle.run_bm3d_lf(noisy, mask, 10.0, 16, 16, 8, 8, 16, 32, 3, 3, t2h, t2w, 2.7)
        assert np.array_equal(rb, ob) and np.array_equal(rd, od) and np.array_equal(rn, on)


def test_partial_window_branch_bit_exact(oracle, ref):
    """`pst != cst` (core:531-821 / :1332-1658): single calls on accumulators with holes, and a complete grayscale run where
    every window takes several core calls (LF_denoised_percent < 100 after the first SAI when C == 1)."""
    _, _, sym = gi.pad_inputs(32, 40, 25.0)
    z = np.zeros_like(sym)
    mask, proc = np.ones(9, np.uint32), np.zeros(9, np.uint32)
    num, den = oracle.run_pass(1, sym, None, z, z, mask, proc, 4, 3, 25.0, 2.7, 18, 6, 16, 8, 4, oracle.ID, oracle.SADCT, oracle.HAAR)
    basic = np.where(den != 0, num / np.where(den != 0, den, 1), sym).astype(np.float32)
    num, den = gi.partial_holes(num, den)
    p2 = proc.copy(); p2[4] = 1
    for pst in (1, 6, 3):
        rn, rd = ref.pass_step1(sym, num, den, mask, p2, 4, pst, 3, 25.0, 2.7, 18, 6, 8, 4, 4, oracle.BIOR, oracle.DCT, oracle.HAAR)
        on, od = oracle.run_pass(1, sym, None, num, den, mask, p2, pst, 3, 25.0, 2.7, 18, 6, 8, 4, 4, oracle.BIOR, oracle.DCT, oracle.HAAR, cst=4)
        assert np.array_equal(rn, on) and np.array_equal(rd, od)
        assert np.array_equal(rn, num) == (pst == 3)          # SAI 3 has no holes: nothing to do
        rn, rd = ref.pass_step2(sym, basic, num, den, mask, p2, 4, pst, 3, 25.0, 18, 6, 8, 8, 4, oracle.DCT, oracle.SADCT, oracle.HAAR)
        on, od = oracle.run_pass(2, sym, basic, num, den, mask, p2, pst, 3, 25.0, 0.0, 18, 6, 8, 8, 4, oracle.DCT, oracle.SADCT, oracle.HAAR, cst=4)
        assert np.array_equal(rn, on) and np.array_equal(rd, od)
    clean = np.ascontiguousarray(lfdata.synth_lf(3, 3, 24, 28)[:, :1])
    noisy = oracle.add_noise(clean, 25.0)
    m = np.ones(9, np.uint32)
    m[4] = 0                                                      # empty centre SAI: already the first call is partial
    rb, rn_ = ref.run_step1(noisy, m, 25.0, 2.7, 3, 3, 1, 8, 18, 6, 16, 4, oracle.ID, oracle.SADCT, oracle.HAAR)
    ob, on_, sched = oracle.run_step1(noisy, m, 25.0, 2.7, 3, 3, 1, 8, 18, 6, 16, 4, oracle.ID, oracle.SADCT, oracle.HAAR)
    assert int(sched[0][3]) > 1 and np.array_equal(rb, ob)


def test_5d_dct_and_sd_weighting_bit_exact(oracle, ref):
    """tau_5D = dct (core:2524-2700 / :2943-3130) and useSD (core:3140-3173; bm3d.cpp:1345-1372 incl. its channel-0 quirk)."""
    _, _, sym = gi.pad_inputs(28, 32, 25.0)
    z = np.zeros_like(sym)
    mask, proc = np.ones(9, np.uint32), np.zeros(9, np.uint32)
    num, den = oracle.run_pass(1, sym, None, z, z, mask, proc, 4, 3, 25.0, 2.7, 18, 6, 16, 8, 4, oracle.ID, oracle.SADCT, oracle.HAAR)
    basic = np.where(den != 0, num / np.where(den != 0, den, 1), sym).astype(np.float32)
    for (step, k, N, t2, t4, t5, sd) in [(1, 16, 8, oracle.ID, oracle.SADCT, oracle.DCT, 0), (2, 8, 16, oracle.DCT, oracle.SADCT, oracle.DCT, 0),
                                         (1, 16, 8, oracle.ID, oracle.SADCT, oracle.HAAR, 1), (2, 8, 16, oracle.DCT, oracle.SADCT, oracle.HAAR, 1)]:
        oracle.lib().orc_set_use_sd(sd)
        try:
            on, od = oracle.run_pass(step, sym, basic if step == 2 else None, z, z, mask, proc, 4, 3, 25.0, 2.7, 18, 6, k, N, 4, t2, t4, t5)
        finally:
            oracle.lib().orc_set_use_sd(0)
        if step == 1:
            rn, rd = ref.pass_step1(sym, z, z, mask, proc, 4, 4, 3, 25.0, 2.7, 18, 6, k, N, 4, t2, t4, t5, useSD=bool(sd))
        else:
            rn, rd = ref.pass_step2(sym, basic, z, z, mask, proc, 4, 4, 3, 25.0, 18, 6, k, N, 4, t2, t4, t5, useSD=bool(sd))
        assert np.array_equal(rn, on) and np.array_equal(rd, od)
    import ctypes as C
    clean = lfdata.synth_lf(1, 1, 36, 40)
    noisy = oracle.add_noise(clean, 20.0)
    one = np.ones(1, np.uint32)
    n = noisy.copy(); rb = np.zeros_like(n); rden = np.zeros_like(n)
    assert ref.lib().ref_run_bm3d_LF(C.c_float(20.0), ref.fp(n), ref.up(ref.u32(one)), ref.fp(rb), ref.fp(rden), 1, 40, 36, 3, 16, 16, 8, 8, 16, 32, 3, 3,
                                     1, 1, oracle.BIOR, oracle.DCT, C.c_float(2.7), oracle.OPP, 1) == 0
    oracle.lib().orc_set_use_sd(1)
    try:
        ob, od, _ = oracle.run_bm3d_lf(noisy, one, 20.0, 16, 16, 8, 8, 16, 32, 3, 3, oracle.BIOR, oracle.DCT, 2.7)
    finally:
        oracle.lib().orc_set_use_sd(0)
    assert np.array_equal(rb, ob) and np.array_equal(rden, od)

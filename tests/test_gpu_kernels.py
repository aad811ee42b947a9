"""GPU parity tests of each C-ABI entry point (through leanfe_b200.engine -> liblfe.so) against the oracle.

Bit-exact for integer / index work (codes, counts, permutations, YOCO keys); fp64 tolerances are written
beside each assertion.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as O  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    from leanfe_b200.engine import get_engine
    return get_engine(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ------------------------------------------------------------------------------------------------
# (1) factorisation: bit-exact with np.unique / first-appearance
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["dense_i64", "shifted_i64", "negative_i32", "huge_range_i64", "u8", "single", "all_distinct"])
def test_factorize_sorted_bit_exact(eng, case):
    rng = np.random.default_rng(1)
    n = 200_003
    v = {
        "dense_i64": rng.integers(0, 5000, n),
        "shifted_i64": rng.integers(0, 700, n) * 7 + 12345,
        "negative_i32": rng.integers(-3000, 3000, n).astype(np.int32),
        "huge_range_i64": rng.choice(rng.integers(-2**62, 2**62, 4000), n),
        "u8": rng.integers(0, 3, n).astype(np.uint8),
        "single": np.full(n, 42, dtype=np.int64),
        "all_distinct": rng.permutation(n).astype(np.int64) * 1_000_003,
    }[case]
    codes, G, uniq = eng.factorize(dev(v), want_uniques=True)
    exp_codes, exp_uniq, _ = O.codes_sorted(v)
    assert G == len(exp_uniq)
    assert codes.dtype == torch.int32
    np.testing.assert_array_equal(codes.cpu().numpy(), exp_codes)
    np.testing.assert_array_equal(uniq.cpu().numpy(), exp_uniq)
    cnt = eng.group_counts(codes, G)
    np.testing.assert_array_equal(cnt.cpu().numpy(), np.bincount(exp_codes, minlength=G))


def test_factorize_float_keys_canonical(eng):
    rng = np.random.default_rng(2)
    v = rng.choice(np.array([0.0, -0.0, 1.5, -2.25, np.nan, np.inf, -np.inf, 1e-300, -1e-300]), 50_000)
    v[::7] = np.float64("nan")
    codes, G, uniq = eng.factorize(dev(v), want_uniques=True)
    exp_codes, exp_uniq, _ = O.codes_sorted(v)
    assert G == len(exp_uniq)
    np.testing.assert_array_equal(codes.cpu().numpy(), exp_codes)
    np.testing.assert_array_equal(uniq.cpu().numpy(), exp_uniq)          # NaN == NaN positionally
    assert not np.signbit(uniq.cpu().numpy()[exp_uniq == 0.0]).any()     # -0.0 was canonicalised to +0.0


def test_factorize_float_keys_large_takes_the_64_bit_radix_sort(eng):
    # 20M continuous doubles: 64-bit keys, 6 x 11-bit passes, 4 tiles per CTA and a ragged last CTA
    rng = np.random.default_rng(29)
    pool = rng.normal(size=700_000)
    v = pool[rng.integers(0, pool.size, 20_000_003)]
    codes, G, _ = eng.factorize(dev(v))
    u, inv = np.unique(v, return_inverse=True)
    assert G == u.size
    np.testing.assert_array_equal(codes.cpu().numpy(), inv)


def test_factorize_integer_valued_doubles_take_the_direct_path(eng):
    rng = np.random.default_rng(22)
    v = rng.integers(-3, 40, 120_000).astype(np.float64)
    v[::11] = -0.0
    codes, G, uniq = eng.factorize(dev(v), want_uniques=True)
    exp_codes, exp_uniq, _ = O.codes_sorted(v)
    assert G == len(exp_uniq)
    np.testing.assert_array_equal(codes.cpu().numpy(), exp_codes)
    np.testing.assert_array_equal(uniq.cpu().numpy(), exp_uniq)
    d = (rng.random(50_000) < 0.3).astype(np.float64)                      # a 0/1 dummy stored as double
    codes, G, uniq = eng.factorize(dev(d), want_uniques=True)
    assert G == 2 and uniq.cpu().tolist() == [0.0, 1.0]
    np.testing.assert_array_equal(codes.cpu().numpy(), d.astype(np.int32))


@pytest.mark.parametrize("kind", ["small_range", "large_range", "sorted", "gappy", "f64_integral"])
def test_factorize_counts_by_product(eng, kind):
    # the direct-address path counts occurrences instead of marking presence: group sizes come for free
    rng = np.random.default_rng(23)
    n = 300_011
    if kind == "small_range":
        a = rng.integers(-7, 490, n)
    elif kind == "large_range":
        a = rng.integers(10**9, 10**9 + 900_000, n)
    elif kind == "sorted":
        a = np.sort(rng.integers(0, 6_000, n))
    elif kind == "gappy":
        a = rng.integers(0, 40_000, n) * 3 + 5
    else:
        a = rng.integers(0, 20_000, n).astype(np.float64)
    codes, G, _ = eng.factorize(dev(a))
    counts = eng.factorize_counts(codes, G)
    u, inv, cnt = np.unique(a, return_inverse=True, return_counts=True)
    assert G == u.size
    np.testing.assert_array_equal(codes.cpu().numpy(), inv)
    assert counts is not None
    np.testing.assert_array_equal(counts.cpu().numpy(), cnt)
    assert eng.factorize_counts(codes, G) is None          # consumed: the handle keeps no copy
    np.testing.assert_array_equal(eng.group_counts(codes, G).cpu().numpy(), cnt)


def test_factorize_first_appearance(eng):
    rng = np.random.default_rng(3)
    v = rng.integers(0, 3000, 100_000) * 11 - 500
    codes, G, uniq = eng.factorize(dev(v), order=1, want_uniques=True)
    exp_codes, exp_uniq = O.codes_first_appearance(v)
    assert G == len(exp_uniq)
    np.testing.assert_array_equal(codes.cpu().numpy(), exp_codes)
    np.testing.assert_array_equal(uniq.cpu().numpy(), exp_uniq)


def test_factorize_empty(eng):
    codes, G, _ = eng.factorize(torch.empty(0, dtype=torch.int64, device="cuda"))
    assert G == 0 and codes.numel() == 0


# ------------------------------------------------------------------------------------------------
# singleton mask / compaction / ordering
# ------------------------------------------------------------------------------------------------
def test_singleton_mask_compact_sort(eng):
    rng = np.random.default_rng(4)
    n = 150_001
    a = rng.integers(0, 60_000, n).astype(np.int32)      # many singletons
    b = rng.integers(0, 40, n).astype(np.int32)
    x = rng.normal(size=n)
    keep, n_keep, dims = eng.singleton_mask([dev(a), dev(b)], [60_000, 40])
    exp = O.singleton_keep_mask([a, b])
    np.testing.assert_array_equal(keep.cpu().numpy().astype(bool), exp)
    assert n_keep == int(exp.sum()) and n_keep < n
    assert dims == (np.unique(a[exp]).size, np.unique(b[exp]).size)
    (xo,), (ao, bo) = eng.compact(keep, n_keep, [dev(x)], [dev(a), dev(b)])
    np.testing.assert_array_equal(xo.cpu().numpy(), x[exp])
    np.testing.assert_array_equal(ao.cpu().numpy(), a[exp])
    np.testing.assert_array_equal(bo.cpu().numpy(), b[exp])
    perm, was_sorted = eng.sort_rows(ao, 60_000)
    assert not was_sorted
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(a[exp], kind="stable"))    # stable => bit-exact
    (xs,), (as_,) = eng.gather(perm, [xo], [ao])
    np.testing.assert_array_equal(as_.cpu().numpy(), np.sort(a[exp], kind="stable"))
    np.testing.assert_array_equal(xs.cpu().numpy(), x[exp][np.argsort(a[exp], kind="stable")])
    perm2, was_sorted2 = eng.sort_rows(as_, 60_000)
    assert was_sorted2
    np.testing.assert_array_equal(perm2.cpu().numpy(), np.arange(n_keep))


@pytest.mark.parametrize("G", [1, 2, 255, 256, 257, 70_000, 3_000_000])
def test_sort_rows_radix_widths(eng, G):
    rng = np.random.default_rng(G)
    n = 300_007
    a = rng.integers(0, G, n).astype(np.int32)
    perm, _ = eng.sort_rows(dev(a), G)
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(a, kind="stable"))


@pytest.mark.parametrize("kind", ["random", "constant", "sorted", "two_values", "zipf"])
def test_sort_rows_large_multi_tile_and_skewed(eng, kind):
    # 20.5M rows: the sort walks 4 tiles per CTA (running digit offsets) and the last CTA is ragged
    rng = np.random.default_rng(5)
    n = 20_500_003
    G = 1_500_000          # 21 bits -> two 11-bit passes
    if kind == "random":
        a = rng.integers(0, G, n)
    elif kind == "constant":
        a = np.full(n, G - 1)
    elif kind == "sorted":
        a = np.sort(rng.integers(0, G, n))
    elif kind == "two_values":
        a = rng.integers(0, 2, n) * (G - 1)
    else:
        a = np.minimum(rng.zipf(1.3, n) - 1, G - 1)
    a = a.astype(np.int32)
    perm, was_sorted = eng.sort_rows(dev(a), G)
    if kind in ("constant", "sorted"):
        assert was_sorted
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(a, kind="stable"))


# ------------------------------------------------------------------------------------------------
# (2) demean
# ------------------------------------------------------------------------------------------------
def _panel(rng, n, Gs, C):
    codes = [rng.integers(0, g, n).astype(np.int32) for g in Gs]
    cols = [rng.normal(size=n) + sum(rng.normal(size=g)[c] for g, c in zip(Gs, codes)) for _ in range(C)]
    return cols, codes


@pytest.mark.parametrize("Gs", [(1_000_000, 500), (500, 1_000_000), (70_000, 60_000), (5, 3), (1, 1), (2_000_000, 1),
                                (40_000, 65_536)])
def test_demean_two_fe_packed_entry_lists(eng, Gs):
    # unweighted two-FE fits whose codes fit one 32-bit word together take the packed-word entry lists
    rng = np.random.default_rng(17)
    n = 400_003
    cols, codes = _panel(rng, n, Gs, 3)
    order = O.fe_order(list(Gs))
    exp_cols, exp_it, exp_mm = O.alt_proj(cols, codes, order, 0.0, 6, None)
    d_cols = [dev(c) for c in cols]
    it, mm, conv = eng.demean(d_cols, [dev(c) for c in codes], list(Gs), order, None, 0.0, 6, 3)
    assert it == exp_it == 6 and not conv
    for got, exp in zip(d_cols, exp_cols):
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=2e-12)
    # (1, 1): the group mean of a constant FE after the sweep is pure summation noise, ~1e-13 either way
    assert mm == pytest.approx(exp_mm, rel=1e-6, abs=1e-12)


@pytest.mark.parametrize("Gs,C,weighted", [((200_000, 5_000, 30), 3, False), ((200_000, 5_000, 30), 3, True),
                                           ((100_000, 12), 6, False), ((30, 150_000), 2, True),
                                           ((40_000, 9, 4, 2), 5, False), ((60_000, 12), 8, False),
                                           ((60_000, 12), 11, False)])
def test_demean_tiny_fe_projected_over_the_big_list(eng, Gs, C, weighted):
    # FEs with G x C <= 96 are projected while streaming the big FE's entry list (thread-private bins)
    rng = np.random.default_rng(31)
    n = 400_003
    cols, codes = _panel(rng, n, Gs, C)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    order = O.fe_order(list(Gs))
    exp_cols, exp_it, exp_mm = O.alt_proj(cols, codes, order, 0.0, 5, w)
    d_cols = [dev(c) for c in cols]
    it, mm, conv = eng.demean(d_cols, [dev(c) for c in codes], list(Gs), order, dev(w) if weighted else None,
                              0.0, 5, 3)
    assert it == exp_it == 5 and not conv
    for got, exp in zip(d_cols, exp_cols):
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=2e-12)
    assert mm == pytest.approx(exp_mm, rel=1e-6, abs=1e-14)


@pytest.mark.parametrize("Gs,C,weighted", [((3000, 40), 3, False), ((3000, 40), 6, True), ((500, 200, 12), 2, False),
                                           ((50,), 4, False), ((2000, 300), 11, False), ((64, 8, 5, 3), 1, True)])
def test_demean_matches_oracle(eng, Gs, C, weighted):
    rng = np.random.default_rng(sum(Gs) + C)
    n = 120_000
    cols, codes = _panel(rng, n, Gs, C)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    order = O.fe_order(list(Gs))
    max_iter = 7
    exp_cols, exp_it, exp_mm = O.alt_proj(cols, codes, order, 0.0, max_iter, w)         # fixed sweep count
    d_cols = [dev(c) for c in cols]
    it, mm, conv = eng.demean(d_cols, [dev(c) for c in codes], list(Gs), order, dev(w) if weighted else None,
                              0.0, max_iter, 3)
    assert it == exp_it == max_iter and not conv
    for got, exp in zip(d_cols, exp_cols):
        # same Gauss-Seidel iterates; only fp64 summation order differs (values are O(1))
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=2e-12)
    assert mm == pytest.approx(exp_mm, rel=1e-6, abs=1e-14)


def test_demean_convergence_and_iteration_parity(eng):
    rng = np.random.default_rng(11)
    n = 400_000
    cols, codes = _panel(rng, n, (20_000, 100), 3)
    order = [1, 0]
    exp_cols, exp_it, exp_mm = O.alt_proj(cols, codes, order, 1e-6, 50)
    d_cols = [dev(c) for c in cols]
    it, mm, conv = eng.demean(d_cols, [dev(c) for c in codes], [20_000, 100], order, None, 1e-6, 50, 3)
    assert conv and it == exp_it and 3 <= it < 10
    assert mm < 1e-6 and mm == pytest.approx(exp_mm, rel=1e-5, abs=1e-13)
    for got, exp in zip(d_cols, exp_cols):
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=1e-12)
    # idempotence: demeaning demeaned columns changes nothing beyond the tolerance
    before = [c.clone() for c in d_cols]
    eng.demean(d_cols, [dev(c) for c in codes], [20_000, 100], order, None, 1e-6, 50, 3)
    for a, b in zip(before, d_cols):
        assert float((a - b).abs().max()) < 5e-6


def test_demean_presorted_and_unsorted_agree(eng):
    rng = np.random.default_rng(12)
    n = 200_000
    cols, codes = _panel(rng, n, (9000, 30), 2)
    p = np.argsort(codes[0], kind="stable")
    a = [dev(c) for c in cols]
    b = [dev(c[p]) for c in cols]
    eng.demean(a, [dev(c) for c in codes], [9000, 30], [1, 0], None, 1e-8, 50, 3)
    eng.demean(b, [dev(c[p]) for c in codes], [9000, 30], [1, 0], None, 1e-8, 50, 3)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x.cpu().numpy()[p], y.cpu().numpy(), rtol=0, atol=1e-12)


def test_demean_single_fe_within_transform(eng):
    rng = np.random.default_rng(13)
    n = 100_000
    cols, codes = _panel(rng, n, (777,), 3)
    d_cols = [dev(c) for c in cols]
    it, _, _ = eng.demean(d_cols, [dev(codes[0])], [777], [0], None, 1e-6, 1, 2)
    assert it == 1
    for got, c in zip(d_cols, cols):
        exp = c - (np.bincount(codes[0], weights=c, minlength=777) / np.bincount(codes[0], minlength=777))[codes[0]]
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=1e-12)


def test_demean_empty_groups_allowed(eng):
    rng = np.random.default_rng(14)
    n = 50_000
    cols, codes = _panel(rng, n, (300, 20), 2)
    codes[0] = codes[0] * 3            # two thirds of the code space is empty
    exp_cols, _, _ = O.alt_proj(cols, codes, [1, 0], 0.0, 5)
    d_cols = [dev(c) for c in cols]
    eng.demean(d_cols, [dev(c) for c in codes], [900, 20], [1, 0], None, 0.0, 5, 3)
    for got, exp in zip(d_cols, exp_cols):
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=2e-12)


# ------------------------------------------------------------------------------------------------
# (4) gram + solve
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,weighted,intercept", [(1, False, True), (3, False, True), (7, True, True), (11, False, True),
                                                  (12, False, False), (13, True, True), (24, False, True), (40, True, False)])
def test_gram_matches_numpy(eng, C, weighted, intercept):
    rng = np.random.default_rng(C)
    n = 77_777
    cols = [rng.normal(size=n) * (1 + j) + j for j in range(C)]
    w = rng.uniform(0.1, 3.0, n) if weighted else None
    Z = np.column_stack(([np.ones(n)] if intercept else []) + cols)
    exp = Z.T @ (Z * (w[:, None] if weighted else 1.0))
    got = eng.gram([dev(c) for c in cols], dev(w) if weighted else None, add_intercept=intercept).cpu().numpy()
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-9)
    np.testing.assert_array_equal(got, got.T)


@pytest.mark.parametrize("p", [1, 2, 6, 11, 33, 100])
def test_solve_spd_matches_numpy(eng, p):
    rng = np.random.default_rng(p)
    M = rng.normal(size=(4 * p + 5, p))
    A, b = M.T @ M, rng.normal(size=p)
    beta, Ainv, fb = eng.solve_spd(dev(A), dev(b))
    eb, eA, efb = O.solve_spd(A, b)
    assert fb == efb is False
    np.testing.assert_allclose(beta.cpu().numpy(), eb, rtol=1e-9)
    np.testing.assert_allclose(Ainv.cpu().numpy(), eA, rtol=1e-8, atol=1e-12)


def test_solve_spd_fallback_and_singular(eng):
    from leanfe_b200._capi import LfeError
    A = np.array([[2.0, 1.0], [1.0, -3.0]])          # symmetric, indefinite -> Cholesky fails -> LU path
    b = np.array([1.0, 2.0])
    beta, Ainv, fb = eng.solve_spd(dev(A), dev(b))
    assert fb
    np.testing.assert_allclose(beta.cpu().numpy(), np.linalg.solve(A, b), rtol=1e-12)
    np.testing.assert_allclose(Ainv.cpu().numpy(), np.linalg.inv(A), rtol=1e-12)
    with pytest.raises(LfeError, match="singular"):
        eng.solve_spd(dev(np.zeros((2, 2))), dev(b))


# ------------------------------------------------------------------------------------------------
# (5) residuals + vcov
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,weighted", [(2, False), (5, True), (10, False)])
def test_resid_vcov_all_kinds(eng, k, weighted):
    rng = np.random.default_rng(100 + k)
    n = 60_000
    X = rng.normal(size=(n, k))
    y = X @ rng.normal(size=k) + rng.normal(size=n) * (1 + 0.5 * np.abs(X[:, 0]))
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    cl1, cl2 = rng.integers(0, 500, n).astype(np.int32), rng.integers(0, 30, n).astype(np.int32)
    reg = O.run_regression(y, X, w, n, absorbed_df=17)
    cols = [dev(y)] + [dev(X[:, j]) for j in range(k)]
    bf, B = dev(reg["beta_full"]), dev(reg["B"])
    dw = dev(w) if weighted else None
    args = (X, reg["B"], reg["resid"], w)
    out = eng.resid_vcov(cols, bf, B, 0, None, None, dw, True, n, reg["df_resid"], want_resid=True)
    np.testing.assert_allclose(out["resid"].cpu().numpy(), reg["resid"], rtol=0, atol=1e-12)
    st = out["stats"].cpu().numpy()
    assert st[0] == pytest.approx(reg["rss"], rel=1e-12) and st[1] == pytest.approx(reg["tss"], rel=1e-11)
    np.testing.assert_allclose(out["se"].cpu().numpy(), O.se_iid(reg["B"], reg["resid"], w, reg["df_resid"]), rtol=1e-11)
    out = eng.resid_vcov(cols, bf, B, 1, None, None, dw, True, n, reg["df_resid"])
    np.testing.assert_allclose(out["se"].cpu().numpy(), O.se_hc1(*args, n, reg["df_resid"])[0], rtol=1e-10)
    for ssc in (True, False):
        out = eng.resid_vcov(cols, bf, B, 2, [dev(cl1)], [500], dw, ssc, n, reg["df_resid"])
        se, G, V = O.se_cluster_oneway(*args, cl1, n, reg["df_resid"], ssc)
        np.testing.assert_allclose(out["se"].cpu().numpy(), se, rtol=1e-10)
        np.testing.assert_allclose(out["vcov"].cpu().numpy(), V, rtol=1e-9, atol=1e-18)
        assert out["n_clusters"] == [G]
        out = eng.resid_vcov(cols, bf, B, 2, [dev(cl1), dev(cl2)], [500, 30], dw, ssc, n, reg["df_resid"])
        se, Gs, _ = O.se_cluster_multiway(*args, [cl1, cl2], n, reg["df_resid"], ssc)
        np.testing.assert_allclose(out["se"].cpu().numpy(), se, rtol=1e-9)
        assert tuple(out["n_clusters"]) == Gs
    # sorted cluster codes take the segmented path: same answer
    p = np.argsort(cl1, kind="stable")
    cols_s = [dev(y[p])] + [dev(X[p, j]) for j in range(k)]
    out = eng.resid_vcov(cols_s, bf, B, 2, [dev(cl1[p])], [500], dev(w[p]) if weighted else None, True, n, reg["df_resid"])
    np.testing.assert_allclose(out["se"].cpu().numpy(), O.se_cluster_oneway(*args, cl1, n, reg["df_resid"], True)[0], rtol=1e-10)


# ------------------------------------------------------------------------------------------------
# (3) YOCO compression: keys and counts bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("weighted", [False, True])
def test_compress_bit_exact_keys_and_counts(eng, weighted):
    rng = np.random.default_rng(21)
    n = 300_000
    keys = [rng.integers(0, 2, n), rng.integers(0, 3, n), rng.integers(0, 50, n) * 3 - 7, rng.integers(0, 12, n)]
    y = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    exp = O.compress(y, keys, w)
    kc, Gs = [], []
    for kcol in keys:
        c, G, _ = eng.factorize(dev(kcol))
        kc.append(c)
        Gs.append(G)
    got = eng.compress(kc, Gs, dev(y), dev(w) if weighted else None)
    assert got["n_compressed"] == exp["n_compressed"]
    for a, b in zip(got["key_codes"], exp["key_codes"]):
        np.testing.assert_array_equal(a.cpu().numpy(), b)                                  # bit-exact, sorted order
    cnt = np.bincount(np.unique(np.column_stack(exp["key_codes"]), axis=0, return_inverse=True)[1].reshape(-1)) * 0
    if not weighted:
        np.testing.assert_array_equal(got["cnt"].cpu().numpy(), exp["n"].astype(np.int64))  # bit-exact counts
    np.testing.assert_allclose(got["n"].cpu().numpy(), exp["n"], rtol=1e-13)
    np.testing.assert_allclose(got["sum_y"].cpu().numpy(), exp["sum_y"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(got["sum_y_sq"].cpu().numpy(), exp["sum_y_sq"], rtol=1e-11)
    r = eng.compress(kc, Gs, None, count_only=True)
    assert r["n_compressed"] == exp["n_compressed"]
    # checksum of checksums: totals are preserved by the group-by
    assert float(got["sum_y"].sum()) == pytest.approx(float((y * (w if weighted else 1.0)).sum()), rel=1e-9, abs=1e-8)
    assert int(got["cnt"].sum()) == n


def test_compress_known_answer_from_reference_docstring(eng):
    # compress.py:318-323: 3 rows -> 2 groups
    kc, Gs = [], []
    for col in (np.array([1, 1, 2]), np.array([1, 1, 2])):
        c, G, _ = eng.factorize(dev(col))
        kc.append(c)
        Gs.append(G)
    got = eng.compress(kc, Gs, dev(np.array([1.0, 2.0, 3.0])))
    assert got["n_compressed"] == 2
    np.testing.assert_array_equal(got["cnt"].cpu().numpy(), [2, 1])
    np.testing.assert_allclose(got["sum_y"].cpu().numpy(), [3.0, 3.0])
    np.testing.assert_allclose(got["sum_y_sq"].cpu().numpy(), [5.0, 9.0])


@pytest.mark.parametrize("weighted", [False, True])
@pytest.mark.parametrize("n,g1,g2", [(700_003, 100, 70), (2_000_000, 300, 100)])
def test_compress_midsize_group_counts_use_smem_windows(eng, weighted, n, g1, g2):
    """A few 10^4 groups with many rows each: shared-memory histogram + windowed shared-memory statistics
    (k_hist_smem_big / k_group_stats_smem) against the oracle; 1 to 4 windows."""
    rng = np.random.default_rng(n + g1)
    keys = [rng.integers(0, g1, n), rng.integers(0, g2, n)]
    y = rng.normal(size=n) + 3.0
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    exp = O.compress(y, keys, w)
    kc, Gs = zip(*[eng.factorize(dev(kcol))[:2] for kcol in keys])
    got = eng.compress(list(kc), list(Gs), dev(y), dev(w) if weighted else None)
    assert got["n_compressed"] == exp["n_compressed"] == g1 * g2
    for a, b_ in zip(got["key_codes"], exp["key_codes"]):
        np.testing.assert_array_equal(a.cpu().numpy(), b_)
    if not weighted:
        np.testing.assert_array_equal(got["cnt"].cpu().numpy(), exp["n"].astype(np.int64))
        np.testing.assert_array_equal(got["n"].cpu().numpy(), exp["n"])                    # exact integers
    np.testing.assert_allclose(got["n"].cpu().numpy(), exp["n"], rtol=1e-12)
    np.testing.assert_allclose(got["sum_y"].cpu().numpy(), exp["sum_y"], rtol=1e-11)
    np.testing.assert_allclose(got["sum_y_sq"].cpu().numpy(), exp["sum_y_sq"], rtol=1e-11)


@pytest.mark.parametrize("n", [5, 4099, 250_001])
def test_compress_raw_uint8_keys_match_factorized_keys(eng, n):
    """uint8 key columns grouped on their raw values (no dictionary pass) give the same groups, counts and
    sums as the same columns factorised first; key_ranges supplies the radix."""
    rng = np.random.default_rng(n)
    d = [rng.integers(0, 2, n).astype(np.uint8) for _ in range(12)]
    d.append((rng.integers(0, 3, n) * 2).astype(np.uint8))            # values {0, 2, 4}: a gap in the raw code space
    ids = rng.integers(0, 37, n)
    y = rng.normal(size=n)
    d_dev = [dev(c) for c in d]
    rng_ = eng.key_ranges(d_dev + [dev(ids), dev(ids.astype(np.int32) - 5)])
    assert rng_ == [(int(c.min()), int(c.max())) for c in d] + [(int(ids.min()), int(ids.max())),
                                                                  (int(ids.min()) - 5, int(ids.max()) - 5)]
    ic, iG, _ = eng.factorize(dev(ids))
    raw = eng.compress(d_dev + [ic], [hi + 1 for _, hi in rng_[:13]] + [iG], dev(y))
    fk, fG, fu = [], [], []
    for c in d_dev:
        cc, G, u = eng.factorize(c, want_uniques=True)
        fk.append(cc); fG.append(G); fu.append(u)
    ref = eng.compress(fk + [ic], fG + [iG], dev(y))
    exp = O.compress(y, d + [ids], None)
    assert raw["n_compressed"] == ref["n_compressed"] == exp["n_compressed"]
    assert torch.equal(raw["cnt"], ref["cnt"])
    for j in range(13):                                                 # raw key out = the value itself
        assert torch.equal(raw["key_codes"][j].long(), fu[j].long()[ref["key_codes"][j].long()])
    assert torch.equal(raw["key_codes"][13], ref["key_codes"][13])
    np.testing.assert_array_equal(raw["cnt"].cpu().numpy(), exp["n"].astype(np.int64))
    np.testing.assert_allclose(raw["sum_y"].cpu().numpy(), exp["sum_y"], rtol=1e-10, atol=1e-10)
    # misaligned views take the scalar path
    off = eng.compress([c[1:] for c in d_dev] + [ic[1:].contiguous()], [hi + 1 for _, hi in rng_[:13]] + [iG], dev(y[1:]))
    exp1 = O.compress(y[1:], [c[1:] for c in d] + [ids[1:]], None)
    assert off["n_compressed"] == exp1["n_compressed"]
    np.testing.assert_array_equal(off["cnt"].cpu().numpy(), exp1["n"].astype(np.int64))


@pytest.mark.parametrize("weighted", [False, True])
def test_compress_merge_of_partial_tables_equals_whole(eng, weighted):
    """lfe_compress_merge (multi-GPU YOCO, here on one device): group-by of the concatenated per-shard tables
    == group-by of all rows; keys and counts bit-exact, sums to rounding."""
    rng = np.random.default_rng(33)
    n = 200_000
    keys = [rng.integers(0, 2, n), rng.integers(0, 40, n), rng.integers(0, 9, n)]
    keys[0][: n // 3] = 0                                   # a key value absent from the first shard
    y = rng.normal(size=n)
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    kc, Gs = [], []
    for kcol in keys:
        c, G, _ = eng.factorize(dev(kcol))
        kc.append(c)
        Gs.append(G)
    whole = eng.compress(kc, Gs, dev(y), dev(w) if weighted else None)
    cuts = [0, n // 3, n // 3 + 17, n]                      # ragged shards, one tiny
    parts = [eng.compress([c[a:b_].contiguous() for c in kc], Gs, dev(y[a:b_]), dev(w[a:b_]) if weighted else None)
             for a, b_ in zip(cuts[:-1], cuts[1:])]
    cat = dict(n_compressed=sum(p["n_compressed"] for p in parts),
               key_codes=[torch.cat([p["key_codes"][j] for p in parts]) for j in range(len(kc))],
               **{f: torch.cat([p[f] for p in parts]) for f in ("n", "cnt", "sum_y", "sum_y_sq")})
    got = eng.compress_merge(cat, Gs)
    assert got["n_compressed"] == whole["n_compressed"]
    for a, b_ in zip(got["key_codes"], whole["key_codes"]):
        assert torch.equal(a, b_)
    assert torch.equal(got["cnt"], whole["cnt"])
    for f in ("n", "sum_y", "sum_y_sq"):
        np.testing.assert_allclose(got[f].cpu().numpy(), whole[f].cpu().numpy(), rtol=1e-12, atol=1e-11)
    again = eng.compress_merge(got, Gs)                     # idempotent on a table without duplicate keys
    assert again["n_compressed"] == got["n_compressed"] and torch.equal(again["cnt"], got["cnt"])
    with eng.replicated() as e2:                            # no communicator attached: a no-op scope
        assert e2.world == 1


# ------------------------------------------------------------------------------------------------
# edge cases: tiny / ragged inputs, plan reuse, many columns, every row its own group
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 257, 1025])
def test_tiny_inputs_every_kernel(eng, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, max(n // 3, 1), n).astype(np.int32)
    b = rng.integers(0, 3, n).astype(np.int32)
    cols = [rng.normal(size=n) for _ in range(3)]
    Ga, Gb = int(a.max()) + 1, 3
    exp_cols, _, _ = O.alt_proj(cols, [a, b], [1, 0] if Gb <= Ga else [0, 1], 0.0, 4)
    d_cols = [dev(c) for c in cols]
    eng.demean(d_cols, [dev(a), dev(b)], [Ga, Gb], [1, 0] if Gb <= Ga else [0, 1], None, 0.0, 4, 3)
    for got, exp in zip(d_cols, exp_cols):
        np.testing.assert_allclose(got.cpu().numpy(), np.nan_to_num(exp), rtol=0, atol=1e-12)
    codes, G, uniq = eng.factorize(dev(a.astype(np.int64) * 5 - 2), want_uniques=True)
    np.testing.assert_array_equal(codes.cpu().numpy(), O.codes_sorted(a)[0])
    perm, _ = eng.sort_rows(dev(a), Ga)
    np.testing.assert_array_equal(perm.cpu().numpy(), np.argsort(a, kind="stable"))
    got = eng.gram([dev(c) for c in cols], None, True).cpu().numpy()
    Z = np.column_stack([np.ones(n)] + cols)
    np.testing.assert_allclose(got, Z.T @ Z, rtol=1e-12, atol=1e-12)
    comp = eng.compress([dev(a), dev(b)], [Ga, Gb], dev(cols[0]))
    exp = O.compress(cols[0], [a, b])
    assert comp["n_compressed"] == exp["n_compressed"]
    np.testing.assert_array_equal(comp["cnt"].cpu().numpy(), exp["n"].astype(np.int64))


def test_demean_plan_reuse_and_many_columns(eng):
    rng = np.random.default_rng(77)
    n, Gs = 80_000, (4000, 25)
    codes = [rng.integers(0, g, n).astype(np.int32) for g in Gs]
    w = rng.uniform(0.5, 2.0, n)
    plan = eng.demean_plan([dev(c) for c in codes], list(Gs), dev(w))
    for C in (1, 9, 20):                                   # 20 columns = three column batches
        cols = [rng.normal(size=n) + 0.01 * codes[0] for _ in range(C)]
        exp_cols, exp_it, _ = O.alt_proj(cols, codes, [1, 0], 0.0, 5, w)
        d_cols = [dev(c) for c in cols]
        it, _, _ = plan.run(d_cols, [1, 0], 0.0, 5, 3)
        assert it == exp_it
        for got, exp in zip(d_cols, exp_cols):
            np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=0, atol=1e-11)
    plan.close()


def test_every_row_its_own_group_and_one_group(eng):
    rng = np.random.default_rng(5)
    n = 40_000
    a = rng.permutation(n).astype(np.int32)                # G == n
    b = np.zeros(n, dtype=np.int32)                        # G == 1
    cols = [rng.normal(size=n) for _ in range(2)]
    d_cols = [dev(c) for c in cols]
    it, _, conv = eng.demean(d_cols, [dev(a), dev(b)], [n, 1], [1, 0], None, 1e-6, 50, 3)
    for got in d_cols:                                      # each row is its own group: everything is absorbed
        assert float(got.abs().max()) < 1e-12
    keep, n_keep, dims = eng.singleton_mask([dev(a), dev(b)], [n, 1])
    assert n_keep == 0 and dims == (0, 0) and int(keep.sum()) == 0


def test_demean_plan_out_of_place_matches_in_place(eng):
    """lfe_demean_plan_run with cols_src: inputs untouched, outputs identical to the in-place run."""
    rng = np.random.default_rng(11)
    n = 40_000
    Gs = (900, 17)
    codes = [dev(rng.integers(0, g, n).astype(np.int32)) for g in Gs]
    src = [dev(rng.normal(size=n)) for _ in range(3)]
    keep = [t.clone() for t in src]
    plan = eng.demean_plan(codes, list(Gs))
    a = [t.clone() for t in src]
    it_a = plan.run(a, [1, 0], 1e-10, 100, 3)
    b = [torch.full_like(t, float("nan")) for t in src]
    it_b = plan.run(b, [1, 0], 1e-10, 100, 3, src=src)
    plan.close()
    assert it_a[0] == it_b[0] and it_a[2] == it_b[2]
    for x, y_, s0, k0 in zip(a, b, src, keep):
        assert torch.equal(s0, k0)
        assert float((x - y_).abs().max()) < 1e-12        # fp64 atomics: summation order differs run to run


# ------------------------------------------------------------------------------------------------
# lfe_demean_plan_fit: the fused tail (x~ formed in registers) equals the materialised chain
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,Gs,weighted,n,offset", [
    (6, (5000, 60), False, 200_003, 0),      # rows-per-thread 4, n % 4 != 0
    (6, (5000, 60), True, 200_000, 0),
    (3, (3000, 700, 12), False, 150_001, 0),  # odd narrow tables (no 128-bit gathers), three FEs
    (8, (4000, 30), False, 120_002, 0),      # rows-per-thread 2
    (11, (4000, 500), True, 90_001, 0),      # rows-per-thread 1, padded table rows (ld 12)
    (1, (300, 7), False, 10_007, 0),         # no regressors
    (5, (2000, 40), False, 50_003, 1),       # unaligned column views -> scalar path
    (4, (800,), False, 30_000, 0),           # single FE (within transform)
])
def test_plan_fit_matches_run_gram_solve_vcov(eng, C, Gs, weighted, n, offset):
    rng = np.random.default_rng(C * 1000 + len(Gs))
    k = C - 1
    fe = [rng.integers(0, g, n).astype(np.int32) for g in Gs]
    X = rng.normal(size=(n, k)) + 0.3 * rng.normal(size=Gs[0])[fe[0]][:, None]
    y = X @ (1 + 0.1 * np.arange(k)) + rng.normal(size=Gs[0])[fe[0]] + rng.normal(size=n) * (1 + 0.3 * (fe[-1] % 3))
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    cl = rng.integers(0, 400, n).astype(np.int32)

    def col(a):
        if not offset:
            return dev(a)
        buf = torch.empty(n + offset, dtype=torch.float64, device="cuda")
        buf[offset:] = torch.from_numpy(a).cuda()
        return buf[offset:]

    raw = [col(y)] + [col(np.ascontiguousarray(X[:, j])) for j in range(k)]
    codes = [dev(f) for f in fe]
    dw = dev(w) if weighted else None
    order = sorted(range(len(Gs)), key=lambda j: Gs[j])
    F = len(Gs)
    mi, mn = (1, 2) if F == 1 else (50, 3)
    plan = eng.demean_plan(codes, list(Gs), dw)
    # the materialised chain
    dm = [torch.empty_like(t) for t in raw]
    it_ref = plan.run(dm, order, 1e-8, mi, mn, src=raw)
    gram = eng.gram(dm[1:] + [dm[0]], dw, add_intercept=True)
    p = k + 1
    beta_full, Ainv, _ = eng.solve_spd(gram[:p, :p].contiguous(), gram[:p, p].contiguous())
    df = n - p - (sum(Gs) - F)
    for kind, clc, G in ((0, None, 0), (1, None, 0), (2, dev(cl), 400)):
        out = plan.fit(raw, order, 1e-8, mi, mn, kind, clc, G, True, n, df)
        assert out["iterations"] == it_ref[0] and out["converged"] == it_ref[2]
        np.testing.assert_allclose(out["beta_full"].cpu().numpy()[1:], beta_full.cpu().numpy()[1:], rtol=1e-10, atol=1e-14)
        st = out["stats"].cpu().numpy()
        if k:
            B = Ainv[1:, 1:].contiguous()
            ref = eng.resid_vcov(dm, beta_full, B, kind, [clc] if clc is not None else None, [G] if G else None, dw,
                                 True, n, df)
            np.testing.assert_allclose(out["se"].cpu().numpy(), ref["se"].cpu().numpy(), rtol=1e-9)
            np.testing.assert_allclose(out["vcov"].cpu().numpy(), ref["vcov"].cpu().numpy(), rtol=1e-8, atol=1e-20)
            rst = ref["stats"].cpu().numpy()
            assert st[0] == pytest.approx(rst[0], rel=1e-11) and st[1] == pytest.approx(rst[1], rel=1e-9)
            if kind == 2:
                assert out["n_clusters"] == ref["n_clusters"][0] == len(np.unique(cl))
        else:
            r = (dm[0] - beta_full[0]).cpu().numpy()
            assert st[0] == pytest.approx(float((r ** 2).sum()), rel=1e-11)
    # the raw columns were only read
    np.testing.assert_array_equal(raw[0].cpu().numpy(), y)
    plan.close()


@pytest.mark.parametrize("Gs,C,weighted", [((300_000, 400), 6, False), ((200_000, 3_000, 25), 6, False),
                                            ((150_000, 2_000), 4, True), ((70_000, 60_000), 11, False)])
def test_demean_l2_blocked_entry_lists_match_oracle(eng, Gs, C, weighted):
    """LFE_OPT_COLUMNS_HINT + a small LFE_OPT_L2_BLOCK_BYTES force the (block of the big code, own code, big code)
    list order: groups are split into several runs, the sums must not change."""
    from leanfe_b200 import _capi as capi
    rng = np.random.default_rng(sum(Gs) + C)
    n = 400_000
    fe = [rng.integers(0, g, n).astype(np.int32) for g in Gs]
    cols = [rng.normal(size=n) + 0.2 * rng.normal(size=Gs[0])[fe[0]] for _ in range(C)]
    w = rng.uniform(0.5, 2.0, n) if weighted else None
    order = sorted(range(len(Gs)), key=lambda j: Gs[j])
    exp, _, _ = O.alt_proj(cols, fe, order, 0.0, 4, w)                           # exactly 4 sweeps
    eng.set_option(capi.OPT_L2_BLOCK_BYTES, 1 << 20)
    try:
        plan = eng.demean_plan([dev(f) for f in fe], list(Gs), dev(w) if weighted else None, columns_hint=C)
        got = [dev(c) for c in cols]
        plan.run(got, order, 0.0, 4, 3)
        plan.close()
    finally:
        eng.set_option(capi.OPT_L2_BLOCK_BYTES, 48 << 20)
    for a, b in zip(got, exp):
        np.testing.assert_allclose(a.cpu().numpy(), b, rtol=0, atol=2e-12)


def test_hard_limits_at_the_edge(eng):
    """Documented limits of the C ABI are enforced at the edge, not one past it: lfe_solve_spd solves p = 512 and
    rejects 513; row counts must stay below 2^31 - 1 (codes and permutations are 32-bit)."""
    import ctypes as C
    from leanfe_b200 import _capi as capi
    from leanfe_b200._capi import LfeError
    rng = np.random.default_rng(512)
    p = 512
    M = rng.normal(size=(3 * p, p))
    A, b = M.T @ M + 1e-3 * np.eye(p), rng.normal(size=p)
    beta, Ainv, fb = eng.solve_spd(dev(A), dev(b))
    assert not fb
    np.testing.assert_allclose(beta.cpu().numpy(), np.linalg.solve(A, b), rtol=1e-7)
    np.testing.assert_allclose(Ainv.cpu().numpy() @ A, np.eye(p), atol=1e-8)
    with pytest.raises(LfeError, match="out of range"):
        eng.solve_spd(dev(np.eye(513)), dev(np.ones(513)))
    # n = 2^31 - 1 rows: refused before any memory is touched (the pointers are never dereferenced)
    tiny = torch.zeros(8, dtype=torch.int64, device="cuda")
    G = C.c_int64(0)
    rc = eng.lib.lfe_factorize(eng.h, C.c_void_p(tiny.data_ptr()), capi.LFE_I64, (1 << 31) - 1, 0,
                               C.c_void_p(tiny.data_ptr()), C.byref(G), None, 0, eng.stream)
    assert rc == -1 and b"n out of range" in eng.lib.lfe_last_error()
    it, mm, cv = C.c_int(0), C.c_double(0), C.c_int(0)
    rc = eng.lib.lfe_demean(eng.h, capi.ptr_array([tiny.data_ptr()]), 1, capi.ptr_array([tiny.data_ptr()]),
                            capi.i64_array([4]), capi.i32_array([0]), 1, 1 << 31, None, 1e-6, 5, 3, -1, 0,
                            C.byref(it), C.byref(mm), C.byref(cv), eng.stream)
    assert rc == -1 and b"out of range" in eng.lib.lfe_last_error()

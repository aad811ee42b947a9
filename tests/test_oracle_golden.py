"""CPU: the NumPy oracle against the committed golden vectors.

'ref_*' fixtures come from the UNMODIFIED reference functions (run under stub imports by
tests/golden/make_golden.py); 'dense_*' from an independent dense dummy-variable least squares.
"""
import json
import os

import numpy as np
import pytest

from oracle import ref_loader as RL
from oracle import ref_numpy as O

RTOL_BETA, RTOL_SE = 1e-8, 1e-6      # north_star tolerances


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def test_formula_grammar_matches_reference(golden_dir):
    J = json.load(open(os.path.join(golden_dir, "structural.json")))
    for formula, exp in J["parse_formula"]:
        got = O.parse_formula(formula)
        assert [list(map(list, v)) if isinstance(v, list) and v and isinstance(v[0], tuple) else v for v in got] == exp
    for formula, err in J["parse_formula_errors"]:
        if err is None:
            O.parse_formula(formula)
        else:
            with pytest.raises(ValueError):
                O.parse_formula(formula)


def test_strategy_rule_matches_reference(golden_dir):
    J = json.load(open(os.path.join(golden_dir, "structural.json")))
    for kwargs, exp in J["determine_strategy"]:
        assert O.determine_strategy(**kwargs) == exp


def test_twoway_against_dense_ols_and_reference_sandwich(golden_dir):
    g = _load(golden_dir, "twoway.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "unit", "time", "w")}
    for tag, kw in (("", {}), ("_w", dict(weights="w"))):
        r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit + time", vcov="iid", demean_tol=1e-13, max_iter=500,
                            return_internals=True, **kw)
        beta = np.array(list(r["coefs"].values()))
        np.testing.assert_allclose(beta, g["dense_beta" + tag], rtol=RTOL_BETA)
        np.testing.assert_allclose(beta, g["tight_beta" + tag], rtol=1e-12)
        reg, X, keep, w = r["_reg"], np.column_stack(r["_cols"][1:]), r["_keep"], r["_w"]
        se = O.se_hc1(X, reg["B"], reg["resid"], w, r["n_obs"], r["df_resid"])[0]
        np.testing.assert_allclose(se, g["ref_se_hc1" + tag], rtol=1e-10)
        for ssc in (True, False):
            se, G, _ = O.se_cluster_oneway(X, reg["B"], reg["resid"], w, d["unit"][keep], r["n_obs"], r["df_resid"], ssc)
            np.testing.assert_allclose(se, g[f"ref_se_cl_unit_ssc{int(ssc)}{tag}"], rtol=1e-10)
            assert G == int(g["ref_G_unit"])
            se, Gs, _ = O.se_cluster_multiway(X, reg["B"], reg["resid"], w, [d["unit"][keep], d["time"][keep]],
                                              r["n_obs"], r["df_resid"], ssc)
            np.testing.assert_allclose(se, g[f"ref_se_cl_2way_ssc{int(ssc)}{tag}"], rtol=1e-10)
            assert tuple(Gs) == tuple(g["ref_G_2way"])
    r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit + time", vcov="iid", demean_tol=1e-13, max_iter=500)
    np.testing.assert_allclose(np.array(list(r["std_errors"].values())), g["dense_se_iid"], rtol=RTOL_SE)


def test_twoway_default_tolerance_is_pinned(golden_dir):
    g = _load(golden_dir, "twoway.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "unit", "time")}
    for vc, cl in (("iid", None), ("hc1", None), ("cluster", ["unit"]), ("cluster", ["unit", "time"])):
        tag = vc + ("_" + "_".join(cl) if cl else "")
        r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit + time", vcov=vc, cluster_cols=cl)
        np.testing.assert_allclose(np.array(list(r["coefs"].values())), g[f"oracle_beta_{tag}"], rtol=1e-13)
        np.testing.assert_allclose(np.array(list(r["std_errors"].values())), g[f"oracle_se_{tag}"], rtol=1e-12)
        assert r["iterations"] == int(g[f"oracle_iterations_{tag}"])
        assert r["df_resid"] == int(g[f"oracle_df_resid_{tag}"])
        # default tolerance already agrees with the exact (dense) solution far inside the north_star band
        np.testing.assert_allclose(np.array(list(r["coefs"].values())), g["dense_beta"], rtol=1e-5)


def test_threeway_singletons_and_dense(golden_dir):
    g = _load(golden_dir, "threeway.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "worker", "firm", "year")}
    r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | worker + firm + year", vcov="iid", demean_tol=1e-13, max_iter=5000)
    assert r["n_obs"] == int(g["oracle_n_obs"]) < d["y"].shape[0]          # singletons were dropped
    np.testing.assert_allclose(np.array(list(r["coefs"].values())), g["dense_beta"], rtol=RTOL_BETA)
    assert r["df_resid"] == r["n_obs"] - int(g["dense_rank"]) or r["df_resid"] <= r["n_obs"]


def test_yoco_against_reference_solver(golden_dir):
    g = _load(golden_dir, "yoco.npz")
    d = {k: g[k] for k in ("y", "xd", "x2", "grp", "per", "st")}
    for vc, ssc, cl in (("iid", True, None), ("HC1", True, None), ("cluster", True, ["grp"]),
                        ("cluster", False, ["grp"]), ("cluster", True, ["st"]), ("cluster", True, ["grp", "st"])):
        tag = f"{vc.lower()}_ssc{int(ssc)}" + ("_" + "_".join(cl) if cl else "")
        r = O.leanfe_oracle(d, formula="y ~ xd + x2 | grp + per", strategy="compress", vcov=vc, cluster_cols=cl, ssc=ssc)
        np.testing.assert_allclose(np.array(list(r["coefs"].values())), g[f"ref_beta_{tag}"], rtol=1e-9)
        np.testing.assert_allclose(np.array(list(r["std_errors"].values())), g[f"ref_se_{tag}"], rtol=1e-8)
        assert r["df_resid"] == int(g[f"ref_df_resid_{tag}"])
        np.testing.assert_allclose(r["rss"], float(g[f"ref_rss_{tag}"]), rtol=1e-10)
        assert r["n_compressed"] == int(g[f"n_compressed_{tag}"])


def test_yoco_equals_fwl_identity(golden_dir):
    """FWL on the raw rows and YOCO on the compressed rows are the same estimator (IID)."""
    g = _load(golden_dir, "yoco.npz")
    d = {k: g[k] for k in ("y", "xd", "x2", "grp", "per")}
    a = O.leanfe_oracle(d, formula="y ~ xd + x2 | grp + per", strategy="compress", vcov="iid")
    b = O.leanfe_oracle(d, formula="y ~ xd + x2 | grp + per", strategy="alt_proj", vcov="iid", demean_tol=1e-13, max_iter=500)
    np.testing.assert_allclose(list(a["coefs"].values()), list(b["coefs"].values()), rtol=1e-9)
    np.testing.assert_allclose(list(a["std_errors"].values()), list(b["std_errors"].values()), rtol=1e-9)


IV_CASES = [  # (tag, strategy, weights, vcov, cluster_cols, ssc) -- formulas are stored in the fixture
    ("exact_iid", "alt_proj", None, "iid", None, True), ("exact2_hc1", "alt_proj", None, "hc1", None, True),
    ("exact2_cl", "alt_proj", None, "cluster", ["unit"], True),
    ("exact2_cl_nossc_w", "alt_proj", "w", "cluster", ["unit"], False),
    ("exact2_2way", "alt_proj", None, "cluster", ["unit", "st"], True),
    ("exact_iid_w_onefe", "demean", "w", "iid", None, True), ("nofe_hc1", "ols", None, "hc1", None, True),
    ("nofe_over_iid", "ols", None, "iid", None, True),
]


def test_iv_2sls_against_reference_functions(golden_dir):
    """IV: the restatement against common.iv_2sls + std_errors._compute_se_*_iv of the unmodified reference
    (tests/golden/make_golden_iv.py)."""
    g = _load(golden_dir, "iv.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "z1", "z2", "unit", "time", "w", "st")}
    for tag, strategy, wname, vc, cl, ssc in IV_CASES:
        r = O.leanfe_oracle(d, formula=str(g[f"formula_{tag}"]), strategy=strategy, weights=wname, vcov=vc,
                            cluster_cols=cl, ssc=ssc, demean_tol=1e-13, max_iter=2000)
        np.testing.assert_allclose(np.array(list(r["coefs"].values())), g[f"ref_beta_{tag}"], rtol=1e-9)
        np.testing.assert_allclose(np.array(list(r["std_errors"].values())), g[f"ref_se_{tag}"], rtol=1e-8)
        assert r["df_resid"] == int(g[f"ref_df_resid_{tag}"]) and r["is_iv"] and r["r_squared"] is None
        np.testing.assert_allclose(r["rss"], float(g[f"ref_rss_{tag}"]), rtol=1e-9)
    with pytest.raises(ValueError, match="Under-identified"):
        O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit | z1", strategy="demean")        # 2 instruments (1 + z1) < 3


def test_reference_docstring_known_answers():
    # compress.py:222-225 ; compress.py:318-323 ; compress.py:835-840
    assert O.estimate_compression_ratio([np.array([1, 1, 2]), np.array([1, 1, 2])]) == pytest.approx(2 / 3)
    c = O.compress(np.array([1.0, 2.0, 3.0]), [np.array([1, 1, 2]), np.array([1, 1, 2])])
    assert c["n_compressed"] == 2 and c["n_obs"] == 3
    np.testing.assert_array_equal(c["n"], [2, 1])
    np.testing.assert_allclose(c["sum_y"], [3.0, 3.0])
    np.testing.assert_allclose(c["sum_y_sq"], [5.0, 9.0])
    codes, uq, cnt = O.codes_sorted(np.array(["A", "A", "B", "B", "C", "A", "C"]))
    assert len(uq) == 3 and sorted(cnt.tolist()) == [2, 2, 3]


def test_group_code_conventions():
    v = np.array([5, 3, 5, 9, 3, 3])
    c, uq, cnt = O.codes_sorted(v)
    np.testing.assert_array_equal(c, [1, 0, 1, 2, 0, 0])
    c1, uq1 = O.codes_first_appearance(v)
    np.testing.assert_array_equal(c1, [0, 1, 0, 2, 1, 1])
    f = np.array([0.0, -0.0, np.nan, 1.0, np.nan])
    c, uq, _ = O.codes_sorted(f)
    assert len(uq) == 3 and c[0] == c[1] and c[2] == c[4]


@pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")
def test_oracle_against_live_reference_functions():
    """Where /root/reference exists: run the reference's own NumPy code next to the restatement."""
    ref = RL.load_reference()
    rng = np.random.default_rng(3)
    N = 4000
    d = dict(y=rng.normal(size=N), a=rng.integers(0, 3, N), b=rng.integers(0, 2, N), f=rng.integers(0, 15, N),
             c=rng.integers(0, 25, N))
    d["y"] = d["y"] + 0.3 * d["a"] - 0.2 * d["b"] + 0.05 * d["f"]
    for vc, cl in (("iid", None), ("HC1", None), ("cluster", ["c"])):
        r = O.leanfe_oracle(d, formula="y ~ a + b | f", strategy="compress", vcov=vc, cluster_cols=cl, return_internals=True)
        rr = RL.reference_yoco(ref, r["_comp"], ["a", "b"], ["f"], vc, True, cl)
        np.testing.assert_allclose(list(r["coefs"].values()), rr["beta"], rtol=1e-10)
        np.testing.assert_allclose(list(r["std_errors"].values()), rr["se"], rtol=1e-9)
    for f in ("y ~ x1 + i(region, ref=B) + t:i(g) | a + b", "y ~ x | fe | z1 + z2"):
        assert tuple(ref.common.parse_formula(f)) == tuple(O.parse_formula(f))


@pytest.mark.parametrize("weighted", [False, True])
def test_window_semantics_against_a_dataframe_engine(weighted):
    """The one part of the path no reference code can pin here (Polars is not installable) is the window
    arithmetic `col - col.mean().over(fe)` / `len().over(fe) > 1` (polars_impl.py:478-508).  pandas'
    groupby-transform is the same window semantics in an independent engine: the oracle's singleton filter,
    Gauss-Seidel sweeps (FE order, weighted means) and y-only unweighted stopping test must agree with the
    reference loop written literally against pandas."""
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(7)
    n = 6000
    df = pd.DataFrame({"unit": rng.integers(0, 900, n), "firm": rng.integers(0, 60, n), "year": rng.integers(0, 7, n)})
    df["unit"] = np.where(rng.random(n) < 0.02, 10_000 + np.arange(n), df["unit"])      # some singleton units
    for c in ("y", "x1", "x2"):
        df[c] = rng.normal(size=n) + 0.3 * (df["unit"] % 13) + 0.2 * df["firm"] + 0.1 * df["year"]
    df["w"] = rng.uniform(0.5, 2.0, n)
    fes, cols = ["unit", "firm", "year"], ["y", "x1", "x2"]

    # --- the reference loop, literally (polars_impl.py:478-526), on pandas -------------------------
    card = {f: df[f].nunique() for f in fes}                                  # :371-373, pre-drop
    keep = np.ones(n, dtype=bool)
    for f in fes:                                                             # :478-482, evaluated once
        keep &= (df.groupby(f)[f].transform("size") > 1).to_numpy()
    d = df[keep].reset_index(drop=True)
    order = sorted(fes, key=lambda f: card[f])                                # :485
    tol, max_iter, iters = 1e-8, 50, 0
    for it in range(1, max_iter + 1):
        for f in order:                                                       # :490-508
            for c in cols:
                if weighted:
                    num = (d[c] * d["w"]).groupby(d[f]).transform("sum")
                    den = d["w"].groupby(d[f]).transform("sum")
                    d[c] = d[c] - num / den
                else:
                    d[c] = d[c] - d.groupby(f)[c].transform("mean")
        iters = it
        if it >= 3:                                                           # :511-525, y only, unweighted
            mm = max(float(d.groupby(f)["y"].transform("mean").abs().max()) for f in fes)
            if mm < tol:
                break

    # --- the oracle ---------------------------------------------------------------------------------
    codes = [O.codes_sorted(df[f].to_numpy())[0] for f in fes]
    k_or = O.singleton_keep_mask(codes)
    np.testing.assert_array_equal(k_or, keep)
    kc = [c[k_or] for c in codes]
    o_order = O.fe_order([card[f] for f in fes])
    assert [fes[j] for j in o_order] == order
    o_cols, o_it, o_mm = O.alt_proj([df[c].to_numpy()[k_or] for c in cols], kc, o_order, tol, max_iter,
                                    df["w"].to_numpy()[k_or] if weighted else None)
    assert o_it == iters
    for c, got in zip(cols, o_cols):
        np.testing.assert_allclose(got, d[c].to_numpy(), rtol=0, atol=1e-11)

"""GPU: the public ``leanfe()`` API against the oracle, the committed goldens and the BASELINE configs.

Tolerances are the north_star's: max relative error 1e-8 on beta, 1e-6 on SEs; iteration counts,
n_obs, df_resid, fe_dims, n_clusters and YOCO group counts must be equal.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ref_numpy as O  # noqa: E402

RTOL_BETA, RTOL_SE = 1e-8, 1e-6


def _check(res, ref, rtol_beta=RTOL_BETA, rtol_se=RTOL_SE):
    assert list(res.coefs) == list(ref["coefs"])
    b = np.array(list(res.coefs.values()))
    eb = np.array(list(ref["coefs"].values()))
    s = np.array(list(res.std_errors.values()))
    es = np.array(list(ref["std_errors"].values()))
    assert np.max(np.abs(b - eb) / np.abs(eb)) < rtol_beta, (b, eb)
    assert np.max(np.abs(s - es) / np.abs(es)) < rtol_se, (s, es)
    assert res.n_obs == ref["n_obs"] and res.df_resid == ref["df_resid"]
    assert res.iterations == ref["iterations"]
    if ref.get("fe_dims") is not None:
        assert tuple(res.fe_dims) == tuple(ref["fe_dims"])
    assert res.n_clusters == ref["n_clusters"]
    assert res.strategy == ref["strategy"]


def _golden(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


@pytest.mark.parametrize("vcov,cl", [("iid", None), ("hc1", None), ("HC1", None), ("cluster", ["unit"]),
                                     ("cluster", ["time"]), ("cluster", ["unit", "time"])])
@pytest.mark.parametrize("weighted", [False, True])
def test_twoway_golden(golden_dir, vcov, cl, weighted):
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "twoway.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "unit", "time", "w")}
    kw = dict(formula="y ~ x1 + x2 | unit + time", vcov=vcov, cluster_cols=cl, weights="w" if weighted else None)
    res = leanfe(d, **kw)
    _check(res, O.leanfe_oracle(d, **kw))
    if not weighted and vcov == "iid":
        np.testing.assert_allclose(list(res.coefs.values()), g["oracle_beta_iid"], rtol=RTOL_BETA)
        assert res.iterations == int(g["oracle_iterations_iid"])
    # tight tolerance -> the exact dense dummy-variable solution and the reference's own sandwich code
    res = leanfe(d, demean_tol=1e-13, max_iter=500, **kw)
    tag = "_w" if weighted else ""
    np.testing.assert_allclose(list(res.coefs.values()), g["dense_beta" + tag], rtol=RTOL_BETA)
    se = np.array(list(res.std_errors.values()))
    if not weighted:          # weighted runs never satisfy the reference's unweighted stopping rule: both hit max_iter
        assert res.iterations == int(g["tight_iterations"])
    if vcov.lower() == "hc1":
        np.testing.assert_allclose(se, g["ref_se_hc1" + tag], rtol=RTOL_SE)
    if cl == ["unit"]:
        np.testing.assert_allclose(se, g["ref_se_cl_unit_ssc1" + tag], rtol=RTOL_SE)
        assert res.n_clusters == int(g["ref_G_unit"])
    if cl == ["unit", "time"]:
        np.testing.assert_allclose(se, g["ref_se_cl_2way_ssc1" + tag], rtol=RTOL_SE)
        assert res.n_clusters == tuple(g["ref_G_2way"])
    if vcov == "iid" and not weighted:
        np.testing.assert_allclose(se, g["dense_se_iid"], rtol=RTOL_SE)


def test_threeway_golden_with_singletons(golden_dir):
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "threeway.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "worker", "firm", "year")}
    kw = dict(formula="y ~ x1 + x2 | worker + firm + year", vcov="hc1")
    res = leanfe(d, **kw)
    ref = O.leanfe_oracle(d, **kw)
    _check(res, ref)
    assert res.n_obs == int(g["oracle_n_obs"]) and tuple(res.fe_dims) == tuple(g["oracle_fe_dims"])
    np.testing.assert_allclose(list(res.coefs.values()), g["oracle_beta"], rtol=RTOL_BETA)
    np.testing.assert_allclose(list(res.std_errors.values()), g["oracle_se"], rtol=RTOL_SE)
    res = leanfe(d, demean_tol=1e-13, max_iter=5000, formula=kw["formula"], vcov="iid")
    np.testing.assert_allclose(list(res.coefs.values()), g["dense_beta"], rtol=RTOL_BETA)
    assert res.iterations == int(g["tight_iterations"])


@pytest.mark.parametrize("vcov,ssc,cl", [("iid", True, None), ("HC1", True, None), ("cluster", True, ["grp"]),
                                         ("cluster", False, ["grp"]), ("cluster", True, ["st"]),
                                         ("cluster", True, ["grp", "st"])])
def test_yoco_golden_reference_solver(golden_dir, vcov, ssc, cl):
    """strategy='compress' against the reference's own build_design_matrix/solve_wls/compute_se_compress."""
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "yoco.npz")
    d = {k: g[k] for k in ("y", "xd", "x2", "grp", "per", "st")}
    tag = f"{vcov.lower()}_ssc{int(ssc)}" + ("_" + "_".join(cl) if cl else "")
    res = leanfe(d, formula="y ~ xd + x2 | grp + per", strategy="compress", vcov=vcov, cluster_cols=cl, ssc=ssc)
    np.testing.assert_allclose(list(res.coefs.values()), g[f"ref_beta_{tag}"], rtol=RTOL_BETA)
    np.testing.assert_allclose(list(res.std_errors.values()), g[f"ref_se_{tag}"], rtol=RTOL_SE)
    assert res.df_resid == int(g[f"ref_df_resid_{tag}"])
    assert res.n_compressed == int(g[f"n_compressed_{tag}"])
    assert res.rss == pytest.approx(float(g[f"ref_rss_{tag}"]), rel=1e-9)
    ncl = g[f"ref_ncl_{tag}"]
    if cl is not None:
        assert res.n_clusters == (int(ncl) if ncl.ndim == 0 else tuple(int(v) for v in ncl))
    assert res.strategy == "compress" and res.iterations == 0 and res.fe_dims == (40, 8)


IV_CASES = [  # (tag, strategy, weights, vcov, cluster_cols, ssc)
    ("exact_iid", "alt_proj", None, "iid", None, True), ("exact2_hc1", "alt_proj", None, "hc1", None, True),
    ("exact2_cl", "alt_proj", None, "cluster", ["unit"], True),
    ("exact2_cl_nossc_w", "alt_proj", "w", "cluster", ["unit"], False),
    ("exact2_2way", "alt_proj", None, "cluster", ["unit", "st"], True),
    ("exact_iid_w_onefe", "demean", "w", "iid", None, True), ("nofe_hc1", "ols", None, "hc1", None, True),
    ("nofe_over_iid", "ols", None, "iid", None, True),
]


@pytest.mark.parametrize("tag,strategy,wname,vcov,cl,ssc", IV_CASES)
def test_iv_2sls_golden_reference_functions(golden_dir, tag, strategy, wname, vcov, cl, ssc):
    """IV/2SLS (formula 'y ~ x | fe | z') against the reference's own iv_2sls / _compute_se_*_iv outputs."""
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "iv.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "z1", "z2", "unit", "time", "w", "st")}
    res = leanfe(d, formula=str(g[f"formula_{tag}"]), strategy=strategy, weights=wname, vcov=vcov, cluster_cols=cl,
                 ssc=ssc, demean_tol=1e-13, max_iter=2000)
    np.testing.assert_allclose(list(res.coefs.values()), g[f"ref_beta_{tag}"], rtol=RTOL_BETA)
    np.testing.assert_allclose(list(res.std_errors.values()), g[f"ref_se_{tag}"], rtol=RTOL_SE)
    assert res.df_resid == int(g[f"ref_df_resid_{tag}"])
    assert res.rss == pytest.approx(float(g[f"ref_rss_{tag}"]), rel=1e-8)
    assert res.is_iv and res.n_instruments == str(g[f"formula_{tag}"]).split("|")[2].count("+") + 1
    assert res.r_squared is None
    ncl = g[f"ref_ncl_{tag}"]
    if cl is not None:
        assert res.n_clusters == (int(ncl) if ncl.ndim == 0 else tuple(int(v) for v in ncl))


def test_iv_default_tolerance_matches_oracle_and_rejects_bad_requests(golden_dir):
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "iv.npz")
    d = {k: g[k] for k in ("y", "x1", "x2", "z1", "z2", "unit", "time", "w", "st")}
    kw = dict(formula="y ~ x1 + x2 | unit + time | z1 + x2", vcov="cluster", cluster_cols=["unit"])
    res, ref = leanfe(d, **kw), O.leanfe_oracle(d, **kw)          # strategy='auto' -> alt_proj, default tol / max_iter
    _check(res, ref)
    with pytest.raises(ValueError, match="Under-identified"):
        leanfe(d, formula="y ~ x1 + x2 | unit | z1")
    with pytest.raises(ValueError, match="instruments"):
        leanfe(d, formula="y ~ x1 | unit + time | z1", strategy="compress")


def _c1(balanced: bool, N=1_000_000):
    """BASELINE config 1: 1M rows, y ~ x1 + x2 | unit(10k) + time(100), IID."""
    rng = np.random.default_rng(42)
    if balanced:
        unit, time_ = np.repeat(np.arange(10_000), 100), np.tile(np.arange(100), 10_000)
    else:
        unit, time_ = rng.integers(0, 10_000, N), rng.integers(0, 100, N)
    a, g = rng.normal(size=10_000), rng.normal(size=100)
    x1, x2 = rng.normal(size=N) + 0.5 * a[unit], rng.normal(size=N) + 0.5 * a[unit]
    y = 1.0 * x1 + 0.5 * x2 + a[unit] + g[time_] + rng.normal(size=N)
    return dict(y=y, x1=x1, x2=x2, unit=unit, time=time_)


@pytest.mark.parametrize("balanced", [True, False])
def test_config1_1m_rows(balanced):
    from leanfe_b200 import leanfe
    d = _c1(balanced)
    kw = dict(formula="y ~ x1 + x2 | unit + time", vcov="iid")
    res = leanfe(d, **kw)                                    # strategy='auto' -> alt_proj (max FE = 10k... rule 3)
    ref = O.leanfe_oracle(d, strategy="alt_proj", **kw)
    assert res.strategy == "alt_proj"
    _check(res, ref)
    assert res.iterations == (3 if balanced else 4)
    assert res.r_squared == pytest.approx(ref["r_squared"], rel=1e-9)


def test_config2_shape_small_clustered_by_unit():
    """BASELINE config 2 at 1/25 scale: 5 regressors | unit + time, clustered by unit, heteroskedastic."""
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(42)
    N, G1, G2, k = 2_000_000, 40_000, 500, 5
    unit, time_ = rng.integers(0, G1, N), rng.integers(0, G2, N)
    a, g, s = rng.normal(size=G1), rng.normal(size=G2), rng.uniform(0.5, 1.5, G1)
    d = dict(unit=unit, time=time_)
    y = a[unit] + g[time_] + s[unit] * rng.normal(size=N)
    for j in range(1, k + 1):
        d[f"x{j}"] = rng.normal(size=N) + 0.5 * a[unit]
        y = y + (1 + 0.1 * j) * d[f"x{j}"]
    d["y"] = y
    kw = dict(formula="y ~ x1 + x2 + x3 + x4 + x5 | unit + time", vcov="cluster", cluster_cols=["unit"])
    res = leanfe(d, **kw)
    _check(res, O.leanfe_oracle(d, **kw))
    assert res.n_clusters == G1 and res.converged


def test_factor_and_interaction_expansion_names_and_values():
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(5)
    n = 20_000
    region = rng.choice(np.array(["A", "B", "C"], dtype=object), n)
    treat = rng.integers(0, 2, n).astype(np.float64)
    d = dict(y=rng.normal(size=n) + (region == "B") * 0.5 + treat * (region == "C") * 0.7, x=rng.normal(size=n),
             region=region, treatment=treat, fe1=rng.integers(0, 10, n), fe2=rng.integers(0, 5, n),
             period=rng.integers(0, 4, n))
    res = leanfe(d, formula="y ~ x + i(region) | fe1 + fe2")
    assert list(res.coefs) == ["x", "region_B", "region_C"]                # test_factor_variables.py:22-25
    _check(res, O.leanfe_oracle(d, formula="y ~ x + i(region) | fe1 + fe2"))
    res = leanfe(d, formula="y ~ x + i(region, ref=B) | fe1 + fe2")
    assert list(res.coefs) == ["x", "region_A", "region_C"]                # test_reference_category.py:43-45
    res = leanfe(d, formula="y ~ x + treatment:i(region) | fe1 + fe2", vcov="cluster", cluster_cols=["fe1"])
    assert list(res.coefs) == ["x", "treatment_B", "treatment_C"] and res.n_clusters == 10
    _check(res, O.leanfe_oracle(d, formula="y ~ x + treatment:i(region) | fe1 + fe2", vcov="cluster", cluster_cols=["fe1"]))
    res = leanfe(d, formula="y ~ x + treatment:i(period, ref=2) | fe1")
    assert list(res.coefs) == ["x", "treatment_0", "treatment_1", "treatment_3"]
    with pytest.raises(ValueError, match="Reference category"):
        leanfe(d, formula="y ~ x + i(region, ref=Z) | fe1")


def test_strategies_demean_ols_auto():
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(6)
    n = 50_000
    d = dict(y=rng.normal(size=n), x1=rng.normal(size=n), x2=rng.normal(size=n), f=rng.integers(0, 300, n),
             c=rng.integers(0, 77, n))
    d["y"] = d["y"] + 0.4 * d["x1"] - 0.2 * d["x2"] + 0.01 * d["f"]
    for vc, cl in (("iid", None), ("hc1", None), ("cluster", ["c"])):
        res = leanfe(d, formula="y ~ x1 + x2 | f", vcov=vc, cluster_cols=cl)          # auto -> 'demean'
        ref = O.leanfe_oracle(d, formula="y ~ x1 + x2 | f", vcov=vc, cluster_cols=cl)
        assert res.strategy == "demean" and res.iterations == 1
        _check(res, ref)
        res = leanfe(d, formula="y ~ x1 + x2", vcov=vc, cluster_cols=cl)               # auto -> 'ols'
        ref = O.leanfe_oracle(d, formula="y ~ x1 + x2", vcov=vc, cluster_cols=cl)
        assert res.strategy == "ols" and res.compression_ratio == pytest.approx(1.0)
        _check(res, ref)
    # discrete design, low-cardinality FEs: the cost model picks YOCO (compress.py:164-181)
    dd = dict(y=d["y"], a=rng.integers(0, 2, n), b=rng.integers(0, 3, n), f=rng.integers(0, 20, n), g=rng.integers(0, 6, n))
    res = leanfe(dd, formula="y ~ a + b | f + g")
    ref = O.leanfe_oracle(dd, formula="y ~ a + b | f + g")
    assert res.strategy == ref["strategy"] == "compress"
    assert res.compression_ratio == pytest.approx(ref["compression_ratio"] if ref["compression_ratio"] else
                                                  O.estimate_compression_ratio([dd[c] for c in "abfg"]))
    np.testing.assert_allclose(list(res.coefs.values()), list(ref["coefs"].values()), rtol=RTOL_BETA)
    np.testing.assert_allclose(list(res.std_errors.values()), list(ref["std_errors"].values()), rtol=RTOL_SE)
    res = leanfe(dd, formula="y ~ a + b", strategy="compress", vcov="HC1")             # YOCO without FEs
    ref = O.leanfe_oracle(dd, formula="y ~ a + b", strategy="compress", vcov="HC1")
    np.testing.assert_allclose(list(res.coefs.values()), list(ref["coefs"].values()), rtol=RTOL_BETA)
    np.testing.assert_allclose(list(res.std_errors.values()), list(ref["std_errors"].values()), rtol=RTOL_SE)


def test_input_kinds_agree(tmp_path):
    import pandas as pd
    import pyarrow as pa
    import pyarrow.parquet as pq
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(7)
    n = 30_000
    d = dict(y=rng.normal(size=n), x=rng.normal(size=n), f=rng.integers(0, 50, n), g=rng.integers(0, 7, n))
    base = leanfe(d, formula="y ~ x | f + g")
    path = str(tmp_path / "d.parquet")
    pq.write_table(pa.table(d), path)
    variants = [pd.DataFrame(d), pa.table(d), path,
                {k: torch.from_numpy(v) for k, v in d.items()},
                {k: torch.from_numpy(v).cuda() for k, v in d.items()}]
    for v in variants:
        r = leanfe(v, formula="y ~ x | f + g")
        assert r.coefs["x"] == pytest.approx(base.coefs["x"], rel=1e-12)
        assert r.std_errors["x"] == pytest.approx(base.std_errors["x"], rel=1e-12)
    # caller data is never modified, string ids work
    before = d["y"].copy()
    ds = dict(d, f=np.array([f"id{v:03d}" for v in d["f"]], dtype=object))
    r = leanfe(ds, formula="y ~ x | f + g")
    assert r.coefs["x"] == pytest.approx(base.coefs["x"], rel=1e-12)
    np.testing.assert_array_equal(d["y"], before)


def test_result_dict_contract():
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(8)
    n = 5000
    d = dict(y=rng.normal(size=n), x=rng.normal(size=n), fe1=rng.integers(0, 10, n), fe2=rng.integers(0, 5, n))
    r = leanfe(d, formula="y ~ x | fe1 + fe2", profile=True)
    for key in ("formula", "coefs", "std_errors", "t_stats", "p_values", "n_obs", "n_compressed", "compression_ratio",
                "fe_cols", "fe_dims", "iterations", "vcov_type", "is_iv", "n_instruments", "n_clusters", "df_resid",
                "r_squared_within"):
        assert key in r.keys()                                     # result.py:248-268
    assert "x" in r["coefficients"] and r["n_obs"] > 0              # test_basic.py:22-25 spelling
    assert r["vcov_type"] == "iid" and "leanfe Regression Results" in str(r)
    assert r.timings["demean_sweeps"] > 0 and r.timings["launches"] > 0


def test_degenerate_inputs_raise_cleanly():
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(9)
    n = 1000
    d = dict(y=rng.normal(size=n), x=rng.normal(size=n), u=np.arange(n), t=rng.integers(0, 5, n))
    with pytest.raises(ValueError, match="singleton"):
        leanfe(d, formula="y ~ x | u + t")                  # every unit is a singleton: nothing left to fit
    with pytest.raises(KeyError):
        leanfe(d, formula="y ~ nope | t")
    with pytest.raises(KeyError):
        leanfe(d, formula="y ~ x | t | z")                  # unknown instrument column


def test_wide_design_twenty_regressors():
    """Shape of the reference's 20-regressor benchmark (python/tests/benchmark_results2.csv:29), small N."""
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(10)
    n, k = 60_000, 20
    d = dict(f1=rng.integers(0, 800, n), f2=rng.integers(0, 150, n), f3=rng.integers(0, 40, n))
    y = rng.normal(size=n)
    for j in range(k):
        d[f"x{j}"] = rng.normal(size=n) + 0.02 * d["f1"] / 800
        y = y + (0.5 + 0.05 * j) * d[f"x{j}"]
    d["y"] = y
    f = "y ~ " + " + ".join(f"x{j}" for j in range(k)) + " | f1 + f2 + f3"
    for vc, cl in (("iid", None), ("hc1", None), ("cluster", ["f2"])):
        res = leanfe(d, formula=f, vcov=vc, cluster_cols=cl)
        _check(res, O.leanfe_oracle(d, formula=f, vcov=vc, cluster_cols=cl))


def test_pageable_ingest_ring_is_bit_exact_for_every_source_kind():
    """lfe_uploader_* (pinned ring + staging threads): the bytes that arrive are the bytes that were sent -- columns
    larger than the whole ring (8 x 32 MB), odd sizes, read-only arrays, narrow dtypes, unpinned tensors -- and the
    fit from such sources equals the fit from device tensors."""
    from leanfe_b200 import leanfe
    from leanfe_b200.engine import get_engine
    from leanfe_b200.frontend import Frame, Uploader
    eng = get_engine(0)
    rng = np.random.default_rng(3)
    big = rng.integers(-2**62, 2**62, 40_000_003)                      # 320 MB > 256 MB of ring
    ro = rng.normal(size=1_000_001)
    ro.setflags(write=False)
    cols = {"big": big, "ro": ro, "u8": rng.integers(0, 255, 70_001).astype(np.uint8),
            "f32": rng.normal(size=33).astype(np.float32), "i32": rng.integers(-9, 9, 5_000_000).astype(np.int32),
            "unpinned": torch.from_numpy(rng.normal(size=2_000_003)), "strided": rng.normal(size=(100_000, 3))[:, 1],
            "pinned": torch.from_numpy(rng.normal(size=300_001)).pin_memory(), "empty": np.zeros(0)}
    up = Uploader(eng, Frame(cols, list(cols)), list(cols))
    for name, src in cols.items():
        got = up.get(name)
        exp = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.array(src)))
        assert got.is_cuda and got.dtype == exp.dtype and torch.equal(got.cpu(), exp), name
    up.close()
    # the public API on pageable / read-only inputs
    n = 400_000
    d = dict(y=rng.normal(size=n), x=rng.normal(size=n), f=rng.integers(0, 5000, n), g=rng.integers(0, 7, n))
    base = leanfe({k: torch.from_numpy(v).cuda() for k, v in d.items()}, formula="y ~ x | f + g", vcov="cluster",
                  cluster_cols=["f"])
    for v in d.values():
        v.setflags(write=False)
    for _ in range(3):                                                  # ring slots are reused across fits
        r = leanfe(d, formula="y ~ x | f + g", vcov="cluster", cluster_cols=["f"])
        assert r.coefs["x"] == pytest.approx(base.coefs["x"], rel=1e-12)
        assert r.std_errors["x"] == pytest.approx(base.std_errors["x"], rel=1e-12)


@pytest.mark.parametrize("vcov,cl", [("iid", None), ("HC1", None), ("cluster", ["grp"])])
def test_yoco_singular_design_takes_the_references_min_norm_fallback(golden_dir, vcov, cl):
    """A regressor that is identically zero makes the normal equations singular: the reference's Cholesky raises and
    it returns np.linalg.lstsq / pinv numbers (compress.py:742-745).  Golden = the reference's own solve_wls +
    compute_se_compress on that design (tests/golden/make_golden.py)."""
    from leanfe_b200 import leanfe
    g = _golden(golden_dir, "yoco_singular.npz")
    d = {k: g[k] for k in ("y", "xd", "x0", "x2", "grp", "per", "st")}
    tag = f"{vcov.lower()}_ssc1" + ("_" + "_".join(cl) if cl else "")
    with pytest.warns(RuntimeWarning, match="collinear"):
        res = leanfe(d, formula="y ~ xd + x0 + x2 | grp + per", strategy="compress", vcov=vcov, cluster_cols=cl)
    eb, es = g[f"ref_beta_{tag}"], g[f"ref_se_{tag}"]
    b, s = np.array(list(res.coefs.values())), np.array(list(res.std_errors.values()))
    for j in (0, 2):
        assert abs(b[j] - eb[j]) / abs(eb[j]) < RTOL_BETA and abs(s[j] - es[j]) / es[j] < RTOL_SE
    assert abs(b[1]) < 1e-10 and abs(eb[1]) < 1e-10 and s[1] < 1e-12 and es[1] < 1e-12       # the dead column
    assert res.df_resid == int(g[f"ref_df_resid_{tag}"])
    assert res.rss == pytest.approx(float(g[f"ref_rss_{tag}"]), rel=1e-9)


@pytest.mark.parametrize("fes,vcov,cl,weighted", [("unit + time", "cluster", ["unit"], False), ("unit + time", "hc1", None, True),
                                                   ("unit + time", "cluster", ["unit", "time"], False),
                                                   ("unit + time + reg", "iid", None, False), ("unit", "cluster", ["time"], True)])
def test_deterministic_mode_is_bit_identical_across_runs(fes, vcov, cl, weighted):
    """deterministic=True (LFE_OPT_DETERMINISTIC): every group sum is taken in a fixed order, so two fits of the same
    data agree in every bit -- and still match the oracle to the usual tolerances."""
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(21)
    n = 300_000
    unit, time_, reg = rng.integers(0, 20_000, n), rng.integers(0, 40, n), rng.integers(0, 300, n)
    a = rng.normal(size=20_000)
    d = dict(unit=unit, time=time_, reg=reg, w=rng.uniform(0.5, 2.0, n))
    d["x1"], d["x2"], d["x3"] = (rng.normal(size=n) + 0.4 * a[unit] for _ in range(3))
    d["y"] = d["x1"] - 0.5 * d["x2"] + 0.25 * d["x3"] + a[unit] + 0.05 * time_ + rng.normal(size=n) * (1 + 0.3 * (unit % 4))
    kw = dict(formula=f"y ~ x1 + x2 + x3 | {fes}", vcov=vcov, cluster_cols=cl, weights="w" if weighted else None,
              strategy="alt_proj" if "+" in fes else "demean")
    dd = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    runs = [leanfe(dd, deterministic=True, **kw) for _ in range(3)]
    for r in runs[1:]:
        assert list(r.coefs.values()) == list(runs[0].coefs.values())               # exact equality
        assert list(r.std_errors.values()) == list(runs[0].std_errors.values())
        assert r.rss == runs[0].rss and r.iterations == runs[0].iterations
    _check(runs[0], O.leanfe_oracle(d, **kw))
    # the default (atomics) path agrees to rounding
    fast = leanfe(dd, **kw)
    np.testing.assert_allclose(list(fast.coefs.values()), list(runs[0].coefs.values()), rtol=1e-11)


@pytest.mark.parametrize("vcov,cl,weighted", [("iid", None, False), ("hc1", None, True), ("cluster", ["grp"], False),
                                              ("cluster", ["grp", "per"], False)])
def test_deterministic_mode_yoco_is_bit_identical_across_runs(vcov, cl, weighted):
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(33)
    n = 500_000
    grp, per = rng.integers(0, 300, n), rng.integers(0, 12, n)
    d = dict(grp=grp, per=per, d1=((grp % 3 == 0) & (per >= 6)).astype(np.uint8), dose=rng.integers(0, 4, n).astype(np.float64),
             w=rng.integers(1, 4, n).astype(np.float64))
    d["y"] = 0.8 * d["d1"] + 0.1 * d["dose"] + rng.normal(size=300)[grp] + 0.05 * per + rng.normal(size=n)
    kw = dict(formula="y ~ d1 + dose | grp + per", strategy="compress", vcov=vcov, cluster_cols=cl,
              weights="w" if weighted else None)
    dd = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    runs = [leanfe(dd, deterministic=True, **kw) for _ in range(3)]
    for r in runs[1:]:
        assert list(r.coefs.values()) == list(runs[0].coefs.values())
        assert list(r.std_errors.values()) == list(runs[0].std_errors.values())
        assert r.rss == runs[0].rss and r.n_compressed == runs[0].n_compressed
    ref = O.leanfe_oracle(d, **kw)
    np.testing.assert_allclose(list(runs[0].coefs.values()), list(ref["coefs"].values()), rtol=RTOL_BETA)
    np.testing.assert_allclose(list(runs[0].std_errors.values()), list(ref["std_errors"].values()), rtol=RTOL_SE)


def test_live_reference_when_polars_is_present():
    """The only route from 'parity partially pinned' to fully pinned: when the reference's engine (Polars) is importable
    on the box, the UNMODIFIED reference `leanfe(..., backend="polars")` (installed --no-deps under baseline/_ref, or
    /root/reference/python) is run on the same data and must agree to the north_star tolerances.  Skips -- and says
    why -- where Polars is absent (this image)."""
    import sys
    try:
        import polars as pl
    except Exception as e:       # noqa: BLE001
        pytest.skip(f"import polars failed ({type(e).__name__}): the reference's engine is not installed here")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.path.join(root, "baseline", "_ref"), "/root/reference/python"):
        if os.path.isdir(os.path.join(cand, "leanfe")) and cand not in sys.path:
            sys.path.insert(0, cand)
    try:
        from leanfe import leanfe as ref_leanfe
    except Exception as e:       # noqa: BLE001
        pytest.skip(f"polars is importable but the reference package is not ({type(e).__name__}: {e})")
    from leanfe_b200 import leanfe
    rng = np.random.default_rng(5)
    n = 200_000
    unit, time_ = rng.integers(0, 5_000, n), rng.integers(0, 40, n)
    a = rng.normal(size=5_000)
    d = dict(unit=unit, time=time_, x1=rng.normal(size=n) + 0.4 * a[unit], x2=rng.normal(size=n))
    d["y"] = d["x1"] - 0.5 * d["x2"] + a[unit] + 0.05 * time_ + rng.normal(size=n)
    for vcov, cl in (("iid", None), ("cluster", ["unit"])):
        kw = dict(formula="y ~ x1 + x2 | unit + time", vcov=vcov, cluster_cols=cl, strategy="alt_proj")
        ref = ref_leanfe(pl.DataFrame(d), backend="polars", **kw)
        res = leanfe(d, **kw)
        for name in ("x1", "x2"):
            assert abs(res.coefs[name] - ref.coefs[name]) / abs(ref.coefs[name]) < RTOL_BETA
            assert abs(res.std_errors[name] - ref.std_errors[name]) / ref.std_errors[name] < RTOL_SE
        assert res.iterations == ref.iterations and res.n_obs == ref.n_obs

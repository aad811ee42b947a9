"""Generate the committed golden fixtures in tests/golden/.

Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Every "ref_*" entry is produced by the UNMODIFIED reference code imported
under stubbed polars/duckdb (oracle/ref_loader.py).  Every "dense_*" entry is an
independent dense dummy-variable least squares.  "oracle_*" entries pin the
NumPy restatement itself (regression guard; they are only as good as the two
former kinds, which the CPU tests cross-check).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as RL  # noqa: E402
from oracle import ref_numpy as O  # noqa: E402


def panel_twoway(seed=42, N=6000, G1=300, G2=12, k=2):
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, G1, N)
    time = rng.integers(0, G2, N)
    a, g = rng.normal(size=G1), rng.normal(size=G2)
    X = [rng.normal(size=N) + 0.5 * a[unit] for _ in range(k)]
    s = rng.uniform(0.5, 1.5, G1)
    y = sum((1 + 0.1 * (j + 1)) * X[j] for j in range(k)) + a[unit] + g[time] + s[unit] * rng.normal(size=N)
    w = rng.uniform(0.5, 2.0, N)
    d = dict(y=y, unit=unit * 7 + 3, time=time - 5, w=w)  # non-dense, shifted raw ids
    for j in range(k):
        d[f"x{j + 1}"] = X[j]
    return d


def panel_threeway(seed=7, N=5000):
    rng = np.random.default_rng(seed)
    worker = rng.integers(0, 1200, N)
    home = rng.integers(0, 150, 1200)
    firm = np.where(rng.random(N) < 0.8, home[worker], rng.integers(0, 150, N))
    year = rng.integers(0, 6, N)
    x1, x2 = rng.normal(size=N), rng.normal(size=N)
    aw, af = rng.normal(size=1200), rng.normal(size=150)
    y = 0.8 * x1 - 0.3 * x2 + aw[worker] + af[firm] + 0.1 * year + rng.normal(size=N)
    return dict(y=y, x1=x1 + 0.3 * aw[worker], x2=x2, worker=worker, firm=firm, year=year)


def panel_yoco(seed=11, N=20000):
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, 40, N)
    per = rng.integers(0, 8, N)
    xd = ((grp < 16) & (per >= 4)).astype(np.int64)
    x2 = rng.integers(0, 3, N)
    st = rng.integers(0, 9, N)
    y = 0.7 * xd + 0.2 * x2 + rng.normal(size=N) + 0.1 * grp + 0.3 * per
    return dict(y=y, xd=xd, x2=x2, grp=grp, per=per, st=st)


def main():
    ref = RL.load_reference()
    J = {}

    # ---- formula grammar (common.py:51-181) --------------------------------
    formulas = [
        "y ~ x1 + x2", "y ~ x1 + x2 | fe1 + fe2", "y ~ x1 + i(region) | fe1",
        "y ~ x1 + i(region, ref=R1) | fe1", "y ~ x1 + treatment:i(region) | fe1",
        "y ~ x1 + treatment:i(region, ref=\"R1\") | fe1 + fe2", "y ~ x1 + x2 | fe1 + fe2 | z1 + z2",
        "y~x", " y ~ x1+x2 |unit+ time ", "y ~ treated:i(period, ref=11) + x | group + period",
    ]
    J["parse_formula"] = [[f, [list(map(list, v)) if isinstance(v, list) and v and isinstance(v[0], tuple) else v
                               for v in ref.common.parse_formula(f)]] for f in formulas]
    bad = ["y ~ x | a | b | c", "y x1", "y ~ i(region | fe"]
    J["parse_formula_errors"] = []
    for f in bad:
        try:
            ref.common.parse_formula(f)
            J["parse_formula_errors"].append([f, None])
        except Exception as e:  # noqa: BLE001
            J["parse_formula_errors"].append([f, type(e).__name__])

    # ---- strategy rule (compress.py:96-184) --------------------------------
    cases = [
        dict(vcov="iid", has_instruments=False, fe_cardinality={"unit": 10000, "time": 100}, n_obs=1_000_000, n_x_cols=2, estimated_compression_ratio=1.0),
        dict(vcov="cluster", has_instruments=False, fe_cardinality={"g": 1000, "p": 24}, n_obs=200_000_000, n_x_cols=23, estimated_compression_ratio=1.2e-4),
        dict(vcov="iid", has_instruments=True, fe_cardinality={"g": 10}),
        dict(vcov="HC1", has_instruments=False, fe_cardinality={"a": 10001, "b": 5}),
        dict(vcov="hc1", has_instruments=False, fe_cardinality={"a": 9000, "b": 9000, "c": 4000}),
        dict(vcov="iid", has_instruments=False, fe_cardinality=None),
        dict(vcov="weird", has_instruments=False, fe_cardinality={"a": 3}),
        dict(vcov="iid", has_instruments=False, fe_cardinality={"a": 500, "b": 100}, n_obs=100_000, n_x_cols=3, estimated_compression_ratio=0.15),
        dict(vcov="iid", has_instruments=False, fe_cardinality={"a": 500, "b": 100}, n_obs=100_000, n_x_cols=3, estimated_compression_ratio=0.001),
        dict(vcov="iid", has_instruments=False, fe_cardinality={"a": 50, "b": 10}),
    ]
    J["determine_strategy"] = [[c, ref.compress.determine_strategy(**c)] for c in cases]

    # ---- LeanFEResult (result.py) -------------------------------------------
    R = ref.result.LeanFEResult(coefs={"x1": 1.25, "x2": -0.5}, std_errors={"x1": 0.1, "x2": 0.25}, n_obs=12345,
                                vcov_type="cluster", iterations=4, n_clusters=321, df_resid=12000,
                                formula="y ~ x1 + x2 | a + b", fe_cols=["a", "b"], fe_dims=(300, 12))
    J["result"] = dict(str=str(R), repr=repr(R), t_stats=R.t_stats, p_values=R.p_values,
                       confint={k: list(v) for k, v in R.confint().items()}, keys=list(R.keys()))

    with open(os.path.join(HERE, "structural.json"), "w") as f:
        json.dump(J, f, indent=1)

    # ---- two-way panel: sandwich algebra via the reference's own functions ---
    d = panel_twoway()
    out = dict(d)
    X = np.column_stack([d["x1"], d["x2"]])
    cu, ct = O.codes_sorted(d["unit"])[0], O.codes_sorted(d["time"])[0]
    b, rank, M = O.dense_dummy_ols(d["y"], X, [cu, ct])
    out["dense_beta"], out["dense_rank"] = b, rank
    bw, _, _ = O.dense_dummy_ols(d["y"], X, [cu, ct], d["w"])
    out["dense_beta_w"] = bw
    # dense IID SEs: sigma^2 (M'M)^-1
    res = d["y"] - M @ np.linalg.lstsq(M, d["y"], rcond=None)[0]
    s2 = res @ res / (len(res) - rank)
    out["dense_se_iid"] = np.sqrt(s2 * np.diag(np.linalg.inv(M.T @ M))[1:3])
    for tag, kw in (("", {}), ("_w", dict(weights="w"))):
        r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit + time", vcov="iid", demean_tol=1e-13, max_iter=500,
                            return_internals=True, **kw)
        reg, Xd, keep, w = r["_reg"], np.column_stack(r["_cols"][1:]), r["_keep"], r["_w"]
        out[f"tight_beta{tag}"] = reg["beta"]
        out[f"tight_se_iid{tag}"] = np.array(list(r["std_errors"].values()))
        out[f"tight_iterations{tag}"] = r["iterations"]
        S = ref.std_errors
        out[f"ref_se_hc1{tag}"] = S._compute_se_hc1_iv(reg["B"], reg["resid"], Xd, w, r["n_obs"], r["df_resid"])[0]
        for ssc in (True, False):
            se, G = S._compute_se_cluster_oneway_iv(reg["B"], reg["resid"], Xd, w, d["unit"][keep], r["n_obs"], r["df_resid"], ssc)
            out[f"ref_se_cl_unit_ssc{int(ssc)}{tag}"], out["ref_G_unit"] = se, G
            se, Gs = S._compute_se_cluster_multiway_iv(reg["B"], reg["resid"], Xd, w,
                                                       np.column_stack([d["unit"], d["time"]])[keep], r["n_obs"], r["df_resid"], ssc)
            out[f"ref_se_cl_2way_ssc{int(ssc)}{tag}"], out["ref_G_2way"] = se, np.array(Gs)
    # default-tolerance run (leanfe() defaults: tol 1e-6, max_iter 50) pins the iteration count
    for vc, cl in (("iid", None), ("hc1", None), ("cluster", ["unit"]), ("cluster", ["unit", "time"])):
        r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | unit + time", vcov=vc, cluster_cols=cl)
        tag = vc + ("_" + "_".join(cl) if cl else "")
        out[f"oracle_beta_{tag}"] = np.array(list(r["coefs"].values()))
        out[f"oracle_se_{tag}"] = np.array(list(r["std_errors"].values()))
        out[f"oracle_iterations_{tag}"] = r["iterations"]
        out[f"oracle_df_resid_{tag}"] = r["df_resid"]
    np.savez_compressed(os.path.join(HERE, "twoway.npz"), **out)

    # ---- three-way with singletons ------------------------------------------
    d = panel_threeway()
    out = dict(d)
    r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | worker + firm + year", vcov="hc1", return_internals=True)
    out.update(oracle_beta=np.array(list(r["coefs"].values())), oracle_se=np.array(list(r["std_errors"].values())),
               oracle_iterations=r["iterations"], oracle_n_obs=r["n_obs"], oracle_fe_dims=np.array(r["fe_dims"]),
               oracle_df_resid=r["df_resid"], keep=r["_keep"])
    r = O.leanfe_oracle(d, formula="y ~ x1 + x2 | worker + firm + year", vcov="iid", demean_tol=1e-13, max_iter=5000,
                        return_internals=True)
    keep = r["_keep"]
    Xk = np.column_stack([d["x1"], d["x2"]])[keep]
    codes = [O.codes_sorted(d[c][keep])[0] for c in ("worker", "firm", "year")]
    b, rank, M = O.dense_dummy_ols(d["y"][keep], Xk, codes)
    out.update(dense_beta=b, dense_rank=rank, tight_beta=r["_reg"]["beta"], tight_iterations=r["iterations"])
    np.savez_compressed(os.path.join(HERE, "threeway.npz"), **out)

    # ---- YOCO: the reference's own solve/SE code ------------------------------
    d = panel_yoco()
    out = dict(d)
    for vc, ssc, cl in (("iid", True, None), ("HC1", True, None), ("cluster", True, ["grp"]), ("cluster", False, ["grp"]),
                        ("cluster", True, ["st"]), ("cluster", True, ["grp", "st"])):
        r = O.leanfe_oracle(d, formula="y ~ xd + x2 | grp + per", strategy="compress", vcov=vc, cluster_cols=cl, ssc=ssc,
                            return_internals=True)
        rr = RL.reference_yoco(ref, r["_comp"], ["xd", "x2"], ["grp", "per"], vc, ssc, cl)
        tag = f"{vc.lower()}_ssc{int(ssc)}" + ("_" + "_".join(cl) if cl else "")
        out[f"ref_beta_{tag}"], out[f"ref_se_{tag}"] = rr["beta"], rr["se"]
        out[f"ref_df_resid_{tag}"], out[f"ref_rss_{tag}"] = rr["df_resid"], rr["rss"]
        out[f"ref_ncl_{tag}"] = np.array(rr["n_clusters"] if rr["n_clusters"] is not None else -1)
        out[f"n_compressed_{tag}"] = r["n_compressed"]
        c = r["_comp"]
        if cl is None and vc == "iid":
            for j, nm in enumerate(["xd", "x2", "grp", "per"]):
                out[f"comp_key_{nm}"] = c["keys"][j]
            out["comp_n"], out["comp_sum_y"], out["comp_sum_y_sq"] = c["n"], c["sum_y"], c["sum_y_sq"]
    np.savez_compressed(os.path.join(HERE, "yoco.npz"), **out)

    # ---- YOCO on a SINGULAR design: a regressor that is identically zero makes the reference's Cholesky raise, so
    # its lstsq / pinv fallback (compress.py:742-745) produces the numbers below --------------------------------
    d = panel_yoco()
    d["x0"] = np.zeros_like(d["y"])
    out = dict(d)
    for vc, ssc, cl in (("iid", True, None), ("HC1", True, None), ("cluster", True, ["grp"])):
        r = O.leanfe_oracle({k: v for k, v in d.items() if k != "x0"}, formula="y ~ xd + x2 | grp + per", strategy="compress",
                            vcov=vc, cluster_cols=cl, ssc=ssc, return_internals=True)
        comp = O.compress(d["y"], [d["xd"], d["x0"], d["x2"], d["grp"], d["per"]])
        comp["n_obs"] = len(d["y"])
        rr = RL.reference_yoco(ref, comp, ["xd", "x0", "x2"], ["grp", "per"], vc, ssc, cl)
        tag = f"{vc.lower()}_ssc{int(ssc)}" + ("_" + "_".join(cl) if cl else "")
        out[f"ref_beta_{tag}"], out[f"ref_se_{tag}"] = rr["beta"], rr["se"]
        out[f"ref_df_resid_{tag}"], out[f"ref_rss_{tag}"] = rr["df_resid"], rr["rss"]
    np.savez_compressed(os.path.join(HERE, "yoco_singular.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

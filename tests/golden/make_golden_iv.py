"""Generate tests/golden/iv.npz: IV/2SLS known answers from the UNMODIFIED reference functions.

Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden_iv.py

The panel is demeaned by the NumPy oracle (tight tolerance); the demeaned y / X / Z then go through the
reference's own ``common.iv_2sls`` (common.py:188-287) and ``std_errors._compute_se_{hc1,cluster_oneway,
cluster_multiway}_iv`` (std_errors.py:448-598), imported under stubbed polars/duckdb (oracle/ref_loader.py).
Only the glue of ``_run_regression`` (polars_impl.py:166-198, 229-232: intercept column, constant joining Z,
(Xh'Xh)^-1, residuals with the fitted Xh) is restated here, because it needs a Polars frame.  The IID formula
(std_errors.py:183-210) is one line of NumPy.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader as RL  # noqa: E402
from oracle import ref_numpy as O  # noqa: E402


def panel_iv(seed=5, N=8000, G1=250, G2=10):
    """x1 endogenous (correlated with the error through u), x2 exogenous; z1, z2 instruments for x1; x2 is
    listed among the instruments too so that the exogenous regressor instruments itself."""
    rng = np.random.default_rng(seed)
    unit, time = rng.integers(0, G1, N), rng.integers(0, G2, N)
    a, g = rng.normal(size=G1), rng.normal(size=G2)
    u = rng.normal(size=N)
    z1, z2 = rng.normal(size=N) + 0.3 * a[unit], rng.normal(size=N)
    x2 = rng.normal(size=N) + 0.2 * g[time]
    x1 = 0.8 * z1 + 0.4 * z2 + 0.5 * u + 0.3 * a[unit] + rng.normal(size=N)
    y = 1.5 * x1 - 0.7 * x2 + a[unit] + g[time] + u + rng.normal(size=N)
    w = rng.uniform(0.5, 2.0, N)
    st = rng.integers(0, 30, N)
    return dict(y=y, x1=x1, x2=x2, z1=z1, z2=z2, unit=unit * 3 + 1, time=time, w=w, st=st)


def reference_iv(ref, y, X, Z, w, n_obs, absorbed, vcov, clusters, ssc):
    """_run_regression's IV branch on demeaned arrays, heavy parts = the reference's own functions."""
    k = X.shape[1]
    Xi = np.hstack([np.ones((X.shape[0], 1)), X])
    Zi = Z
    if Xi.shape[1] > Zi.shape[1] and not any(np.allclose(c, 1.0) for c in Zi.T):
        Zi = np.column_stack([np.ones(Zi.shape[0]), Zi])
    beta_full, Xh = ref.common.iv_2sls(y, Xi, Zi, w)
    A = (Xh * np.sqrt(w)[:, None]).T @ (Xh * np.sqrt(w)[:, None]) if w is not None else Xh.T @ Xh
    L = np.linalg.cholesky(A)
    XtX_inv = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(L.shape[0])))
    resid = y - Xh @ beta_full
    df_resid = n_obs - (k + 1) - absorbed
    S = ref.std_errors
    ncl = None
    if vcov == "iid":
        s2 = np.sum((w if w is not None else 1.0) * resid ** 2) / df_resid
        se = np.sqrt(np.maximum(s2 * np.diag(XtX_inv), 0.0))
    elif vcov == "hc1":
        se, _ = S._compute_se_hc1_iv(XtX_inv, resid, Xh, w, n_obs, df_resid)
    elif len(clusters) == 1:
        se, ncl = S._compute_se_cluster_oneway_iv(XtX_inv, resid, Xh, w, clusters[0], n_obs, df_resid, ssc)
    else:
        se, ncl = S._compute_se_cluster_multiway_iv(XtX_inv, resid, Xh, w, np.column_stack(clusters), n_obs, df_resid, ssc)
    return beta_full[1:], se[1:], df_resid, ncl, float(np.sum(resid ** 2))


CASES = [  # (tag, x_cols, instruments, fe, weights, vcov, cluster_cols, ssc)
    ("exact_iid", ["x1"], ["z1"], "unit + time", None, "iid", None, True),
    ("exact2_hc1", ["x1", "x2"], ["z1", "x2"], "unit + time", None, "hc1", None, True),
    ("exact2_cl", ["x1", "x2"], ["z1", "x2"], "unit + time", None, "cluster", ["unit"], True),
    ("exact2_cl_nossc_w", ["x1", "x2"], ["z1", "x2"], "unit + time", "w", "cluster", ["unit"], False),
    ("exact2_2way", ["x1", "x2"], ["z1", "x2"], "unit + time", None, "cluster", ["unit", "st"], True),
    ("exact_iid_w_onefe", ["x1"], ["z1"], "unit", "w", "iid", None, True),
    ("nofe_hc1", ["x1", "x2"], ["z1", "x2"], "", None, "hc1", None, True),
    ("nofe_over_iid", ["x1"], ["z1", "z2"], "", None, "iid", None, True),
]


def main():
    ref = RL.load_reference()
    d = panel_iv()
    out = dict(d)
    for tag, xs, zs, fe, wname, vc, cl, ssc in CASES:
        formula = f"y ~ {' + '.join(xs)}" + (f" | {fe}" if fe else " | ") + f" | {' + '.join(zs)}"
        strategy = "alt_proj" if "+" in fe else ("demean" if fe else "ols")
        r = O.leanfe_oracle(d, formula=formula, strategy=strategy, weights=wname, vcov=vc, cluster_cols=cl, ssc=ssc,
                            demean_tol=1e-13, max_iter=2000, return_internals=True)
        cols, keep, w = r["_cols"], r["_keep"], r["_w"]
        k = len(xs)
        y, X, Z = cols[0], np.column_stack(cols[1:1 + k]), np.column_stack(cols[1 + k:])
        fe_dims = r["fe_dims"] or ()
        absorbed = (sum(fe_dims) - len(fe_dims)) if fe_dims else 0
        clusters = [np.asarray(d[c])[keep] for c in cl] if cl else None
        b, se, df, ncl, rss = reference_iv(ref, y, X, Z, w, r["n_obs"], absorbed, vc, clusters, ssc)
        out[f"ref_beta_{tag}"], out[f"ref_se_{tag}"] = b, se
        out[f"ref_df_resid_{tag}"], out[f"ref_rss_{tag}"] = df, rss
        out[f"ref_ncl_{tag}"] = np.array(ncl if ncl is not None else -1)
        out[f"formula_{tag}"] = formula
        # the restatement must agree with the reference's functions before it is trusted as the GPU tests' checker
        ob, os_ = np.array(list(r["coefs"].values())), np.array(list(r["std_errors"].values()))
        assert np.allclose(ob, b, rtol=1e-10, atol=0), (tag, ob, b)
        assert np.allclose(os_, se, rtol=1e-9, atol=0), (tag, os_, se)
        assert r["df_resid"] == df and (r["n_clusters"] == ncl or ncl is None)
        print(tag, formula, b, se, df, ncl)
    np.savez_compressed(os.path.join(HERE, "iv.npz"), **out)
    print("written", os.path.join(HERE, "iv.npz"))


if __name__ == "__main__":
    main()

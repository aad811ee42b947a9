"""Import the UNMODIFIED reference package under stubbed ``polars`` / ``duckdb``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Works only where
``/root/reference`` exists (the build container, not the GPU box), and is used
solely by ``tests/golden/make_golden.py`` to produce the committed fixtures
and by CPU tests that are skipped when the reference tree is absent.

The reference's arithmetic for the Polars-expression parts cannot run (Polars
1.36.1 is not installable here), but its NumPy/SciPy functions do run as is:
``common.parse_formula``, ``compress.determine_strategy``,
``compress.build_design_matrix`` (dict-of-arrays ``DuckDBResult`` branch),
``compress.solve_wls``, ``compress.compute_rss_grouped``,
``compress.compute_se_compress``, ``compress._build_sparse_cluster_matrix``,
``std_errors._compute_se_*_iv`` (same sandwich algebra as the Polars variants)
and ``result.LeanFEResult``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LEANFE_REFERENCE_ROOT", "/root/reference/python")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "leanfe"))


def _stub_modules():
    saved = {}
    pl = types.ModuleType("polars")

    class _Cfg:
        @staticmethod
        def set_engine_affinity(*a, **k):
            return None

    pl.DataFrame = type("DataFrame", (), {})
    pl.LazyFrame = type("LazyFrame", (), {})
    pl.Series = type("Series", (), {})
    pl.Config = _Cfg
    for name in ("String", "Categorical", "Float64", "UInt8", "UInt32"):
        setattr(pl, name, type(name, (), {}))
    sel = types.ModuleType("polars.selectors")
    pl.selectors = sel
    dd = types.ModuleType("duckdb")
    dd.DuckDBPyConnection = type("DuckDBPyConnection", (), {})
    for name, mod in (("polars", pl), ("polars.selectors", sel), ("duckdb", dd)):
        saved[name] = sys.modules.get(name)
        sys.modules[name] = mod
    return saved


def load_reference():
    """Return the reference's ``leanfe`` package (modules: common, compress, std_errors, result)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    saved = _stub_modules()
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        for m in [k for k in sys.modules if k == "leanfe" or k.startswith("leanfe.")]:
            del sys.modules[m]
        pkg = importlib.import_module("leanfe")
        mods = {n: importlib.import_module(f"leanfe.{n}") for n in ("common", "compress", "std_errors", "result")}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for name, mod in saved.items():
            if mod is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = mod
        # keep the reference out of the importable namespace afterwards
        for m in [k for k in sys.modules if k == "leanfe" or k.startswith("leanfe.")]:
            del sys.modules[m]
    ns = types.SimpleNamespace(pkg=pkg, **mods)
    return ns


def reference_yoco(ref, comp: dict, x_names, fe_names, vcov, ssc, cluster_names=None):
    """Run the reference's own YOCO solve + SE code (compress.py:1098-1162) on a compressed table.

    ``comp`` is the dict produced by ``oracle.ref_numpy.compress`` with key
    columns ordered x, fe, extra-cluster.  Uses the ``DuckDBResult`` branch of
    ``build_design_matrix`` (compress.py:487-499), i.e. plain NumPy inputs.
    """
    import numpy as np
    C = ref.compress
    names = list(x_names) + list(fe_names) + [c for c in (cluster_names or []) if c not in list(x_names) + list(fe_names)]
    table = {n: np.asarray(v) for n, v in zip(names, comp["keys"])}
    table.update(_n=comp["n"], _sum_y=comp["sum_y"], _sum_y_sq=comp["sum_y_sq"], _mean_y=comp["mean_y"], _wts=comp["wts"])
    res = C.DuckDBResult(table)
    M, Y, wts, all_cols, _ = C.build_design_matrix(res, list(x_names), list(fe_names))
    beta, XtX_inv = C.solve_wls(M, Y, wts)
    rss, rss_g = C.compute_rss_grouped(res, M, beta, backend="duckdb")
    p = len(all_cols)
    df_resid = comp["n_obs"] - p
    cluster_ids = e0 = None
    if vcov.lower() == "cluster":
        yhat = np.asarray(M @ beta).flatten()
        e0 = comp["sum_y"] - comp["n"] * yhat
        if len(cluster_names) == 1:
            cluster_ids = table[cluster_names[0]]
        else:
            cluster_ids = np.column_stack([table[c] for c in cluster_names])
    kx = len(x_names) + 1
    ctx = C.CompressionContext(XtX_inv=XtX_inv, rss_total=rss, rss_per_group=rss_g, n_obs=comp["n_obs"],
                               df_resid=df_resid, vcov=vcov, design_matrix=M, x_cols=all_cols[:kx],
                               cluster_ids=cluster_ids, residual_sums_per_group=e0,
                               apply_small_sample_correction=ssc)
    se, n_clusters = C.compute_se_compress(ctx)
    return dict(beta=np.asarray(beta)[1:kx], se=np.asarray(se)[1:], rss=float(rss), df_resid=df_resid,
                n_clusters=n_clusters, all_cols=all_cols)

"""CPU: host-side logic of the product (formula, strategy, result object, C-ABI surface)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import leanfe_b200 as L
from leanfe_b200 import _capi
from leanfe_b200.frontend import Frame, resolve_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_follow_reference_init():
    for name in ("leanfe", "leanfe_polars", "parse_formula", "LeanFEResult"):
        assert hasattr(L, name)


def test_parse_formula_matches_reference_goldens(golden_dir):
    J = json.load(open(os.path.join(golden_dir, "structural.json")))
    for formula, exp in J["parse_formula"]:
        got = L.parse_formula(formula)
        assert [list(map(list, v)) if isinstance(v, list) and v and isinstance(v[0], tuple) else v for v in got] == exp
        assert got.y_col == exp[0] and got.fe_cols == exp[2]
    for formula, err in J["parse_formula_errors"]:
        if err is None:
            L.parse_formula(formula)
        else:
            with pytest.raises(ValueError):
                L.parse_formula(formula)


def test_determine_strategy_matches_reference_goldens(golden_dir):
    J = json.load(open(os.path.join(golden_dir, "structural.json")))
    for kwargs, exp in J["determine_strategy"]:
        assert L.determine_strategy(**kwargs) == exp


def test_result_object_matches_reference(golden_dir):
    J = json.load(open(os.path.join(golden_dir, "structural.json")))["result"]
    R = L.LeanFEResult(coefs={"x1": 1.25, "x2": -0.5}, std_errors={"x1": 0.1, "x2": 0.25}, n_obs=12345,
                       vcov_type="cluster", iterations=4, n_clusters=321, df_resid=12000,
                       formula="y ~ x1 + x2 | a + b", fe_cols=["a", "b"], fe_dims=(300, 12))
    assert str(R) == J["str"]
    assert repr(R) == J["repr"]
    assert list(R.keys()) == J["keys"]
    for k in ("x1", "x2"):
        assert R.t_stats[k] == pytest.approx(J["t_stats"][k], rel=1e-14)
        assert R.p_values[k] == pytest.approx(J["p_values"][k], rel=1e-12, abs=1e-300)
        assert list(R.confint()[k]) == pytest.approx(J["confint"][k], rel=1e-12)
    assert R["coefs"] is R.coefs and R["coefficients"] is R.coefs and "coefficients" in R
    assert R.get("nope", 7) == 7 and repr(R["n_obs"]) == "12_345"
    multi = L.LeanFEResult(coefs={"x": 1.0}, std_errors={"x": 0.0}, n_obs=10, vcov_type="cluster", n_clusters=(3, 4))
    assert "3 × 4" in str(multi) and np.isnan(multi.t_stats["x"])


def test_reference_level_resolution():
    assert resolve_reference(["A", "B"], None, "r") == "A"
    assert resolve_reference([1, 2, 3], "2", "r") == 2
    with pytest.raises(ValueError, match="Reference category"):
        resolve_reference(["A", "B"], "Z", "r")


def test_frame_adapters(tmp_path):
    import pandas as pd
    import pyarrow as pa
    import pyarrow.parquet as pq
    d = {"y": np.arange(4.0), "g": np.array(["a", "b", "a", "c"], dtype=object), "i": np.arange(4)}
    for data in (d, pd.DataFrame(d), pa.table(d)):
        f = Frame(data, ["y", "g"])
        np.testing.assert_array_equal(np.asarray(f.column("y")), d["y"])
        assert list(f.column("g")) == list(d["g"])
    path = str(tmp_path / "t.parquet")
    pq.write_table(pa.table(d), path)
    np.testing.assert_array_equal(Frame(path, ["y", "i"]).column("i"), d["i"])
    with pytest.raises(KeyError):
        Frame(d, ["missing"])
    with pytest.raises(ValueError):
        Frame(pa.table({"y": pa.array([1.0, None])}), ["y"]).column("y")


def test_c_abi_library_exports_every_declared_symbol():
    """liblfe.so loads without a GPU and exports exactly what include/lfe.h declares."""
    header = open(os.path.join(ROOT, "include", "lfe.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+(lfe_\w+)\s*\(", header, flags=re.M)))
    assert declared, "no declarations parsed from lfe.h"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in lfe.h but not exported"
    assert sorted(_capi.EXPORTS) == declared
    lib.lfe_version.restype = ctypes.c_int
    assert lib.lfe_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.leanfe({"y": np.zeros(3), "x": np.zeros(3)}, formula="y ~ x")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "leanfe_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_argument_errors_follow_reference():
    from leanfe_b200.api import InvalidVcov, leanfe
    d = {"y": np.zeros(4), "x": np.zeros(4)}
    with pytest.raises(ValueError, match="backend must be"):
        leanfe(d, formula="y ~ x", backend="spark")
    with pytest.raises(ValueError, match="dataset must be provided"):
        leanfe(None, formula="y ~ x")
    with pytest.raises(ValueError, match="Must provide either"):
        leanfe(d)
    with pytest.raises(AssertionError):
        leanfe(d, formula="y ~ x", vcov="hc3")
    with pytest.raises(ValueError):
        leanfe(d, formula="y ~ x", vcov="hc3")
    with pytest.raises(ValueError, match="cluster_cols required"):
        leanfe(d, formula="y ~ x", vcov="cluster")
    with pytest.raises(ValueError, match="Unknown strategy"):
        leanfe(d, formula="y ~ x", strategy="magic")
    assert issubclass(InvalidVcov, AssertionError)


def test_bench_live_reference_branch_with_a_stand_in(monkeypatch):
    """bench.cpu_reference_run takes the real reference when `import polars` works (it cannot on this image): drive
    that branch with stand-ins for polars and for leanfe() so that it cannot rot unexecuted."""
    import sys
    import types

    import numpy as np

    import bench

    calls = {}

    class FakeRes:
        coefs = {f"x{j}": 1.0 + 0.1 * j for j in range(1, 6)}
        std_errors = {f"x{j}": 0.01 * j for j in range(1, 6)}
        iterations, n_obs = 4, 1000

    def fake_leanfe(df, **kw):
        calls.update(kw, rows=len(df["y"]))
        return FakeRes()

    monkeypatch.setitem(sys.modules, "polars", types.SimpleNamespace(DataFrame=lambda d: d))
    monkeypatch.setattr(bench, "live_reference", lambda: (fake_leanfe, "polars importable"))
    d, _ = bench.make_panel(1000, 1)
    r = bench.cpu_reference_run(d, steps=2, warmup=1)
    assert r["kind"] == "reference" and r["iterations"] == 4 and r["n_obs"] == 1000 and r["rate"] > 0
    np.testing.assert_allclose(r["beta"], [1.1, 1.2, 1.3, 1.4, 1.5])
    assert calls["backend"] == "polars" and calls["strategy"] == "alt_proj" and calls["cluster_cols"] == ["unit"]
    assert calls["rows"] == 1000


def test_bench_panels_are_pure_functions_of_the_global_row_index():
    """The strong-scaling records of bench.py claim 'the same panel at every N': every generator must give bit-identical
    rows whatever the shard boundaries and chunking (counter-based RNG keyed by the global row index)."""
    import torch

    import bench

    n_total = 3_000_000
    whole5 = bench.make_c5_rows(torch, "cpu", 0, 120_000, n_total, chunk=50_000)
    whole3 = bench.make_c3_rows(torch, "cpu", 0, 120_000, n_total, chunk=70_000)
    whole4, names4 = bench.make_c4_rows(torch, "cpu", 0, 120_000, chunk=33_333)
    for lo, hi in ((0, 40_000), (40_000, 120_000), (77_777, 99_999)):
        for whole, part in ((whole5, bench.make_c5_rows(torch, "cpu", lo, hi, n_total)),
                            (whole3, bench.make_c3_rows(torch, "cpu", lo, hi, n_total)),
                            (whole4, bench.make_c4_rows(torch, "cpu", lo, hi)[0])):
            for k, v in part.items():
                assert torch.equal(v, whole[k][lo:hi]), k
    # shapes of the named configs: ids in range, 23 event-study dummies, treated groups only
    assert int(whole5["unit"].max()) < n_total // bench.C5_ROWS_PER_UNIT and int(whole5["time"].max()) < bench.N_TIME
    assert int(whole3["worker"].max()) < n_total // 10 and int(whole3["firm"].max()) < n_total // 100
    assert len(names4) == 23 and "d11" not in names4
    on = whole4["d12"].bool()
    assert bool((whole4["group"][on] < 400).all()) and bool((whole4["period"][on] == 12).all())

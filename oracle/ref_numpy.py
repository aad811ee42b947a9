"""NumPy restatement of the reference's estimation hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Every function cites
the reference lines it restates (paths relative to ``/root/reference/``).  The
Polars-expression parts of the reference (``.over()`` windows, ``group_by``)
are rewritten with ``np.unique`` / ``np.bincount``; the NumPy parts
(`_run_regression`, the sandwich algebra, the YOCO solve) are restated
operation for operation.

Parity unpinned by the reference's own tests (it has no golden vectors); pinned
here by ``oracle/ref_loader.py`` (reference functions run under stub imports)
and by the dense dummy-variable OLS in :func:`dense_dummy_ols`.
"""
from __future__ import annotations

import re
from itertools import combinations

import numpy as np

# python/leanfe/polars_impl.py:24 ; python/leanfe/compress.py:40-51
MAX_FE_LEVELS = 10_000
DEFAULT_DEMEANING_ITERATIONS = 10


# ---------------------------------------------------------------------------
# (a1/A.1) group codes
# ---------------------------------------------------------------------------
def canonical_float_keys(v: np.ndarray) -> np.ndarray:
    """-0.0 -> +0.0 and every NaN -> one NaN payload (Polars group-by semantics)."""
    v = np.asarray(v, dtype=np.float64) + 0.0
    v = np.where(np.isnan(v), np.nan, v)
    return v


def codes_sorted(v: np.ndarray):
    """Sorted-rank codes: ``np.unique(v, return_inverse=True)``.

    compress.py:842, compress.py:585-590, std_errors.py:490,555.
    Returns (codes int32, uniques, counts int64).
    """
    v = np.asarray(v)
    if v.dtype.kind == "f":
        v = canonical_float_keys(v)
    uniq, inv, cnt = np.unique(v, return_inverse=True, return_counts=True)
    return inv.reshape(-1).astype(np.int32), uniq, cnt.astype(np.int64)


def codes_first_appearance(v: np.ndarray):
    """First-appearance codes (Polars Categorical physical order, polars_impl.py:130)."""
    v = np.asarray(v)
    uniq, first, inv = np.unique(v, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv.reshape(-1)].astype(np.int32), uniq[order]


# ---------------------------------------------------------------------------
# formula grammar  (common.py:51-181)
# ---------------------------------------------------------------------------
_I_RE = re.compile(r'i\((\w+)(?:\s*,\s*ref\s*=\s*["\']?([^"\')\s]+)["\']?)?\)')
_INT_RE = re.compile(r'(\w+):i\((\w+)(?:\s*,\s*ref\s*=\s*["\']?([^"\')\s]+)["\']?)?\)')


def parse_formula(formula: str):
    """common.py:51-100 (+ 103-181 for the x terms).

    Returns (y_col, x_cols, fe_cols, factor_vars, interactions, instruments).
    """
    parts = [p.strip() for p in formula.split("|")]
    if len(parts) > 3:
        raise ValueError("Formula has too many parts")
    lr = parts[0].split("~")
    if len(lr) != 2:
        raise ValueError("Formula must have exactly one '~'")
    y = lr[0].strip()
    x_cols, factor_vars, interactions = [], [], []
    for term in (t.strip() for t in lr[1].split("+")):
        if term == "":
            continue
        if ":i(" in term and term.endswith(")"):
            m = _INT_RE.match(term)
            if not m:
                raise ValueError(f"Invalid interaction syntax: {term}")
            interactions.append((m.group(1), m.group(2), m.group(3)))
        elif term.startswith("i(") and term.endswith(")"):
            m = _I_RE.match(term)
            if not m:
                raise ValueError(f"Invalid i() syntax: {term}")
            factor_vars.append((m.group(1), m.group(2)))
        else:
            x_cols.append(term)
    fe = [f.strip() for f in parts[1].split("+") if f.strip()] if len(parts) >= 2 and parts[1].strip() else []
    iv = [z.strip() for z in parts[2].split("+") if z.strip()] if len(parts) == 3 and parts[2].strip() else []
    return y, x_cols, fe, factor_vars, interactions, iv


def _sorted_categories(col: np.ndarray) -> list:
    return np.unique(np.asarray(col)).tolist()


def _resolve_ref(categories: list, ref, var: str):
    """polars_impl.py:44-55 / 91-102 (with the line-50 typo fixed)."""
    if ref is None:
        return categories[0]
    ref_cat = ref
    if categories and not isinstance(categories[0], type(ref)):
        try:
            ref_cat = type(categories[0])(ref)
        except (ValueError, TypeError):
            pass
    if ref_cat not in categories:
        raise ValueError(f"Reference category '{ref}' not found in {var}. Available: {categories}")
    return ref_cat


def expand_interactions(data: dict, interactions):
    """polars_impl.py:72-115: ``var:i(factor)`` -> Float64 columns ``f"{var}_{cat}"``."""
    out, names = {}, []
    for var, factor, ref in interactions:
        cats = _sorted_categories(data[factor])
        ref_cat = _resolve_ref(cats, ref, factor)
        for cat in cats:
            if cat == ref_cat:
                continue
            name = f"{var}_{cat}"
            out[name] = np.asarray(data[var], dtype=np.float64) * (np.asarray(data[factor]) == cat)
            names.append(name)
    return out, names


def expand_factors(data: dict, factor_vars):
    """polars_impl.py:27-69: ``i(var)`` -> UInt8 dummies ``f"{var}_{cat}"``, reference dropped.

    Labels come from the ORIGINAL values, not physical codes.
    """
    out, names = {}, []
    for var, ref in factor_vars:
        cats = _sorted_categories(data[var])
        ref_cat = _resolve_ref(cats, ref, var)
        for cat in cats:
            if cat == ref_cat:
                continue
            name = f"{var}_{cat}"
            out[name] = (np.asarray(data[var]) == cat).astype(np.float64)
            names.append(name)
    return out, names


# ---------------------------------------------------------------------------
# (a2-a6) singleton filter, ordering, alternating projections
# ---------------------------------------------------------------------------
def singleton_keep_mask(fe_codes: list[np.ndarray]) -> np.ndarray:
    """polars_impl.py:478-482 (434-435 single FE): keep iff every FE group size > 1, one pass."""
    n = fe_codes[0].shape[0]
    keep = np.ones(n, dtype=bool)
    for c in fe_codes:
        cnt = np.bincount(c)
        keep &= cnt[c] > 1
    return keep


def fe_order(cardinality: list[int]) -> list[int]:
    """polars_impl.py:485: stable ascending sort by PRE-drop cardinality."""
    return sorted(range(len(cardinality)), key=lambda j: cardinality[j])


def _group_mean(col, codes, G, w=None):
    if w is None:
        s = np.bincount(codes, weights=col, minlength=G)
        n = np.bincount(codes, minlength=G).astype(np.float64)
    else:
        s = np.bincount(codes, weights=col * w, minlength=G)
        n = np.bincount(codes, weights=w, minlength=G)
    with np.errstate(invalid="ignore", divide="ignore"):
        return s / n


def alt_proj(cols: list[np.ndarray], fe_codes: list[np.ndarray], order: list[int],
             tol: float, max_iter: int, w: np.ndarray | None = None, min_iter: int = 3):
    """polars_impl.py:490-526.  ``cols[0]`` is y.  Returns (demeaned cols, iterations, max_mean).

    Gauss-Seidel: for each FE in ``order`` every column c <- c - mean_g(c)
    (weighted: sum(w c)/sum(w), :492-500).  From iteration 3 on, after the full
    sweep, ``max_fe max_rows |mean(y).over(fe)|`` -- y only, UNWEIGHTED (:511-521)
    -- is compared with ``tol`` (:522).
    """
    cols = [np.array(c, dtype=np.float64, copy=True) for c in cols]
    Gs = [int(c.max()) + 1 if c.size else 0 for c in fe_codes]
    iterations, max_mean = 0, np.inf
    for it in range(1, max_iter + 1):
        for j in order:
            for c in cols:
                c -= _group_mean(c, fe_codes[j], Gs[j], w)[fe_codes[j]]
        if it >= min_iter:
            max_mean = 0.0
            for j in range(len(fe_codes)):
                m = _group_mean(cols[0], fe_codes[j], Gs[j])
                m = m[np.isfinite(m)]
                if m.size:
                    max_mean = max(max_mean, float(np.abs(m).max()))
            if max_mean < tol:
                iterations = it
                break
        iterations = it
    return cols, iterations, max_mean


# ---------------------------------------------------------------------------
# (a7-a8) regression  (polars_impl.py:141-285)
# ---------------------------------------------------------------------------
def solve_spd(A: np.ndarray, b: np.ndarray):
    """polars_impl.py:212-220: Cholesky solve + inverse via two solves, LinAlgError -> solve/inv."""
    try:
        L = np.linalg.cholesky(A)
        beta = np.linalg.solve(L.T, np.linalg.solve(L, b))
        Ainv = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(L.shape[0])))
        return beta, Ainv, False
    except np.linalg.LinAlgError:
        return np.linalg.solve(A, b), np.linalg.inv(A), True


def run_regression(y: np.ndarray, X: np.ndarray, w: np.ndarray | None, n_obs: int, absorbed_df: int):
    """polars_impl.py:166-232, 281-283.  X WITHOUT intercept (n x k); intercept is prepended (:168-169)."""
    k = X.shape[1]
    Xi = np.hstack([np.ones((X.shape[0], 1)), X]) if k > 0 else np.ones((n_obs, 1))
    if w is not None:
        sw = np.sqrt(w)
        Xw, Yw = Xi * sw[:, None], y * sw
        XtX, Xty = Xw.T @ Xw, Xw.T @ Yw
    else:
        XtX, Xty = Xi.T @ Xi, Xi.T @ y
    beta_full, XtX_inv, used_fallback = solve_spd(XtX, Xty)
    resid = y - Xi @ beta_full
    df_resid = n_obs - (k + 1) - absorbed_df
    rss = float(np.sum(resid ** 2))
    tss = float(np.sum((y - np.mean(y)) ** 2))
    r2 = 1 - rss / tss if tss > 0 else None
    return dict(beta_full=beta_full, beta=beta_full[1:], XtX=XtX, Xty=Xty, XtX_inv=XtX_inv,
                B=XtX_inv[1:, 1:], resid=resid, df_resid=df_resid, rss=rss, tss=tss,
                r_squared=r2, used_fallback=used_fallback)


# ---------------------------------------------------------------------------
# (a9-a12) standard errors  (std_errors.py:183-441)
# ---------------------------------------------------------------------------
def se_iid(B, resid, w, df_resid):
    """std_errors.py:183-210."""
    s2 = float(np.sum((w if w is not None else 1.0) * resid ** 2)) / df_resid
    return np.sqrt(np.maximum(s2 * np.diag(B), 0.0))


def se_hc1(X, B, resid, w, n_obs, df_resid):
    """std_errors.py:217-282: meat_ij = sum (w) x_i x_j r^2 ; V = B meat B * n/df."""
    r2 = resid ** 2 if w is None else w * resid ** 2
    meat = X.T @ (X * r2[:, None])
    V = B @ meat @ B
    return np.sqrt(np.maximum((n_obs / df_resid) * np.diag(V), 0.0)), V * (n_obs / df_resid)


def _cluster_scores(X, resid, w, cl_codes, G):
    xr = X * (resid if w is None else resid * w)[:, None]
    S = np.zeros((G, X.shape[1]))
    for j in range(X.shape[1]):
        S[:, j] = np.bincount(cl_codes, weights=xr[:, j], minlength=G)
    return S


def se_cluster_oneway(X, B, resid, w, cluster, n_obs, df_resid, ssc):
    """std_errors.py:289-347."""
    codes, uniq, _ = codes_sorted(cluster)
    G = len(uniq)
    S = _cluster_scores(X, resid, w, codes, G)
    meat = S.T @ S
    adj = (G / (G - 1)) * ((n_obs - 1) / df_resid) if ssc else G / (G - 1)
    V = adj * (B @ meat @ B)
    return np.sqrt(np.maximum(np.diag(V), 0.0)), G, V


def se_cluster_multiway(X, B, resid, w, clusters: list[np.ndarray], n_obs, df_resid, ssc):
    """std_errors.py:354-441 (CGM inclusion-exclusion, single G_min adjustment)."""
    W = len(clusters)
    V = np.zeros_like(B)
    n_list = []
    cl_codes = [codes_sorted(c)[0] for c in clusters]
    for size in range(1, W + 1):
        sign = (-1) ** (size - 1)
        for sub in combinations(range(W), size):
            if size == 1:
                codes = cl_codes[sub[0]]
                G = int(codes.max()) + 1
            else:
                stacked = np.column_stack([cl_codes[i] for i in sub])
                uq, codes = np.unique(stacked, axis=0, return_inverse=True)
                codes = codes.reshape(-1)
                G = uq.shape[0]
            if size == 1:
                n_list.append(G)
            if G <= 1:
                continue
            S = _cluster_scores(X, resid, w, codes, G)
            V += sign * (B @ (S.T @ S) @ B)
    if n_list:
        Gmin = min(n_list)
        if Gmin > 2:
            V *= Gmin / (Gmin - 1)
    if ssc:
        V *= (n_obs - 1) / df_resid
    return np.sqrt(np.maximum(np.diag(V), 0.0)), tuple(n_list), V


def standard_errors(vcov, X, B, resid, w, clusters, n_obs, df_resid, ssc):
    """Dispatcher std_errors.py:30-176, accepting both 'hc1' and 'HC1' (SURVEY A.6.1)."""
    v = vcov.lower()
    if v == "iid":
        return se_iid(B, resid, w, df_resid), None
    if v == "hc1":
        return se_hc1(X, B, resid, w, n_obs, df_resid)[0], None
    if v == "cluster":
        if not clusters:
            raise ValueError("cluster_cols required for vcov='cluster'")
        if len(clusters) == 1:
            se, G, _ = se_cluster_oneway(X, B, resid, w, clusters[0], n_obs, df_resid, ssc)
            return se, G
        se, Gs, _ = se_cluster_multiway(X, B, resid, w, clusters, n_obs, df_resid, ssc)
        return se, Gs
    raise ValueError(f"Unknown vcov type: {vcov}")


# ---------------------------------------------------------------------------
# IV / 2SLS  (common.py:188-287, polars_impl.py:176-198 / 229-278, std_errors.py:448-598)
# ---------------------------------------------------------------------------
def iv_2sls(Y, X, Z, w=None):
    """common.py:188-287.  X (n x p) are ALL treated as endogenous and regressed on Z (n x m), m >= p."""
    if Z.shape[1] < X.shape[1]:
        raise ValueError(f"Under-identified: {Z.shape[1]} instruments for {X.shape[1]} endogenous variables")
    if w is not None:                                              # :251-274
        sw = np.sqrt(w)
        Zw, Xw, Yw = Z * sw[:, None], X * sw[:, None], Y * sw
        gamma = np.linalg.solve(Zw.T @ Zw, Zw.T @ Xw)
        Xh = Z @ gamma
        Xhw = Xh * sw[:, None]
        beta = np.linalg.solve(Xhw.T @ Xhw, Xhw.T @ Yw)
    else:                                                          # :277-287
        gamma = np.linalg.solve(Z.T @ Z, Z.T @ X)
        Xh = Z @ gamma
        beta = np.linalg.solve(Xh.T @ Xh, Xh.T @ Y)
    return beta, Xh


def run_regression_iv(y, X, Z, w, n_obs, absorbed_df):
    """polars_impl.py:166-198, 229-232 with instruments.  X, Z WITHOUT intercept.  The intercept column is one of
    the 'endogenous' columns; a constant joins Z only if Z has fewer columns than [1 | X] and no constant
    column (:179-182).  Residuals use the FITTED Xh (:229), (X'X)^-1 is (Xh'WXh)^-1 (:187-198)."""
    k = X.shape[1]
    Xi = np.hstack([np.ones((X.shape[0], 1)), X]) if k > 0 else np.ones((n_obs, 1))
    Zi = Z
    if Xi.shape[1] > Zi.shape[1]:
        if not any(np.allclose(col, 1.0) for col in Zi.T):
            Zi = np.column_stack([np.ones(Zi.shape[0]), Zi])
    beta_full, Xh = iv_2sls(y, Xi, Zi, w)
    A = (Xh * np.sqrt(w)[:, None]).T @ (Xh * np.sqrt(w)[:, None]) if w is not None else Xh.T @ Xh
    try:
        L = np.linalg.cholesky(A)
        XtX_inv = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(L.shape[0])))
    except np.linalg.LinAlgError:
        XtX_inv = np.linalg.inv(A)
    resid = y - Xh @ beta_full
    df_resid = n_obs - (k + 1) - absorbed_df
    return dict(beta_full=beta_full, beta=beta_full[1:], XtX_inv=XtX_inv, Xh=Xh, resid=resid, df_resid=df_resid,
                rss=float(np.sum(resid ** 2)), n_z=Zi.shape[1])


def standard_errors_iv(vcov, Xh, XtX_inv, resid, w, clusters, n_obs, df_resid, ssc):
    """std_errors.py:85-145 dispatch with is_iv=True -> :183-210 (iid, full (X'X)^-1) and :448-598 (hc1,
    one-way, multi-way on Xh).  The IV variants use the same sandwich algebra and scalar corrections as the
    OLS ones, applied to Xh and the full inverse; the caller strips the intercept entry (polars_impl.py:271-274)."""
    v = vcov.lower()
    if v == "iid":
        return se_iid(XtX_inv, resid, w, df_resid), None
    if v == "hc1":
        return se_hc1(Xh, XtX_inv, resid, w, n_obs, df_resid)[0], None
    if not clusters:
        raise ValueError("cluster_cols required for vcov='cluster'")
    if len(clusters) == 1:
        se, G, _ = se_cluster_oneway(Xh, XtX_inv, resid, w, clusters[0], n_obs, df_resid, ssc)
        return se, G
    se, Gs, _ = se_cluster_multiway(Xh, XtX_inv, resid, w, clusters, n_obs, df_resid, ssc)
    return se, Gs


# ---------------------------------------------------------------------------
# (a13-a18) YOCO  (compress.py)
# ---------------------------------------------------------------------------
def determine_strategy(vcov, has_instruments, fe_cardinality=None, max_fe_levels=MAX_FE_LEVELS,
                       n_obs=None, n_x_cols=None, estimated_compression_ratio=None):
    """compress.py:96-184."""
    if has_instruments:
        return "alt_proj"
    if vcov.lower() not in ("iid", "hc1", "cluster"):
        return "alt_proj"
    if fe_cardinality is None:
        return "compress"
    total = sum(fe_cardinality.values())
    mx = max(fe_cardinality.values()) if fe_cardinality else 0
    if mx > max_fe_levels:
        return "alt_proj"
    if total > max_fe_levels * 2:
        return "alt_proj"
    if estimated_compression_ratio is not None and n_obs is not None:
        n_comp = int(n_obs * estimated_compression_ratio)
        yoco = 1.0 * n_obs + 1.0 * n_comp * total + total ** 2
        fwl = DEFAULT_DEMEANING_ITERATIONS * len(fe_cardinality) * n_obs
        return "compress" if yoco < fwl else "alt_proj"
    return "compress"


def _key_matrix(key_cols: list[np.ndarray]):
    """Per-column sorted-rank codes; lexicographic order of the code tuples == order of the raw tuples."""
    codes, uniqs = [], []
    for c in key_cols:
        cd, uq, _ = codes_sorted(c)
        codes.append(cd)
        uniqs.append(uq)
    return codes, uniqs


def estimate_compression_ratio(key_cols: list[np.ndarray]) -> float:
    """compress.py:232-253: #unique rows over set(x_cols + fe_cols) / N."""
    if not key_cols:
        return 1.0
    n = key_cols[0].shape[0]
    codes, _ = _key_matrix(key_cols)
    uq = np.unique(np.column_stack(codes), axis=0)
    return uq.shape[0] / max(n, 1)


def compress(y, key_cols: list[np.ndarray], w=None):
    """compress.py:325-356.  Group by the key columns (x_cols + fe_cols + extra cluster cols).

    The reference leaves the output row order unspecified; the oracle (and the
    device path) emit groups in ascending lexicographic key order.
    Returns dict(keys=[per-column arrays], n, sum_y, sum_y_sq, mean_y, wts, n_obs).
    """
    n_obs = y.shape[0]
    codes, uniqs = _key_matrix(key_cols)
    if codes:
        uq, inv = np.unique(np.column_stack(codes), axis=0, return_inverse=True)
        inv = inv.reshape(-1)
        G = uq.shape[0]
    else:
        uq, inv, G = np.zeros((1, 0), dtype=np.int64), np.zeros(n_obs, dtype=np.int64), 1
    if w is None:
        n = np.bincount(inv, minlength=G).astype(np.float64)
        sy = np.bincount(inv, weights=y, minlength=G)
        syy = np.bincount(inv, weights=y * y, minlength=G)
    else:
        n = np.bincount(inv, weights=w, minlength=G)
        sy = np.bincount(inv, weights=y * w, minlength=G)
        syy = np.bincount(inv, weights=y * y * w, minlength=G)
    keys = [uniqs[j][uq[:, j]] for j in range(len(key_cols))]
    key_codes = [uq[:, j].astype(np.int32) for j in range(len(key_cols))]
    return dict(keys=keys, key_codes=key_codes, n=n, sum_y=sy, sum_y_sq=syy, mean_y=sy / n,
                wts=np.sqrt(n), n_obs=n_obs, n_compressed=G)


def build_design_dense(X: np.ndarray, fe_vals: list[np.ndarray]):
    """compress.py:503-564, 622-652: [1 | X | D], first SORTED level of each FE dropped."""
    G = X.shape[0]
    blocks = [np.ones((G, 1)), X]
    n_levels = 0
    for f in fe_vals:
        cats = np.unique(f)
        n_levels += len(cats)
        for cat in cats[1:]:
            blocks.append((f == cat).astype(np.float64)[:, None])
    return np.hstack(blocks), n_levels


def yoco_fit(comp: dict, k: int, n_fe: int, vcov="iid", ssc=False, cluster_key_idx: list[int] | None = None):
    """compress.py:1098-1174: design, WLS (659-747), grouped RSS (754-810), SEs (854-1042).

    ``comp['keys']`` = x columns (k) then FE columns (n_fe) then extra cluster columns.
    ``cluster_key_idx`` indexes into ``comp['keys']``.
    """
    X = np.column_stack([np.asarray(c, dtype=np.float64) for c in comp["keys"][:k]]) if k else np.zeros((comp["n"].shape[0], 0))
    fe_vals = comp["keys"][k:k + n_fe]
    M, _ = build_design_dense(X, fe_vals)
    p = M.shape[1]
    wts, Y = comp["wts"], comp["mean_y"]
    Mw, Yw = M * wts[:, None], Y * wts
    XtX, Xty = Mw.T @ Mw, Mw.T @ Yw
    try:
        L = np.linalg.cholesky(XtX)
        beta = np.linalg.solve(L.T, np.linalg.solve(L, Xty))
        XtX_inv = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(p)))
    except np.linalg.LinAlgError:
        beta = np.linalg.lstsq(XtX, Xty, rcond=None)[0]
        XtX_inv = np.linalg.pinv(XtX)
    yhat = M @ beta
    rss_g = comp["sum_y_sq"] - 2 * yhat * comp["sum_y"] + comp["n"] * yhat ** 2
    rss = float(rss_g.sum())
    n_obs = comp["n_obs"]
    df_resid = n_obs - p
    kx = k + 1
    v = vcov.lower()
    n_clusters = None
    if v == "iid":
        se_full = np.sqrt(np.maximum(np.diag(XtX_inv) * (rss / df_resid), 0.0))
    elif v == "hc1":
        meat = M.T @ (M * rss_g[:, None])
        se_full = np.sqrt(np.maximum(np.diag(XtX_inv @ meat @ XtX_inv) * (n_obs / df_resid), 0.0))
    elif v == "cluster":
        e0 = comp["sum_y"] - comp["n"] * yhat
        Sg = M * e0[:, None]
        cl = [comp["keys"][i] for i in cluster_key_idx]
        if len(cl) == 1:
            codes, uq, _ = codes_sorted(cl[0])
            G = len(uq)
            S = np.zeros((G, p))
            np.add.at(S, codes, Sg)
            adj = (G / (G - 1)) * ((n_obs - 1) / df_resid) if ssc else G / (G - 1)
            se_full = np.sqrt(np.maximum(np.diag(adj * (XtX_inv @ (S.T @ S) @ XtX_inv)), 0.0))
            n_clusters = G
        else:
            V = np.zeros_like(XtX_inv)
            firsts = []
            cl_codes = [codes_sorted(c)[0] for c in cl]
            for size in range(1, len(cl) + 1):
                sign = (-1) ** (size - 1)
                for sub in combinations(range(len(cl)), size):
                    if size == 1:
                        codes = cl_codes[sub[0]]
                        G = int(codes.max()) + 1
                        firsts.append(G)
                    else:
                        uq, codes = np.unique(np.column_stack([cl_codes[i] for i in sub]), axis=0, return_inverse=True)
                        codes = codes.reshape(-1)
                        G = uq.shape[0]
                    if G <= 1:
                        continue
                    S = np.zeros((G, p))
                    np.add.at(S, codes, Sg)
                    V += sign * (XtX_inv @ (S.T @ S) @ XtX_inv)
            if firsts:
                Gmin = min(firsts)
                if Gmin > 2:
                    V *= Gmin / (Gmin - 1)
            if ssc:
                V *= (n_obs - 1) / df_resid
            se_full = np.sqrt(np.maximum(np.diag(V), 0.0))
            n_clusters = tuple(firsts)
    else:
        raise ValueError(f"vcov must be 'iid', 'HC1', or 'cluster', got '{vcov}'")
    return dict(beta=beta[1:kx], se=se_full[:kx][1:], beta_full=beta, XtX_inv=XtX_inv, rss=rss,
                rss_g=rss_g, df_resid=df_resid, p=p, n_clusters=n_clusters,
                fe_dims=tuple(len(np.unique(f)) for f in fe_vals) if n_fe else None)


# ---------------------------------------------------------------------------
# independent cross-check: dense dummy-variable OLS
# ---------------------------------------------------------------------------
def dense_dummy_ols(y, X, fe_codes: list[np.ndarray], w=None):
    """lstsq on [1 | X | D_fe(-first)] -- what alt_proj must approach as tol -> 0."""
    n = y.shape[0]
    blocks = [np.ones((n, 1)), X]
    for c in fe_codes:
        for g in np.unique(c)[1:]:
            blocks.append((c == g).astype(np.float64)[:, None])
    M = np.hstack(blocks)
    if w is not None:
        sw = np.sqrt(w)
        beta = np.linalg.lstsq(M * sw[:, None], y * sw, rcond=None)[0]
    else:
        beta = np.linalg.lstsq(M, y, rcond=None)[0]
    return beta[1:1 + X.shape[1]], np.linalg.matrix_rank(M), M


# ---------------------------------------------------------------------------
# end-to-end: leanfe_polars()  (polars_impl.py:287-579)
# ---------------------------------------------------------------------------
def leanfe_oracle(data: dict, y_col=None, x_cols=None, fe_cols=None, formula=None, strategy="auto",
                  weights=None, demean_tol=1e-6, max_iter=50, vcov="iid", cluster_cols=None, ssc=True,
                  return_internals=False):
    """Restatement of ``leanfe()`` -> ``leanfe_polars()`` on a dict of NumPy columns.

    Follows leanfe.py:14-30 defaults and polars_impl.py:311-579, with the
    documented quirk fixes of SURVEY A.6 (hc1 spelling, expanded columns kept,
    factor labels from original values).  IV and sample_frac are out of scope.
    """
    if formula is not None:
        y_col, x_cols, fe_cols, factor_vars, interactions, instruments = parse_formula(formula)
    elif y_col is None or x_cols is None or fe_cols is None:
        raise ValueError("Must provide either 'formula' or (y_col, x_cols, fe_cols)")
    else:
        factor_vars, interactions, instruments = [], [], []
    instruments = list(instruments)
    if instruments and strategy == "compress":
        raise ValueError("strategy='compress' does not support instruments")
    if vcov.lower() not in ("iid", "hc1", "cluster"):
        raise ValueError(f"Invalid vcov method: {vcov}")
    x_cols, fe_cols = list(x_cols), list(fe_cols or [])
    data = dict(data)
    if interactions:
        new, names = expand_interactions(data, interactions)
        data.update(new)
        x_cols += names
    if factor_vars:
        new, names = expand_factors(data, factor_vars)
        data.update(new)
        x_cols += names

    n0 = np.asarray(data[y_col]).shape[0]
    fe_codes0, fe_card = [], []
    for fe in fe_cols:
        c, uq, _ = codes_sorted(data[fe])
        fe_codes0.append(c)
        fe_card.append(len(uq))
    est_ratio = None
    if strategy == "auto":
        est_ratio = estimate_compression_ratio([np.asarray(data[c]) for c in dict.fromkeys(x_cols + fe_cols)])
        if not fe_cols:
            strategy = "ols" if est_ratio >= 0.8 or instruments else "compress"
        elif len(fe_cols) == 1:
            strategy = "demean"
        else:
            strategy = determine_strategy(vcov, bool(instruments), dict(zip(fe_cols, fe_card)), MAX_FE_LEVELS,
                                          n0, len(x_cols), est_ratio)
    w_full = np.asarray(data[weights], dtype=np.float64) if weights is not None else None
    y_full = np.asarray(data[y_col], dtype=np.float64)
    X_full = [np.asarray(data[c], dtype=np.float64) for c in x_cols + instruments]   # instruments demeaned too (:486)
    n_z = len(instruments)

    if strategy == "compress":
        cl = list(cluster_cols) if cluster_cols else []
        group_cols = x_cols + fe_cols
        extra = [c for c in cl if c not in group_cols]
        group_cols = group_cols + extra
        comp = compress(y_full, [np.asarray(data[c]) for c in group_cols], w_full)
        idx = [group_cols.index(c) for c in cl] if vcov.lower() == "cluster" and cl else None
        fit = yoco_fit(comp, len(x_cols), len(fe_cols), vcov, ssc, idx)
        out = dict(coefs=dict(zip(x_cols, fit["beta"])), std_errors=dict(zip(x_cols, fit["se"])),
                   n_obs=comp["n_obs"], n_compressed=comp["n_compressed"], iterations=0, vcov_type=vcov,
                   df_resid=fit["df_resid"], rss=fit["rss"], n_clusters=fit["n_clusters"],
                   fe_dims=fit["fe_dims"], strategy="compress", compression_ratio=None, r_squared=None)
        if return_internals:
            out["_comp"], out["_fit"] = comp, fit
        return out

    if strategy == "demean":
        if len(fe_cols) != 1:
            raise ValueError("Strategy 'demean' requires exactly one FE column.")
    elif strategy == "alt_proj":
        if not fe_cols:
            raise ValueError("Strategy 'alt_proj' requires FE-cols.")
    elif strategy != "ols":
        raise ValueError(f"Unknown strategy: {strategy}")

    if strategy == "ols":
        keep = np.ones(n0, dtype=bool)
        cols = [y_full] + X_full
        iterations, absorbed, fe_dims, max_mean = 0, 0, None, None
        w = w_full
    else:
        keep = singleton_keep_mask(fe_codes0)
        fe_codes = [c[keep] for c in fe_codes0]
        w = w_full[keep] if w_full is not None else None
        cols0 = [c[keep] for c in [y_full] + X_full]
        if strategy == "demean":
            G = fe_card[0]
            cols = [c - _group_mean(c, fe_codes[0], G, w)[fe_codes[0]] for c in cols0]
            iterations, max_mean = 1, None
        else:
            cols, iterations, max_mean = alt_proj(cols0, fe_codes, fe_order(fe_card), demean_tol, max_iter, w)
        fe_dims = tuple(int(np.unique(c).size) for c in fe_codes)
        absorbed = (sum(fe_dims) - len(fe_cols)) if strategy == "alt_proj" else fe_dims[0] - 1
    n_obs = int(keep.sum())
    y = cols[0]
    X = np.column_stack(cols[1:]) if len(cols) > 1 else np.zeros((n_obs, 0))
    clusters = [np.asarray(data[c])[keep] for c in cluster_cols] if cluster_cols else None
    if n_z:
        X, Z = X[:, : X.shape[1] - n_z], X[:, X.shape[1] - n_z:]
        reg = run_regression_iv(y, X, Z, w, n_obs, absorbed)
        se_full, n_clusters = standard_errors_iv(vcov, reg["Xh"], reg["XtX_inv"], reg["resid"], w, clusters, n_obs,
                                                 reg["df_resid"], ssc)
        out = dict(coefs=dict(zip(x_cols, reg["beta"])), std_errors=dict(zip(x_cols, se_full[1:])), n_obs=n_obs,
                   iterations=iterations, vcov_type=vcov, n_clusters=n_clusters, df_resid=reg["df_resid"],
                   fe_dims=fe_dims, r_squared=None, compression_ratio=est_ratio, strategy=strategy,
                   n_compressed=None, rss=reg["rss"], is_iv=True, n_instruments=n_z)
        if return_internals:
            out.update(_reg=reg, _cols=cols, _keep=keep, _max_mean=max_mean, _w=w)
        return out
    reg = run_regression(y, X, w, n_obs, absorbed)
    se, n_clusters = standard_errors(vcov, X, reg["B"], reg["resid"], w, clusters, n_obs, reg["df_resid"], ssc)
    out = dict(coefs=dict(zip(x_cols, reg["beta"])), std_errors=dict(zip(x_cols, se)), n_obs=n_obs,
               iterations=iterations, vcov_type=vcov, n_clusters=n_clusters, df_resid=reg["df_resid"],
               fe_dims=fe_dims, r_squared=reg["r_squared"], compression_ratio=est_ratio,
               strategy=strategy, n_compressed=None, rss=reg["rss"])
    if return_internals:
        out.update(_reg=reg, _cols=cols, _keep=keep, _max_mean=max_mean, _w=w)
    return out

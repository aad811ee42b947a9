"""``leanfe()`` -- the reference's public entry point, backed by the B200 library.

Signature, defaults, strategy selection and error behaviour follow the reference
(python/leanfe/leanfe.py:14-184 -> python/leanfe/polars_impl.py:287-579).  Everything per-row runs in
liblfe.so (``include/lfe.h``); this module only orchestrates.  Deviations (all listed in DESIGN.md):
'hc1' and 'HC1' are both accepted, expanded factor/interaction columns are kept (reference bug), dummy
names use the original labels, IV/2SLS is implemented (single-GPU, strategies ols / demean / alt_proj;
rejected for 'compress' and multi-GPU), ``compression_ratio`` is only computed when the strategy rule
actually needs it.
"""
from __future__ import annotations

import contextlib
import os
import time
import warnings

import numpy as np
import torch

from . import _capi as capi
from .engine import Engine, get_engine
from .formula import parse_formula
from .frontend import Frame, KeyColumn, Uploader, resolve_reference, to_device_f64, to_device_key
from .result import LeanFEResult
from .strategy import MAX_FE_LEVELS, determine_strategy


class InvalidVcov(AssertionError, ValueError):
    """The reference asserts on the vcov name (polars_impl.py:158) and raises ValueError later
    (std_errors.py:176); this is both."""


_VCOV = {"iid": capi.VCOV_IID, "hc1": capi.VCOV_HC1, "cluster": capi.VCOV_CLUSTER}


def leanfe(data=None, y_col=None, x_cols=None, fe_cols=[], formula=None, strategy="auto", weights=None,
           demean_tol=1e-6, max_iter=50, vcov="iid", cluster_cols=None, ssc=True, sample_frac=None,
           backend="polars", con=None, device=None, profile=False, local_fe=None, deterministic=None) -> LeanFEResult:
    """Fast fixed-effects regression; drop-in for ``leanfe.leanfe`` with the compute on a B200.

    ``backend`` accepts "polars" (the backend this package replaces), "b200" or "cuda"; "duckdb" is a
    different engine of the reference and is out of scope.
    """
    if backend in ("polars", "b200", "cuda"):
        if data is None:
            raise ValueError(f"A dataset must be provided when using backend = {backend}.")
        return leanfe_b200(data=data, y_col=y_col, x_cols=x_cols, fe_cols=fe_cols, formula=formula,
                           strategy=strategy, weights=weights, demean_tol=demean_tol, max_iter=max_iter, vcov=vcov,
                           cluster_cols=cluster_cols, ssc=ssc, sample_frac=sample_frac, device=device,
                           profile=profile, local_fe=local_fe, deterministic=deterministic)
    if backend == "duckdb":
        raise ValueError("backend='duckdb' is not provided by leanfe_b200 (out of scope: it replaces the Polars "
                         "backend's compute); use backend='polars'")
    raise ValueError(f"backend must be 'polars' or 'duckdb', got '{backend}'")


def leanfe_b200(data, demean_tol=1e-6, y_col=None, x_cols=None, fe_cols=None, formula=None, strategy="auto",
                weights=None, max_iter=50, vcov="iid", cluster_cols=None, ssc=True, sample_frac=None, device=None,
                profile=False, local_fe=None, deterministic=None) -> LeanFEResult:
    """Counterpart of ``leanfe_polars`` (polars_impl.py:287-301) on the device.

    Multi-GPU (one process per GPU, ``Engine.comm_init_from_torch()`` called): ``data`` holds THIS rank's
    rows.  With ``local_fe=None`` (default) the rows may be sharded ARBITRARILY: the library factorises every
    key column in one shared code space, cuts the code range of the highest-cardinality FE into ``world``
    row-balanced ranges and moves the rows with one NCCL all-to-all per column (``lfe_shard_plan`` /
    ``lfe_shard_exchange``), so the singleton filter and FE order are those of the GLOBAL frame
    (polars_impl.py:478-485).  ``local_fe=<name>`` declares that the caller already partitioned the rows so
    that the groups of that FE never span ranks, and skips the exchange.  All other FE tables, the Gram matrix
    and the meat are all-reduced over NCCL inside liblfe; every rank returns the same result.
    """
    t_start = time.perf_counter()
    if formula is not None:
        y_col, x_cols, fe_cols, factor_vars, interactions, instruments = parse_formula(formula)
    elif y_col is None or x_cols is None or fe_cols is None:
        raise ValueError("Must provide either 'formula' or (y_col, x_cols, fe_cols)")
    else:
        factor_vars, interactions, instruments = [], [], []
    instruments = list(instruments)
    if instruments and strategy == "compress":
        # the reference's YOCO path takes no instruments and would silently fit OLS (compress.py:1049-1060)
        raise ValueError("strategy='compress' does not support instruments; use 'alt_proj', 'demean' or 'ols'")
    if not isinstance(vcov, str) or vcov.lower() not in _VCOV:
        raise InvalidVcov(f"Invalid vcov method: {vcov}")
    vkind = _VCOV[vcov.lower()]
    cluster_cols = list(cluster_cols) if cluster_cols else None
    if vkind == capi.VCOV_CLUSTER and not cluster_cols:
        raise ValueError("cluster_cols required for vcov='cluster'")
    x_cols = list(x_cols)
    fe_cols = list(fe_cols) if fe_cols is not None else []
    if strategy not in ("auto", "compress", "alt_proj", "demean", "ols"):
        raise ValueError(f"Unknown strategy: {strategy}")

    # needed columns (polars_impl.py:325-339)
    needed = [y_col] + x_cols + fe_cols + [z for z in instruments if z not in x_cols + fe_cols + [y_col]]
    for var, _ in factor_vars:
        if var not in needed:
            needed.append(var)
    for var, factor, _ in interactions:
        for c in (var, factor):
            if c not in needed:
                needed.append(c)
    if cluster_cols:
        needed += [c for c in cluster_cols if c not in needed]
    if weights is not None and weights not in needed:
        needed.append(weights)

    eng = get_engine(device)
    if deterministic is None:
        deterministic = os.environ.get("LFE_DETERMINISTIC", "0") not in ("", "0")
    if deterministic:
        if eng.world > 1:
            raise NotImplementedError("deterministic=True is single-GPU only")
    eng.set_option(capi.OPT_DETERMINISTIC, 1 if deterministic else 0)
    if profile:
        eng.set_profiling(True)
        eng.reset_timers()
    frame = Frame(data, needed)
    dist = eng.world > 1
    if dist:
        if strategy not in ("alt_proj", "demean", "compress"):
            raise NotImplementedError("multi-GPU fits support strategy='alt_proj' / 'demean' / 'compress' "
                                      "(pass it explicitly)")
        if strategy == "compress":
            local_fe = None          # YOCO shards rows arbitrarily: every key column uses the shared code space
        elif local_fe is not None and local_fe not in fe_cols:
            raise ValueError(f"multi-GPU fit: local_fe={local_fe!r} is not one of the fixed effects {fe_cols}")
        elif not fe_cols:
            raise ValueError("multi-GPU alt_proj / demean fits need at least one fixed effect")
        if sample_frac is not None:
            raise NotImplementedError("sample_frac is not supported in multi-GPU fits")
        if instruments:
            raise NotImplementedError("IV/2SLS is single-GPU only")
        if cluster_cols and len(cluster_cols) > 1 and strategy != "compress":
            raise NotImplementedError("multi-way clustering is single-GPU only (except strategy='compress')")

    # every H2D copy is enqueued now, ids first: the id-only work below overlaps the fp64 uploads
    key_first = fe_cols + (cluster_cols or []) + [f for _, f, _ in interactions] + [v for v, _ in factor_vars]
    up = Uploader(eng, frame, key_first + ([weights] if weights is not None else []) + needed)
    try:
        numeric: dict[str, torch.Tensor] = {}
        keys: dict[str, KeyColumn] = {}

        owned_ids: set[int] = set()      # device tensors created by this call (uploads, casts): safe to overwrite

        def num(name):
            if name not in numeric:
                src = up.get(name)
                numeric[name] = to_device_f64(eng, src)
                if name in up.events_seen or numeric[name] is not src:
                    owned_ids.add(id(numeric[name]))
            return numeric[name]

        def key(name, labels=False):
            if name not in keys or (labels and keys[name].labels is None):
                # multi-GPU: every key column except the shard key needs ONE code space shared by all ranks
                keys[name] = to_device_key(eng, up.get(name), want_labels=labels, global_codes=dist and name != local_fe)
            return keys[name]

        fe_keys = [key(c) for c in fe_cols]
        cl_keys = [key(c) for c in cluster_cols] if cluster_cols else []
        base_x = list(x_cols)
        extra_z = [z for z in instruments if z not in x_cols]
        fe_card_order = [k_.n_groups for k_ in fe_keys]      # cardinalities that fix the sweep order (:485)
        if dist and strategy != "compress":
            if local_fe is None:
                # library-side repartition: all key columns are coded in the shared code space; the rows
                # move so that every group of the highest-cardinality FE lives on one rank
                j_sh = max(range(len(fe_cols)), key=lambda j: fe_keys[j].n_groups)
                local_fe = fe_cols[j_sh]
                num_names = list(dict.fromkeys([y_col] + base_x + extra_z + ([weights] if weights is not None else [])
                                               + [v for v, _, _ in interactions]))
                key_names = list(dict.fromkeys(fe_cols + (cluster_cols or []) + [f for _, f, _ in interactions]
                                               + [v for v, _ in factor_vars]))
                for nm in [f for _, f, _ in interactions] + [v for v, _ in factor_vars]:
                    key(nm, labels=True)
                tensors = [num(nm) for nm in num_names] + [key(nm).codes for nm in key_names]
                sp = eng.shard_plan(fe_keys[j_sh].codes, fe_keys[j_sh].n_groups)
                lo, hi = sp["cuts"][eng.rank], sp["cuts"][eng.rank + 1]
                moved = eng.shard_exchange(sp, tensors, [0] * len(num_names) + [lo if nm == local_fe else 0 for nm in key_names])
                del tensors
                for nm, t in zip(num_names, moved):
                    numeric[nm] = t
                    owned_ids.add(id(t))
                for nm, t in zip(key_names, moved[len(num_names):]):
                    keys[nm] = KeyColumn(t, (hi - lo) if nm == local_fe else keys[nm].n_groups, keys[nm].labels, None)
                fe_keys = [keys[c] for c in fe_cols]
                cl_keys = [keys[c] for c in cluster_cols] if cluster_cols else []
            else:
                # caller-partitioned rows: the shard key was coded rank-locally; its global cardinality (for the sweep
                # order, which must be the same on every rank) is the sum over ranks
                j_sh = fe_cols.index(local_fe)
                fe_card_order[j_sh] = int(eng.allreduce(torch.tensor([float(fe_keys[j_sh].n_groups)], dtype=torch.float64,
                                                                     device=eng.device)).item())

        def xsrc(name, keep_int):
            # YOCO only groups by the regressors: integer-typed columns (dummies, counts) are factorised as they are,
            # without a 8-byte-per-row fp64 copy
            t = up.get(name)
            if keep_int and isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in (torch.uint8, torch.int32, torch.int64):
                return t
            return num(name)

        iv_state = {"k": None, "z_idx": []}

        def columns(keep_int=False):
            """(y, X, names, w): the numeric columns, waited for / expanded only when first needed."""
            y_ = num(y_col)
            X_ = [xsrc(c, keep_int) for c in base_x]
            names = list(base_x)
            w_ = num(weights) if weights is not None else None
            # interactions then factors (polars_impl.py:350-365); labels come from the original values
            for var, factor, ref in interactions:
                kc = key(factor, labels=True)
                ref_cat = resolve_reference(kc.labels, ref, factor)
                lv = [j for j, cat in enumerate(kc.labels) if cat != ref_cat]
                X_ += eng.expand_dummies(kc.codes, lv, base=num(var))          # one pass for all levels
                names += [f"{var}_{kc.labels[j]}" for j in lv]
            for var, ref in factor_vars:
                kc = key(var, labels=True)
                ref_cat = resolve_reference(kc.labels, ref, var)
                lv = [j for j, cat in enumerate(kc.labels) if cat != ref_cat]
                X_ += eng.expand_dummies(kc.codes, lv, as_uint8=keep_int)
                names += [f"{var}_{kc.labels[j]}" for j in lv]
            # instruments ride along as extra columns: they are demeaned with y and X (polars_impl.py:431, 486); an
            # exogenous regressor listed among the instruments is the same column, demeaned once
            iv_state["k"] = len(names)
            for z in instruments:
                if z not in names:
                    X_.append(num(z))
                    names.append(z)
            iv_state["z_idx"] = [names.index(z) for z in instruments]
            return y_, X_, names, w_

        cols_cache = []

        def get_columns():
            if not cols_cache:
                cols_cache.append(columns())
                owned_ids.update(id(t) for t in cols_cache[0][1][len(base_x):])      # expanded dummies are ours
            return cols_cache[0]

        get_columns.owned = lambda t: id(t) in owned_ids
        if not interactions and not factor_vars and weights is None:
            # plain columns: they can be handed out one by one, as each upload completes
            get_columns.lazy = dict(y=y_col, x=base_x + extra_z, get=num,
                                    pending=lambda: any(nm in up.events for nm in [y_col] + base_x + extra_z))
            iv_state.update(k=len(base_x), z_idx=[(base_x + extra_z).index(z) for z in instruments])

        n0 = fe_keys[0].codes.numel() if fe_keys else get_columns()[0].numel()

        if sample_frac is not None:
            # the reference samples with Polars' RNG (seed 42, polars_impl.py:359-360); not reproducible without
            # Polars, so sampled fits are parity-unpinned
            g = torch.Generator(device=eng.device)
            g.manual_seed(42)
            m = int(n0 * sample_frac)
            idx = torch.randperm(n0, generator=g, device=eng.device)[:m].sort().values
            y, X, names, w = get_columns()
            cols_cache[0] = (y[idx].contiguous(), [c[idx].contiguous() for c in X], names,
                             w[idx].contiguous() if w is not None else None)
            # cardinalities, group sizes and the strategy rule see the SAMPLED frame, as in the reference
            # (polars_impl.py:359-373: sample first, n_unique afterwards): re-rank the sampled codes so that levels
            # without a sampled row do not count towards fe_dims / df_resid
            def resample(k_):
                c_, G_, _ = eng.factorize(k_.codes[idx].contiguous())
                return KeyColumn(c_, G_, None, eng.factorize_counts(c_, G_))

            by_id = {id(k_): resample(k_) for k_ in {id(k_): k_ for k_ in fe_keys + cl_keys}.values()}
            fe_keys = [by_id[id(k_)] for k_ in fe_keys]
            cl_keys = [by_id[id(k_)] for k_ in cl_keys]
            fe_card_order = [k_.n_groups for k_ in fe_keys]
            n0 = m

        fe_card = {c: k_.n_groups for c, k_ in zip(fe_cols, fe_keys)}
        est_ratio = None
        if strategy == "auto":
            def ratio():
                # distinct rows over set(x_cols + fe_cols) / N   (compress.py:232-253)
                _, X, names, _ = get_columns()
                names = names[: iv_state["k"]]
                kcs = []
                for nm in dict.fromkeys(names + fe_cols):
                    if nm in fe_cols:
                        kcs.append(fe_keys[fe_cols.index(nm)])
                    else:
                        c_, G_, _ = eng.factorize(X[names.index(nm)])
                        kcs.append(KeyColumn(c_, G_))
                if not kcs:
                    return 1.0
                r = eng.compress([k_.codes for k_ in kcs], [k_.n_groups for k_ in kcs], None, count_only=True)
                return r["n_compressed"] / max(n0, 1)

            if not fe_cols:
                est_ratio = ratio()
                # (with instruments the reference would pick 'compress' here and silently drop them; we fit 2SLS)
                strategy = "ols" if est_ratio >= 0.8 or instruments else "compress"
            elif len(fe_cols) == 1:
                strategy = "demean"
            else:
                # rules 1/2 of determine_strategy do not need the ratio; compute it only for the cost model
                total, mx = sum(fe_card.values()), max(fe_card.values())
                if mx > MAX_FE_LEVELS or total > 2 * MAX_FE_LEVELS:
                    strategy = "alt_proj"
                else:
                    est_ratio = ratio()
                    strategy = determine_strategy(vcov, bool(instruments), fe_card, MAX_FE_LEVELS, n0,
                                                  iv_state["k"], est_ratio)

        if strategy == "compress":
            y, X, x_cols, w = cols_cache[0] if cols_cache else columns(keep_int=True)
            res = _fit_compress(eng, y, X, x_cols, fe_keys, fe_cols, cl_keys, cluster_cols, w, vcov, vkind, ssc, n0)
            res.formula, res.fe_cols, res.compression_ratio = formula, fe_cols, est_ratio
        else:
            lf = fe_cols.index(local_fe) if dist else -1
            cl_local = bool(dist and cluster_cols and len(cluster_cols) == 1 and cluster_cols[0] == local_fe)
            w_early = num(weights) if weights is not None else None
            res = _fit_fwl(eng, strategy, get_columns, fe_keys, fe_cols, cl_keys, w_early, vcov, vkind, ssc, demean_tol,
                           max_iter, formula, est_ratio, lf, cl_local, iv_state if instruments else None, fe_card_order,
                       deterministic)
        if profile:
            torch.cuda.synchronize(eng.device)
            res.timings = dict(eng.timers(), wall_ms=(time.perf_counter() - t_start) * 1e3, **eng.counters())
            eng.set_profiling(False)
        return res
    finally:
        up.close()      # every staged host->device copy has completed before the caller's arrays may go away
        if deterministic:
            eng.set_option(capi.OPT_DETERMINISTIC, 0)      # the engine is shared: leave it as it was found


# ------------------------------------------------------------------------------------------------
def _fit_fwl(eng: Engine, strategy, get_columns, fe_keys, fe_cols, cl_keys, w, vcov, vkind, ssc, tol, max_iter,
             formula, est_ratio, local_fe=-1, cl_local=False, iv=None, fe_card=None, deterministic=False) -> LeanFEResult:
    """'alt_proj' (polars_impl.py:468-537), 'demean' (:424-465) and 'ols' (:541-547) + _run_regression.

    ``get_columns()`` -> (y, X, names, w) is called as late as possible: the singleton test and the demean plan
    (entry-list sorts) need only the ids and run while the fp64 columns are still being uploaded.
    """
    F = len(fe_keys)
    if strategy == "demean" and F != 1:
        raise ValueError("Strategy 'demean' requires exactly one FE column.")
    if strategy == "alt_proj" and F == 0:
        raise ValueError("Strategy 'alt_proj' requires FE-cols. Use strategy='ols' instead for OLS without FE.")
    cl_codes = [c.codes for c in cl_keys]
    iterations, converged, fe_dims, absorbed = 0, None, None, 0
    if strategy != "ols":
        codes = [f.codes for f in fe_keys]
        Gs = [f.n_groups for f in fe_keys]
        card = fe_card if fe_card is not None else Gs   # multi-GPU: the shard key's cardinality over ALL ranks
        order = sorted(range(F), key=lambda j: card[j])  # ascending PRE-drop cardinality, stable (:485)
        # singleton filter: one pass, all FEs (polars_impl.py:478-482)
        # group sizes, shared by the filter and the plan: a by-product of the factorisation when it counted
        counts = [f.counts if f.counts is not None else eng.group_counts(f.codes, f.n_groups) for f in fe_keys]
        keep, n_keep, fe_dims = eng.singleton_mask(codes, Gs, local_fe, counts)
        n_obs = eng.last_n_keep_global
        if n_obs == 0:
            raise ValueError("no observations left after dropping singleton groups (every row is alone in one of "
                             f"the fixed effects {fe_cols})")
        if n_obs < eng.last_n_rows_global:       # some rank dropped rows: every rank takes the compaction path
            y, X, x_cols, w = get_columns()
            k = len(X)
            f_in = [y] + X + ([w] if w is not None else [])
            extra = [c for c in cl_codes if not any(c is f for f in codes)]     # a cluster column may be an FE
            f_out, i_out = eng.compact(keep, n_keep, f_in, codes + extra)
            cols, w = f_out[: k + 1], (f_out[k + 1] if w is not None else None)
            new_of = {id(o): n_ for o, n_ in zip(codes + extra, i_out)}
            codes, cl_codes = i_out[:F], [new_of[id(c)] for c in cl_codes]
            plan = eng.demean_plan(codes, Gs, w, local_fe, columns_hint=k + 1)
        else:
            # rows are never permuted: the plan sorts only the int32 code columns (one entry list per FE)
            lazy_ = getattr(get_columns, "lazy", None)
            per_column = lazy_ is not None and strategy == "alt_proj" and lazy_["pending"]()
            hint = 1 if per_column else (len(lazy_["x"]) + 1 if lazy_ is not None else None)
            plan = eng.demean_plan(codes, Gs, w, local_fe, counts, columns_hint=hint)
            cols = None
        del keep, counts
        lazy = getattr(get_columns, "lazy", None)
        if cols is None and lazy is not None and strategy == "alt_proj" and lazy["pending"]():
            # Columns are still crossing PCIe.  Demeaning is independent per column and the stopping rule looks
            # at y only (polars_impl.py:511-525), so: demean y as soon as it lands (this fixes the sweep count
            # T), then give every regressor exactly T sweeps as it arrives -- the same iterates as the
            # all-columns loop, but only the last column's work is left when the upload ends.
            own = get_columns.owned

            def fetch(name):
                t = lazy["get"](name)
                return (t, None) if own(t) else (torch.empty_like(t), [t])    # never overwrite caller-owned data

            ycol, ysrc = fetch(lazy["y"])
            iterations, _, converged = plan.run([ycol], order, tol, max_iter, 3, src=ysrc)
            cols, x_cols = [ycol], list(lazy["x"])
            for name in x_cols:
                c, csrc = fetch(name)
                plan.run([c], order, tol, iterations, iterations + 1, src=csrc)   # T sweeps, no stopping test
                cols.append(c)
            k = len(x_cols)
            absorbed = sum(fe_dims) - F
        else:
            if cols is None:
                y, X, x_cols, w = get_columns()
                k = len(X)
                raw = [y] + X
            else:
                raw = cols
            kind = vkind & 0xFF
            if (iv is None and len(raw) <= plan.FIT_MAX_COLS and (kind != capi.VCOV_CLUSTER or len(cl_codes) == 1)
                    and not (deterministic and kind == capi.VCOV_CLUSTER)
                    and all(t.dtype == torch.float64 and t.is_contiguous() for t in raw)):
                # the usual fit: one library call, the demeaned columns are formed in registers and never written
                # (lfe_demean_plan_fit); df_resid needs only fe_dims, known since the singleton pass
                absorbed = (fe_dims[0] - 1) if strategy == "demean" else (sum(fe_dims) - F)
                p = k + 1
                df_resid = n_obs - p - absorbed
                if kind == capi.VCOV_CLUSTER and cl_local:
                    vkind |= capi.VCOV_CLUSTER_LOCAL
                mi, mn = (1, 2) if strategy == "demean" else (max_iter, 3)
                out = plan.fit(raw, order, tol, mi, mn, vkind, cl_codes[0] if kind == capi.VCOV_CLUSTER else None,
                               cl_keys[0].n_groups if kind == capi.VCOV_CLUSTER else 0, ssc, n_obs, df_resid)
                plan.close()
                iterations, converged = (1, True) if strategy == "demean" else (out["iterations"], out["converged"])
                host = out["packed"].cpu().numpy()
                beta, se = host[1:p], host[p:p + k]
                rss, tss = host[-4], host[-3]
                n_clusters = out["n_clusters"] if kind == capi.VCOV_CLUSTER else None
                r2 = float(1 - rss / tss) if tss > 0 else None
                return LeanFEResult(coefs=dict(zip(x_cols, beta.tolist())), std_errors=dict(zip(x_cols, se.tolist())),
                                    n_obs=n_obs, iterations=iterations, vcov_type=vcov, is_iv=False, n_instruments=None,
                                    n_clusters=n_clusters, df_resid=df_resid, formula=formula, fe_cols=fe_cols,
                                    fe_dims=fe_dims, r_squared=r2, compression_ratio=est_ratio, rss=float(rss),
                                    tss=float(tss), strategy=strategy, converged=converged)
            if cols is None:
                own = getattr(get_columns, "owned", lambda t: False)
                src = [y] + X
                # caller-owned device columns are read in place and the demeaned values written to fresh buffers
                cols = [c if own(c) else torch.empty_like(c) for c in src]
            else:
                src = cols
            src_arg = None if all(a is b for a, b in zip(src, cols)) else src
            if strategy == "demean":
                plan.run(cols, order, tol, 1, 2, src=src_arg)
                iterations, converged = 1, True
                absorbed = fe_dims[0] - 1
            else:
                iterations, _, converged = plan.run(cols, order, tol, max_iter, 3, src=src_arg)
                absorbed = sum(fe_dims) - F
        plan.close()
    else:
        y, X, x_cols, w = get_columns()
        k = len(X)
        cols = [y] + list(X)
        n_obs = cols[0].numel()
    if iv is not None:
        # 2SLS instead of OLS (polars_impl.py:176-198); iv["z_idx"] picks the (demeaned) instrument columns
        k = iv["k"]
        Z = [cols[1 + i] for i in iv["z_idx"]]
        beta, se, rss, df_resid, n_clusters = _regress_iv(eng, cols[0], cols[1:1 + k], Z, w, vkind, cl_codes,
                                                          [c.n_groups for c in cl_keys], ssc, n_obs, absorbed)
        x_cols, n_instr = list(x_cols)[:k], len(Z)
        return LeanFEResult(coefs=dict(zip(x_cols, beta)), std_errors=dict(zip(x_cols, se)), n_obs=n_obs,
                            iterations=iterations, vcov_type=vcov, is_iv=True, n_instruments=n_instr,
                            n_clusters=n_clusters, df_resid=df_resid, formula=formula, fe_cols=fe_cols, fe_dims=fe_dims,
                            r_squared=None, compression_ratio=est_ratio, rss=rss, strategy=strategy,
                            converged=converged)
    if vkind == capi.VCOV_CLUSTER and cl_local:
        vkind |= capi.VCOV_CLUSTER_LOCAL
    # regression on [1 | X~] (polars_impl.py:166-226)
    gram = eng.gram(cols[1:] + [cols[0]], w, add_intercept=True)       # order: 1, x.., y
    p = k + 1
    XtX, Xty = gram[:p, :p].contiguous(), gram[:p, p].contiguous()
    beta_full, XtX_inv, _ = eng.solve_spd(XtX, Xty)
    df_resid = n_obs - p - absorbed
    B = XtX_inv[1:, 1:].contiguous()
    out = eng.resid_vcov(cols, beta_full, B, vkind, cl_codes, [c.n_groups for c in cl_keys], w, ssc, n_obs, df_resid)
    host = torch.cat([beta_full, out["se"], out["stats"]]).cpu().numpy()
    beta, se, (rss, tss) = host[1:p], host[p:p + k], host[p + k:p + k + 2]
    n_clusters = None
    if (vkind & 0xFF) == capi.VCOV_CLUSTER:
        n_clusters = out["n_clusters"][0] if len(cl_keys) == 1 else tuple(out["n_clusters"])
    r2 = float(1 - rss / tss) if tss > 0 else None
    return LeanFEResult(coefs=dict(zip(x_cols, beta.tolist())), std_errors=dict(zip(x_cols, se.tolist())), n_obs=n_obs,
                        iterations=iterations, vcov_type=vcov, is_iv=False, n_instruments=None, n_clusters=n_clusters,
                        df_resid=df_resid, formula=formula, fe_cols=fe_cols, fe_dims=fe_dims, r_squared=r2,
                        compression_ratio=est_ratio, rss=float(rss), tss=float(tss), strategy=strategy,
                        converged=converged)


def _regress_iv(eng: Engine, y, X, Z, w, vkind, cl_codes, cl_G, ssc, n_obs, absorbed):
    """2SLS exactly as the reference runs it (polars_impl.py:176-198, 229-278; common.py:188-287;
    std_errors.py:448-598): EVERY column of X = [1 | x] is projected on the instruments Z (a constant joins Z
    only when Z has fewer columns than X), beta = (Xh'Xh)^-1 Xh'y, residuals are taken with the FITTED Xh, and
    the sandwiches use Xh.  On the device this needs no n x k temporary: with gamma = (Z'WZ)^-1 Z'WX,

        Xh'WXh = gamma' Z'WZ gamma,   Xh'Wy = gamma' Z'Wy,   resid = y - Z (gamma beta),
        V = A V_Z A',  A = (Xh'WXh)^-1 gamma',

    where V_Z is the ordinary meat (HC1 / one-way / multi-way, with the same scalar corrections) of the
    regressor set Z -- one Gram pass over [1, x, z, y] and one lfe_resid_vcov pass over [y, Z].
    -> (beta[k], se[k], rss, df_resid, n_clusters)"""
    k, m = len(X), len(Z)
    n = y.numel()
    add_const = (k + 1) > m and not any(bool(((z - 1.0).abs() <= 1e-8 + 1e-5).all()) for z in Z)   # np.allclose(col, 1)
    mz = m + (1 if add_const else 0)
    if mz < k + 1:
        raise ValueError(f"Under-identified: {mz} instruments for {k + 1} endogenous variables")
    gram = eng.gram(list(X) + list(Z) + [y], w, add_intercept=True)        # order: 1, x.., z.., y
    ix = torch.arange(0, k + 1, device=eng.device)
    iz = torch.tensor(([0] if add_const else []) + list(range(k + 1, k + 1 + m)), device=eng.device)
    iy = k + 1 + m
    ZtZ = gram.index_select(0, iz).index_select(1, iz).contiguous()
    ZtX = gram.index_select(0, iz).index_select(1, ix).contiguous()
    ZtY = gram.index_select(0, iz)[:, iy].contiguous()
    _, ZtZ_inv, _ = eng.solve_spd(ZtZ, ZtY)
    gamma = ZtZ_inv @ ZtX                                                  # first stage, (mz, k+1)
    XhXh = gamma.T @ ZtZ @ gamma
    XhXh = (0.5 * (XhXh + XhXh.T)).contiguous()
    beta_full, XtX_inv, _ = eng.solve_spd(XhXh, (gamma.T @ ZtY).contiguous())
    delta = gamma @ beta_full                                              # resid = y - Z delta
    df_resid = n_obs - (k + 1) - absorbed
    kind = vkind & 0xFF
    if kind == capi.VCOV_IID:
        b_z = delta if add_const else torch.cat([torch.zeros(1, dtype=torch.float64, device=eng.device), delta])
        out = eng.resid_vcov([y] + list(Z), b_z.contiguous(), torch.eye(m, dtype=torch.float64, device=eng.device),
                             capi.VCOV_IID, None, None, w, ssc, n_obs, df_resid)
        V = XtX_inv * (out["stats"][2] / df_resid)                         # std_errors.py:183-210 on the full (X'X)^-1
        n_clusters = None
    else:
        Zc = ([torch.ones(n, dtype=torch.float64, device=eng.device)] if add_const else []) + list(Z)
        b_z = torch.cat([torch.zeros(1, dtype=torch.float64, device=eng.device), delta]).contiguous()
        out = eng.resid_vcov([y] + Zc, b_z, torch.eye(mz, dtype=torch.float64, device=eng.device), kind, cl_codes, cl_G,
                             w, ssc, n_obs, df_resid)
        A = XtX_inv @ gamma.T
        V = A @ out["vcov"] @ A.T
        n_clusters = None
        if kind == capi.VCOV_CLUSTER:
            n_clusters = out["n_clusters"][0] if len(cl_G) == 1 else tuple(out["n_clusters"])
    se_full = torch.sqrt(torch.clamp(torch.diagonal(V), min=0.0))
    host = torch.cat([beta_full, se_full, out["stats"][:1]]).cpu().numpy()
    p = k + 1
    return host[1:p].tolist(), host[p + 1:2 * p].tolist(), float(host[2 * p]), df_resid, n_clusters


def _fit_compress(eng: Engine, y, X, x_cols, fe_keys, fe_cols, cl_keys, cluster_cols, w, vcov, vkind, ssc, n_obs):
    """YOCO (compress.py:1049-1174): device group-by, then FWL on the compressed rows.

    Multi-GPU: each rank groups its own rows -- all key columns coded in one shared code space --
    the partial tables are all-gathered and re-aggregated (``lfe_compress_merge``), and the small solve on the
    merged table is repeated identically on every rank (no collectives inside)."""
    k, F = len(X), len(fe_keys)
    if k == 0:
        raise ValueError("strategy='compress' needs at least one regressor")
    dist = eng.world > 1
    if dist:
        n_obs = int(eng.allreduce(torch.tensor([float(n_obs)], dtype=torch.float64, device=eng.device)).item())
    group_names = x_cols + fe_cols
    xk = [None] * k
    x_order = capi.ORDER_SORTED | (capi.ORDER_GLOBAL if dist else 0)
    # uint8 regressors (dummies, treatment x period indicators) are grouped on their raw values, 1 byte per row:
    # no dictionary pass, value of code c is c itself
    raw = [j for j, c in enumerate(X) if c.dtype == torch.uint8 and c.is_contiguous()]
    if raw:
        for j, (_, hi) in zip(raw, eng.key_ranges([X[j] for j in raw])):
            radix = max(int(hi), 0) + 1
            xk[j] = (X[j], radix, torch.arange(radix, dtype=torch.float64, device=eng.device))
    for j, c in enumerate(X):
        if xk[j] is None:
            codes, G, uniq = eng.factorize(c, x_order, want_uniques=True, uniques_cap=1 << 20)
            xk[j] = (codes, G, uniq.to(torch.float64))
    kc = [(c, G) for c, G, _ in xk] + [(f.codes, f.n_groups) for f in fe_keys]
    cl_idx = []
    for name, ck in zip(cluster_cols or [], cl_keys):
        if name in group_names:
            cl_idx.append(group_names.index(name))
        else:
            group_names = group_names + [name]
            kc.append((ck.codes, ck.n_groups))
            cl_idx.append(len(kc) - 1)
    comp = eng.compress([c for c, _ in kc], [g for _, g in kc], y, w)
    if dist:
        comp = eng.compress_merge(comp, [g for _, g in kc])
    return _solve_compressed(eng, comp, xk, kc, cl_idx, x_cols, fe_keys, k, F, vcov, vkind, ssc, n_obs, dist)


def _solve_compressed(eng: Engine, comp, xk, kc, cl_idx, x_cols, fe_keys, k, F, vcov, vkind, ssc, n_obs, dist=False):
    """Weighted FWL + YOCO standard errors on the compressed table (compress.py:503-1042, 1105-1174).

    ``dist``: the table is the merged one, identical on every rank; the solve runs rank-locally on each of
    them and rank 0's numbers are then shared, so all ranks return bit-identical results."""
    with (eng.replicated() if dist else contextlib.nullcontext()):
        G = comp["n_compressed"]
        kout = comp["key_codes"]
        # regressor values of each compressed row = uniques[code]; group means; weights n_g
        Xg = [xk[j][2].index_select(0, kout[j].long()).contiguous() for j in range(k)]
        nn, sy, syy = comp["n"], comp["sum_y"], comp["sum_y_sq"]
        cols = [(sy / nn).contiguous()] + Xg
        fe_codes = [kout[k + j] for j in range(F)]
        fe_G = [f.n_groups for f in fe_keys]
        fe_dims = tuple(fe_G) if F else None
        # no FE: the intercept alone is absorbed, i.e. one "fixed effect" with a single level
        dm_codes = fe_codes if F else [torch.zeros(G, dtype=torch.int32, device=eng.device)]
        dm_G = fe_G if F else [1]
        scale = float(torch.max(torch.abs(cols[0])).item()) + 1.0
        order = sorted(range(len(dm_G)), key=lambda j: dm_G[j])
        sweeps, _, converged = eng.demean(cols, dm_codes, dm_G, order, nn, 1e-14 * scale, 20000, 1, strict_check=True)
        if not converged:
            # the reference solves the dummy-variable system directly (compress.py:700-747); an iterative solve that
            # stops short must not pass silently
            warnings.warn(f"strategy='compress': the weighted FWL solve on the {G} compressed rows did not reach its "
                          f"fixed point in {sweeps} sweeps (poorly connected fixed effects); coefficients and standard "
                          "errors may be inaccurate", RuntimeWarning, stacklevel=3)
        gram = eng.gram(cols[1:] + [cols[0]], nn, add_intercept=False)
        try:
            beta, B, _ = eng.solve_spd(gram[:k, :k].contiguous(), gram[:k, k].contiguous())
        except capi.LfeError as e:
            if "singular" not in str(e):
                raise
            # Singular design (e.g. a dummy that is never 1): the reference's Cholesky raises and it falls back to
            # the minimum-norm solution, np.linalg.lstsq + np.linalg.pinv of the normal equations
            # (compress.py:742-745).  Same two LAPACK calls here, on the k x k FWL system (k <= 512, host): for
            # all-zero / duplicated regressors this is the x-block of the reference's full-system answer; a
            # regressor collinear with the FE dummies gets coefficient 0 here (the reference spreads it over the
            # dummies it then discards).
            warnings.warn("strategy='compress': the regressors are collinear after absorbing the fixed effects; "
                          "returning the minimum-norm solution (as the reference's lstsq/pinv fallback)",
                          RuntimeWarning, stacklevel=3)
            g_host = gram.cpu().numpy()
            beta_h = np.linalg.lstsq(g_host[:k, :k], g_host[:k, k], rcond=None)[0]
            B_h = np.linalg.pinv(g_host[:k, :k])
            beta = torch.from_numpy(np.ascontiguousarray(beta_h)).to(eng.device)
            B = torch.from_numpy(np.ascontiguousarray(B_h)).to(eng.device)
        p = 1 + k + sum(g - 1 for g in fe_G)
        df_resid = n_obs - p
        out = eng.yoco_vcov(cols, beta, B, nn, sy, syy, vkind, [kout[i] for i in cl_idx], [kc[i][1] for i in cl_idx],
                            ssc, n_obs, df_resid)
        dev = torch.cat([beta, out["se"], out["rss"]])
    if dist:
        if eng.rank != 0:
            dev.zero_()
        eng.allreduce(dev)
    host = dev.cpu().numpy()
    n_clusters = None
    if vkind == capi.VCOV_CLUSTER:
        n_clusters = out["n_clusters"][0] if len(cl_idx) == 1 else tuple(out["n_clusters"])
    return LeanFEResult(coefs=dict(zip(x_cols, host[:k].tolist())), std_errors=dict(zip(x_cols, host[k:2 * k].tolist())),
                        n_compressed=G, n_obs=n_obs, vcov_type=vcov, df_resid=df_resid, rss=float(host[2 * k]),
                        n_clusters=n_clusters, fe_dims=fe_dims, strategy="compress", iterations=0, converged=converged,
                        solver_sweeps=sweeps)

"""``LeanFEResult`` -- the result object of the reference, kept field for field.

Contract (reference python/leanfe/result.py:11-296): attributes ``coefs``, ``std_errors``, ``t_stats``,
``p_values``, ``n_obs``, ``iterations``, ``n_compressed``, ``compression_ratio``, ``vcov_type``, ``is_iv``,
``n_instruments``, ``n_clusters``, ``df_resid``, ``r_squared``, ``r_squared_within``, ``rss``, ``tss``,
``formula``, ``fe_cols``, ``fe_dims``; printable table; ``coef/se/tstat/pvalue/confint``; dict-like access
over ``to_dict()``.  Supersets (harmless to reference users): ``strategy``, ``converged``, ``timings``, ``solver_sweeps``; the
``"coefficients"`` key is accepted as an alias of ``"coefs"`` (the reference's docs/tests still use it);
``_vcov_description`` no longer breaks on multi-way cluster tuples.
"""
from __future__ import annotations

import math
from typing import Any

from scipy import stats as _st

_RULE = "=" * 70
_THIN = "-" * 70


class PrettyInt(int):
    """int whose repr groups digits with underscores."""

    def __repr__(self) -> str:
        return format(int(self), "_")


def wrap_int(value: Any):
    return None if value is None else PrettyInt(value)


def _stars(p: float) -> str:
    for cut, mark in ((0.001, "***"), (0.01, "**"), (0.05, "*"), (0.1, ".")):
        if p < cut:
            return mark
    return ""


class LeanFEResult:
    def __init__(self, coefs, std_errors, n_obs, vcov_type, iterations=0, n_compressed=None, compression_ratio=None,
                 is_iv=False, n_instruments=None, n_clusters=None, df_resid=None, r_squared=None,
                 r_squared_within=None, rss=None, tss=None, formula=None, fe_cols=None, fe_dims=None,
                 strategy=None, converged=None, timings=None, solver_sweeps=None):
        self.coefs = coefs
        self.std_errors = std_errors
        self.n_obs = n_obs
        self.vcov_type = vcov_type
        self.iterations = iterations
        self.n_compressed = n_compressed
        self.compression_ratio = compression_ratio
        self.is_iv = is_iv
        self.n_instruments = n_instruments
        self.n_clusters = n_clusters
        self.df_resid = df_resid or (n_obs - len(coefs))
        self.r_squared = r_squared
        self.r_squared_within = r_squared_within
        self.rss = rss
        self.tss = tss
        self.formula = formula
        self.fe_cols = fe_cols or []
        self.fe_dims = fe_dims
        self.strategy = strategy
        self.converged = converged
        self.timings = timings
        self.solver_sweeps = solver_sweeps   # strategy='compress': sweeps of the weighted FWL solve (iterations stays 0)
        self.t_stats, self.p_values = {}, {}
        for name, b in coefs.items():
            s = std_errors[name]
            if s > 0:
                t = b / s
                self.t_stats[name] = t
                self.p_values[name] = 2 * (1 - _st.t.cdf(abs(t), self.df_resid))
            else:
                self.t_stats[name] = math.nan
                self.p_values[name] = math.nan

    # -- printing ----------------------------------------------------------------------------
    def _vcov_description(self) -> str:
        if self.vcov_type == "iid":
            return "IID"
        if self.vcov_type in ("HC1", "hc1"):
            return "Heteroskedasticity-robust (HC1)"
        if self.vcov_type == "cluster":
            if isinstance(self.n_clusters, tuple):
                return "Clustered (" + " × ".join(f"{c:,}" for c in self.n_clusters) + " clusters)"
            return f"Clustered ({self.n_clusters:,} clusters)"
        return self.vcov_type

    def __repr__(self) -> str:
        return f"LeanFEResult(n_obs={self.n_obs:_}, n_coef={len(self.coefs)}, vcov='{self.vcov_type}')"

    def __str__(self) -> str:
        out = ["", _RULE, " " * 25 + "leanfe Regression Results", _RULE]
        if self.formula:
            out += [f"Formula:      {self.formula}", f"Observations: {self.n_obs:_}"]
        if self.fe_cols:
            if isinstance(self.fe_cols, dict):
                out += [f"Fixed Effect ({fe}): {cnt:_} groups" for fe, cnt in self.fe_cols.items()]
            else:
                out.append("Fixed Effects: " + ", ".join(self.fe_cols))
                if self.fe_dims:
                    out.append("FE Dimensions: " + " × ".join(f"{d:_}" for d in self.fe_dims))
        if self.r_squared_within is not None:
            out.append(f"R² (within):  {self.r_squared_within:.4f}")
        out.append(f"Std. Errors:  {self._vcov_description()}")
        if self.n_clusters:
            shown = (" × ".join(f"{c:_}" for c in self.n_clusters) if isinstance(self.n_clusters, tuple)
                     else f"{self.n_clusters:_}")
            out.append(f"Clusters:     {shown}")
        out += [_THIN, f"{'Variable':<20} {'Estimate':>12} {'Std.Err':>12} {'t-stat':>10} {'p-value':>10}", _THIN]
        for name, b in self.coefs.items():
            label = name if len(name) <= 20 else name[:18] + ".."
            p = self.p_values[name]
            out.append(f"{label:<20} {b:>12.6f} {self.std_errors[name]:>12.6f} {self.t_stats[name]:>10.3f} "
                       f"{p:>9.4f}{_stars(p)}")
        out += [_THIN, "Signif. codes: 0 '***' 0.001 '**' 0.01 '*' 0.05 '.' 0.1", _RULE, ""]
        return "\n".join(out)

    def summary(self) -> str:
        return str(self)

    # -- accessors ---------------------------------------------------------------------------
    @staticmethod
    def _pick(table: dict, var):
        return dict(table) if var is None else table.get(var)

    def coef(self, var=None):
        return self._pick(self.coefs, var)

    def se(self, var=None):
        return self._pick(self.std_errors, var)

    def tstat(self, var=None):
        return self._pick(self.t_stats, var)

    def pvalue(self, var=None):
        return self._pick(self.p_values, var)

    def confint(self, level: float = 0.95) -> dict:
        crit = _st.t.ppf(1 - (1 - level) / 2, self.df_resid)
        return {v: (b - crit * self.std_errors[v], b + crit * self.std_errors[v]) for v, b in self.coefs.items()}

    # -- dict-like ---------------------------------------------------------------------------
    def to_dict(self) -> dict:
        return {
            "formula": self.formula, "coefs": self.coefs, "std_errors": self.std_errors, "t_stats": self.t_stats,
            "p_values": self.p_values, "n_obs": wrap_int(self.n_obs), "n_compressed": wrap_int(self.n_compressed),
            "compression_ratio": self.compression_ratio, "fe_cols": self.fe_cols, "fe_dims": self.fe_dims,
            "iterations": self.iterations, "vcov_type": self.vcov_type, "is_iv": self.is_iv,
            "n_instruments": self.n_instruments, "n_clusters": self.n_clusters, "df_resid": wrap_int(self.df_resid),
            "r_squared_within": self.r_squared_within,
        }

    def _extended(self) -> dict:
        d = self.to_dict()
        d.update(coefficients=self.coefs, strategy=self.strategy, converged=self.converged, r_squared=self.r_squared,
                 rss=self.rss, tss=self.tss, timings=self.timings)
        return d

    def __getitem__(self, key):
        return self._extended()[key]

    def __contains__(self, key):
        return key in self._extended()

    def get(self, key, default=None):
        return self._extended().get(key, default)

    def keys(self):
        return self.to_dict().keys()

    def values(self):
        return self.to_dict().values()

    def items(self):
        return self.to_dict().items()

"""Strategy rule of the reference (python/leanfe/compress.py:96-184), host-side metadata only."""
from __future__ import annotations

MAX_FE_LEVELS = 10_000            # polars_impl.py:24 / compress.py:40
ASSUMED_FWL_SWEEPS = 10           # compress.py:41
_KNOWN_VCOV = ("iid", "hc1", "cluster")


def determine_strategy(vcov: str, has_instruments: bool, fe_cardinality: dict | None = None,
                       max_fe_levels: int = MAX_FE_LEVELS, n_obs: int | None = None, n_x_cols: int | None = None,
                       estimated_compression_ratio: float | None = None) -> str:
    """'compress' (YOCO) or 'alt_proj' (FWL), decided exactly as the reference does."""
    if has_instruments or vcov.lower() not in _KNOWN_VCOV:
        return "alt_proj"
    if fe_cardinality is None:
        return "compress"
    levels = list(fe_cardinality.values())
    total, largest = sum(levels), max(levels, default=0)
    if largest > max_fe_levels or total > 2 * max_fe_levels:
        return "alt_proj"
    if estimated_compression_ratio is None or n_obs is None:
        return "compress"
    n_groups = int(n_obs * estimated_compression_ratio)
    cost_yoco = n_obs + n_groups * total + total ** 2
    cost_fwl = ASSUMED_FWL_SWEEPS * len(levels) * n_obs
    return "compress" if cost_yoco < cost_fwl else "alt_proj"

"""leanfe_b200 -- B200-native (sm_100a) implementation of the leanfe estimation hot path.

Export list follows the reference's ``python/leanfe/__init__.py:3-16`` (minus the DuckDB backend);
``leanfe_polars`` is kept as an alias of the device backend so reference call sites keep working.
"""
from .formula import FormulaComponents, parse_formula
from .result import LeanFEResult
from .strategy import determine_strategy


def leanfe(*args, **kwargs):
    from .api import leanfe as _leanfe
    return _leanfe(*args, **kwargs)


def leanfe_b200(*args, **kwargs):
    from .api import leanfe_b200 as _impl
    return _impl(*args, **kwargs)


leanfe_polars = leanfe_b200

__version__ = "0.1.0"
__all__ = ["leanfe", "leanfe_b200", "leanfe_polars", "parse_formula", "FormulaComponents", "LeanFEResult",
           "determine_strategy"]

"""R-style formula grammar of the reference front end.

Behaviour mirrors ``parse_formula`` in the reference (python/leanfe/common.py:51-181):

    y ~ x1 + x2                       no fixed effects
    y ~ x1 + x2 | fe1 + fe2           fixed effects
    y ~ x1 + i(region)                factor dummies, first sorted level is the reference
    y ~ x1 + i(region, ref=R1)        explicit reference level (quotes optional)
    y ~ x1 + treat:i(region)          heterogeneous slopes
    y ~ x1 | fe1 | z1 + z2            third part = instruments (IV/2SLS, estimated by api._regress_iv)
"""
from __future__ import annotations

import re
from typing import NamedTuple


class FormulaComponents(NamedTuple):
    y_col: str
    x_cols: list
    fe_cols: list
    factor_vars: list       # [(var, ref | None)]
    interactions: list      # [(var, factor, ref | None)]
    instruments: list


_REF = r"(?:\s*,\s*ref\s*=\s*[\"']?([^\"')\s]+)[\"']?)?"
_FACTOR = re.compile(r"i\((\w+)" + _REF + r"\)")
_INTERACT = re.compile(r"(\w+):i\((\w+)" + _REF + r"\)")


def _split_plus(text: str) -> list:
    return [tok for tok in (piece.strip() for piece in text.split("+")) if tok]


def _classify(term: str, out: FormulaComponents):
    if term.endswith(")") and ":i(" in term:
        m = _INTERACT.match(term)
        if m is None:
            raise ValueError(f"Invalid interaction syntax: {term}")
        out.interactions.append((m.group(1), m.group(2), m.group(3)))
    elif term.endswith(")") and term.startswith("i("):
        m = _FACTOR.match(term)
        if m is None:
            raise ValueError(f"Invalid i() syntax: {term}. Use i(var) or i(var, ref=value)")
        out.factor_vars.append((m.group(1), m.group(2)))
    else:
        out.x_cols.append(term)


def parse_formula(formula: str) -> FormulaComponents:
    sections = [sec.strip() for sec in formula.split("|")]
    if len(sections) > 3:
        raise ValueError("Formula has too many parts. Use: 'y ~ x' or 'y ~ x | fe' or 'y ~ x | fe | z' (IV)")
    sides = sections[0].split("~")
    if len(sides) != 2:
        raise ValueError("Formula must have exactly one '~' separating y and x variables")
    out = FormulaComponents(sides[0].strip(), [], [], [], [], [])
    for term in _split_plus(sides[1]):
        _classify(term, out)
    if len(sections) >= 2:
        out.fe_cols.extend(_split_plus(sections[1]))
    if len(sections) == 3:
        out.instruments.extend(_split_plus(sections[2]))
    return out

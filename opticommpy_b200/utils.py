"""Boundary type of the drop-in: the ``parameters`` attribute bag.

Mirrors the public behaviour of ``optic.utils.parameters`` (optic/utils.py:29-146): a plain
object whose attributes are read with ``getattr(param, name, default)`` by every function, plus
``view`` / ``table`` / ``latex_table`` / ``copy`` helpers.  Any object with attributes works as
a parameter bag (duck typing) — including the reference's own class.
"""
from __future__ import annotations

import copy as _copy
import math

import numpy as np

_SI_PREFIX = {-12: "p", -9: "n", -6: "µ", -3: "m", 0: "", 3: "k", 6: "M", 9: "G", 12: "T", 15: "P"}


class parameters:
    """Struct-like bag of parameters."""

    def view(self):
        """Print every attribute; large numbers in scientific notation."""
        for name, value in vars(self).items():
            big = isinstance(value, (int, float)) and value > 10000
            print(f"{name}: {value:.2e}" if big else f"{name}: {value}")

    def to_engineering_notation(self, value):
        """Format ``value`` with an SI prefix when it is very large or very small."""
        if not isinstance(value, (int, float)):
            return value
        mag = abs(value)
        if not (mag >= 10000 or 0 < mag < 0.0001):
            return value
        exp3 = int(math.floor(math.log10(mag))) // 3 * 3
        return f"{value / 10 ** exp3:.1f} {_SI_PREFIX.get(exp3, '')}"

    def _rows(self):
        for name, value in vars(self).items():
            if isinstance(value, (list, tuple, np.ndarray)):
                yield name, "Array"
            else:
                yield name, self.to_engineering_notation(value)

    def table(self):
        """Print a Markdown table of the parameters."""
        lines = ["| Parameter Name | Value |", "|----------------|-----------------|"]
        lines += [f"| {n} | {v} |" for n, v in self._rows()]
        return print("\n".join(lines) + "\n")

    def latex_table(self):
        """Print a LaTeX tabular of the parameters."""
        out = ["\\begin{tabular}{|c|c|}", "\\hline", "Parameter Name & Value \\\\", "\\hline"]
        for n, v in self._rows():
            out += [f"{n} & {v} \\\\", "\\hline"]
        out.append("\\end{tabular}")
        return print("\n".join(out))

    def copy(self):
        """Deep copy."""
        return _copy.deepcopy(self)

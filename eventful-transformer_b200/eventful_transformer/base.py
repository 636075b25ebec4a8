"""
Module infrastructure: operation counters, state reset and sub-module filtering.

API mirror of the reference's eventful_transformer/base.py (Counts :7-78,
ExtendedModule :81-149, numeric_tuple :152, dict_csv_header/line/dict_string
:165-195).  Counters stay host-side Python integers: they are bookkeeping, not
device work.
"""

import sys
from collections import defaultdict
from numbers import Number

from torch import nn


def _sorted_items(d):
    return [(key, d[key]) for key in sorted(d.keys())]


def dict_csv_header(x):
    """Comma-joined sorted keys."""
    return ",".join(str(key) for key, _ in _sorted_items(x))


def dict_csv_line(x):
    """Comma-joined values in sorted-key order, %g formatted."""
    return ",".join(format(value, "g") for _, value in _sorted_items(x))


def dict_string(x, indent=4, value_format=".4g"):
    """Aligned `key: value` lines in sorted-key order."""
    width = max(len(str(key)) for key in x.keys()) + 1
    pad = " " * indent
    rows = []
    for key, value in _sorted_items(x):
        label = f"{key}:"
        rows.append(f"{pad}{label:<{width}} {format(value, value_format)}")
    return "\n".join(rows)


class Counts(defaultdict):
    """A dict of named operation counts with element-wise arithmetic (missing keys count as 0)."""

    def __init__(self, *args, **kwargs):
        if args or kwargs:
            super().__init__(*args, **kwargs)
        else:
            super().__init__(int)

    def _combined(self, other, sign):
        out = self.copy()
        if isinstance(other, Counts):
            for key, value in other.items():
                out[key] += sign * value
        else:
            for key in out:
                out[key] += sign * other
        return out

    def __add__(self, other):
        return self._combined(other, 1)

    __radd__ = __add__

    def __sub__(self, other):
        return self._combined(other, -1)

    def __neg__(self):
        out = self.copy()
        for key in out:
            out[key] = -out[key]
        return out

    def __rsub__(self, other):
        return (-self)._combined(other, 1)

    def __mul__(self, factor):
        out = self.copy()
        for key in out:
            out[key] *= factor
        return out

    __rmul__ = __mul__

    def __truediv__(self, divisor):
        return self * (1.0 / divisor)

    def csv_header(self):
        return dict_csv_header(self)

    def csv_line(self):
        return dict_csv_line(self)

    def pretty_print(self, indent=4, value_format=".3e", file=sys.stdout, flush=False):
        print(dict_string(self, indent, value_format), file=file, flush=flush)


class ExtendedModule(nn.Module):
    """nn.Module with operation counting, state reset and typed sub-module enumeration."""

    def __init__(self):
        super().__init__()
        self.count_mode = False
        self.counts = Counts()

    # -- enumeration -----------------------------------------------------------------
    def modules_of_type(self, module_type):
        return (m for m in self.modules() if isinstance(m, module_type))

    def extended_modules(self):
        return self.modules_of_type(ExtendedModule)

    # -- counting ----------------------------------------------------------------------
    def counting(self, mode=True):
        for m in self.extended_modules():
            m.count_mode = mode

    def no_counting(self):
        self.counting(mode=False)

    def clear_counts(self):
        for m in self.extended_modules():
            m.counts.clear()

    def total_counts(self):
        return sum(m.counts for m in self.extended_modules())

    # -- state -------------------------------------------------------------------------
    def reset(self):
        for m in self.extended_modules():
            m.reset_self()

    def reset_self(self):
        """Hook: drop this module's own temporal state (children are visited by reset())."""


def numeric_tuple(x, length):
    """Scalar -> tuple of `length` copies; anything else -> tuple(x)."""
    if isinstance(x, (bool, Number)):
        return (x,) * length
    return tuple(x)

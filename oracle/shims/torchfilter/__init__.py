"""ORACLE (test infrastructure): restatement of the torchfilter API surface the reference uses.

Follows SURVEY.md Appendix A; upstream source is absent (ref: setup.py:14, un-pinned tarball).
PARITY UNPINNED -- see oracle/__init__.py.
"""
from . import base, filters, types  # noqa: F401

__all__ = ["base", "filters", "types"]

"""ORACLE (test infrastructure): the slices of fannypack the hot path touches (Appendix A.8)."""
from . import nn, utils  # noqa: F401

"""Drop-in for the slice of ``torchfilter`` the reference uses, with the
filter recursion running in hand-written sm_100a CUDA kernels (``libmmf_b200.so``)."""
from . import base, data, filters, train, types  # noqa: F401

"""Drop-in for the two corners of ``fannypack`` that sit on the filtering hot path
(SURVEY.md Appendix A.8): ``nn.resblocks`` and ``utils.SliceWrapper``."""
from . import nn, utils  # noqa: F401

"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement of the filtering recursion that brentyi/multimodalfilter runs:

* ``oracle/shims/torchfilter``  -- restatement of the *external, un-vendored* dependency
  ``torchfilter @ stanford-iprl-lab/torchfilter/tarball/master`` (ref: setup.py:12-15; floating
  HEAD, no pin).  Its source is absent from ``/root/reference`` and it cannot be installed here
  (no network), so the restatement follows SURVEY.md Appendix A (A.1-A.6) and the reference's own
  call sites (ref: crossmodal/push_models/pf.py:14-27, crossmodal/door_models/kf.py:14-28,
  crossmodal/eval_helpers.py:125-142, crossmodal/base_models/crossmodal_kf.py:169-206).
  **PARITY UNPINNED at this boundary**: the reference ships no test, golden vector or fixture
  for it.  What we pin instead is listed in DESIGN.md ("Oracle pins").
* ``oracle/shims/fannypack``    -- the two pieces of ``fannypack`` that sit inside the hot MLPs
  (``nn.resblocks.Linear/Conv2d``) and at the API boundary (``utils.SliceWrapper``), Appendix A.8.
* ``oracle/crossmodal_port``    -- independent restatement of ref: crossmodal/base_models/*,
  crossmodal/{push,door}_models/* (architectures + fusion math).  This half IS pinned: when
  ``/root/reference`` is mounted, ``oracle/make_golden.py`` imports the reference's own
  ``crossmodal`` package on top of the shims and writes ``tests/golden/*.npz``; the port, and the
  CUDA product, are checked against those fixtures.
* ``oracle/pinned``             -- C restatement (gcc) of the *pinned* resampling arithmetic
  (exp polynomial, CDF summation orders, lower-bound search) that defines "bit-exact resampling".

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs -- as the checker / CPU baseline only.
"""

import os
import sys

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install_shims() -> None:
    """Make ``import torchfilter`` / ``import fannypack`` resolve to the oracle shims.

    Used by ``oracle/make_golden.py`` (to run the reference's own ``crossmodal`` package on CPU)
    and by the oracle port itself.  Refuses to run if the product's drop-in was installed under
    the same names in this process (the two must never be mixed).
    """
    for name in ("torchfilter", "fannypack"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "")).startswith(_SHIMS):
            raise RuntimeError(
                f"'{name}' is already imported from {getattr(mod, '__file__', '?')}; "
                "the oracle shims must not be mixed with another implementation in one process"
            )
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)

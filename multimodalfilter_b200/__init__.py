"""B200-native (sm_100a) differentiable filtering recursion behind the torchfilter / crossmodal API.

    import multimodalfilter_b200 as mmf
    mmf.install()                      # `import torchfilter`, `import fannypack` now resolve here
    import crossmodal                  # the reference's own package runs unchanged on top

or use the mirrored model classes directly: ``mmf.crossmodal.push_models.PushCrossmodalParticleFilter``.
Nothing here imports ``oracle/`` and nothing falls back to the CPU: the recursion runs in
``libmmf_b200.so`` or raises.
"""
import sys

__version__ = "0.1.0"


def install() -> None:
    """Register this package's drop-ins under the names the reference imports
    (ref: crossmodal/push_models/pf.py:5-7: ``import torchfilter``, ``from fannypack.nn import resblocks``)."""
    from . import fannypack, torchfilter

    for alias, pkg in (("torchfilter", torchfilter), ("fannypack", fannypack)):
        existing = sys.modules.get(alias)
        if existing is not None and existing is not pkg:
            raise RuntimeError(f"'{alias}' is already imported from {getattr(existing, '__file__', '?')}")
        prefix = pkg.__name__
        for name, mod in list(sys.modules.items()):
            if name == prefix or name.startswith(prefix + "."):
                sys.modules[alias + name[len(prefix):]] = mod


def __getattr__(name):
    if name in ("torchfilter", "fannypack", "crossmodal", "ops", "fused", "synthetic"):
        import importlib

        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)

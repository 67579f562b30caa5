"""pfnl_b200 - B200 (sm_100a) implementation of the PFNL 4x multi-frame forward hot path.

`PFNL`, `Engine`, `weights`, `dist` are loaded on first use; touching any of them loads
libpfnl_b200.so (built in-tree by `python -m pfnl_b200.build`).  There is no fallback
implementation - a missing or stale library is an ImportError.
"""
import importlib

__all__ = ["PFNL", "Engine", "weights", "dist", "build"]

_LAZY = {"PFNL": ("model", "PFNL"), "Engine": ("model", "Engine")}
_MODULES = {"weights", "dist", "build", "model", "_lib"}


def __getattr__(name):
    if name in _LAZY:
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    if name in _MODULES:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

"""stable-ts word-timestamp hot path, B200-native (sm_100a).  See DESIGN.md.

Public surface mirrors the reference's: ``load_model()`` -> model with ``align`` / ``transcribe`` / ``refine`` and the
plugin closures the reference's ``Aligner`` / ``Refiner`` accept.  Heavy imports are lazy so that CPU-only tooling
(ABI test, docs) can import the package without CUDA.
"""
__version__ = "0.1.0"


def __getattr__(name):
    import importlib
    lazy = {
        "load_model": ("stable_ts_b200.api", "load_model"),
        "B200Whisper": ("stable_ts_b200.model", "B200Whisper"),
    }
    if name in lazy:
        mod, attr = lazy[name]
        return getattr(importlib.import_module(mod), attr)
    raise AttributeError(name)

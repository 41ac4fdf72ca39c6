"""qrdm_b200 — B200-native (sm_100a) drop-in for the factorisation hot path of mdessole/qrdm.

``qrdm_b200.QRDM`` mirrors the reference's CPython extension; ``qrdm_b200.dgeqrdm`` is a
convenience wrapper returning the outputs as a dict; ``qrdm_b200.dgeqrdm_device`` runs on a
device-resident matrix.  All of them go through the C ABI of ``libqrdm_b200.so``
(include/qrdm_b200.h).  ``qrdm_b200.generators`` (pure NumPy) is importable without the library;
everything else fails loudly if it has not been built — there is no CPU fallback.
"""
from . import generators  # noqa: F401


def __getattr__(name):
    # lazy so that `from qrdm_b200 import generators` works before the library is built
    if name in ("dgeqrdm", "dgeqrdm_device", "dgeqrdm_batched", "dgeqrdm_batched_device", "dormqr", "dormqr_device", "low_rank", "stats",
                "set_profile", "fp64_peak", "copy_gbs", "dgeqp3", "dgeqp3_device"):
        from . import api
        return getattr(api, name)
    if name in ("QRDM", "api", "_lib", "sharded"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)

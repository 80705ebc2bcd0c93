"""pyrecode_b200 -- B200-native implementation of pyReCoDe's per-frame reduce-and-compress hot path and of the
matching decompress-and-unpack read path, behind the reference's own API:

    pyrecode_b200.recode_writer.ReCoDeWriter / print_run_metrics
    pyrecode_b200.recode_reader.ReCoDeReader / merge_parts
    pyrecode_b200.params.InputParams / InitParams
    pyrecode_b200.recode_header.ReCoDeHeader, pyrecode_b200.structures.ReCoDeStructures, pyrecode_b200.misc
    pyrecode_b200.c_recode.Reader           (shim of the reference's native extension)

Importing this package does not import torch or touch the GPU; the first writer.start() / reader decode does, and
fails loudly when the CUDA library or device is missing (no CPU fallback).
"""
import sys

__all__ = ['install_as_pyrecode']


def install_as_pyrecode():
    """Alias this package as `pyrecode` (and the shim as `c_recode`) so unmodified user code such as
    `from pyrecode.recode_writer import ReCoDeWriter` picks up the GPU implementation."""
    import importlib
    pkg = sys.modules[__name__]
    sys.modules.setdefault('pyrecode', pkg)
    for sub in ('misc', 'params', 'recode_header', 'structures', 'recode_writer', 'recode_reader', 'c_recode'):
        mod = importlib.import_module(__name__ + '.' + sub)
        sys.modules.setdefault('pyrecode.' + sub, mod)
    sys.modules.setdefault('c_recode', sys.modules[__name__ + '.c_recode'])
    return pkg

"""mlsgpu_b200 — B200-native reconstruction core with the mlsgpu surface.

The product is ``libmlsb200.so`` (hand-written CUDA for sm_100a behind the C ABI of
``include/mlsb200.h``).  This package is the Python binding used by the tests and by
``bench.py``; the C++ mirror of the reference classes lives in ``host/mlsb200.hpp``.

There is no CPU fallback: importing works anywhere, but every compute call needs the built
library and a CUDA device and raises :class:`MlsbError` otherwise.
"""
from .capi import (MlsbError, SPLAT_DTYPE, Worker, WorkerConfig, WorkerGroup, device_count, lib, library_path,
                   marching_tables, measure_fp32_peak, version)

__all__ = ["MlsbError", "SPLAT_DTYPE", "Worker", "WorkerConfig", "WorkerGroup", "device_count", "lib", "library_path",
           "marching_tables", "measure_fp32_peak", "version"]

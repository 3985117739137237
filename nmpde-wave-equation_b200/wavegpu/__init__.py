"""Python test/bench bindings of libwavegpu.so (ctypes over the C ABI of include/wavegpu.h).

The product is the shared library and the C++ host classes; this package only lets the pytest
suite and bench.py drive the same C entry points.  It contains no numerics."""
from .api import WaveError, WaveSolver, lib, library_path, partition_plan, cell_dofs  # noqa: F401
from .problems import problem, NAMES  # noqa: F401

"""B200-native store for the cube-transform hot path of olap-in-memory.

Public surface = the reference's (/root/reference/src/index.js:1-6):
Cube, GenericDimension, TimeDimension, getParser.  The per-measure store
behind Cube is GpuStore (device-resident, hand-written sm_100a kernels reached
through the C ABI in include/olap_gpu.h); it is imported lazily so that the
host-side classes can be used for dimension work without a GPU, but any store
operation without the CUDA library raises."""
from .cube import Cube
from .dimension import CatchAll, GenericDimension, TimeDimension, TimeSlot
from .parser import getParser

__all__ = ["Cube", "GenericDimension", "TimeDimension", "CatchAll", "TimeSlot", "getParser", "GpuStore"]


def __getattr__(name):
    if name == "GpuStore":
        from .store import GpuStore

        return GpuStore
    raise AttributeError(name)

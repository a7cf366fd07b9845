"""b200geo — B200-native cell-update engine behind the LibGeoDecomp plugin API.

csrc/          hand-written sm_100a kernels + the C ABI (include/b200geo.h) -> libb200geo.so
capi.py        ctypes binding of that ABI (no fallback)
models.py      cell-model bindings (twin of B200KernelBinding<CELL>)
simulator.py   Initializer / Writer / Steerer / B200Grid / B200Simulator mirror of the reference API
striping.py    slab partition + halo exchange across one-process-per-GPU ranks
synth.py       deterministic synthetic grids for the BASELINE.json configs
"""
__version__ = "0.1"

"""quromorphic.jl_b200 — B200-native state-vector engine beneath Quromorphic.jl's src/Quantum API.

Layout: csrc/ (CUDA kernels + planner + the C ABI of include/qm_b200.h -> libqm_b200.so),
qsim.py / qinfo.py (host mirror of the reference's QSim / QInfo functions for this path),
circuits.py (the synthetic reservoir workloads), julia/B200QSim.jl (the same host shim in Julia).
The directory name contains a dot, so it is imported through __graft_entry__.load_package() under
the module name `quromorphic_jl_b200`.
"""
from . import _lib, circuits, dist, qinfo, qsim, reservoir  # noqa: F401
from ._lib import QMError  # noqa: F401
from .qsim import B200Density, B200State, QSimError  # noqa: F401

"""cm3_b200 - B200-native batched environment stepper for CM3's Checkers and
cooperative-navigation (multi-goal_spread) Markov games.

Product path: Python facades -> ctypes -> libcm3env.so (hand-written sm_100a CUDA kernels).
There is no CPU implementation in this package; without the compiled library or without a
CUDA device the compute entry points raise.
"""
from .build import build_library, library_path  # noqa: F401
from ._lib import Cm3Error, abi_version, device_count, load_library  # noqa: F401
from . import presets  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import cm3_b200` stays cheap
    if name in ("VecCheckers",):
        from .vec_checkers import VecCheckers
        return VecCheckers
    if name in ("VecParticle",):
        from .vec_particle import VecParticle
        return VecParticle
    raise AttributeError(name)

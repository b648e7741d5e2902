"""Import-time shims that let the UNMODIFIED reference (/root/reference) run under
Python 3.12 / NumPy 2.x without gym.  TEST INFRASTRUCTURE ONLY.

Only the golden-vector generator (oracle/gen_golden.py) and the oracle-validation
scripts import this module, and only inside the build container: /root/reference does
not exist on the GPU box, so nothing under tests/ -m gpu, smoke() or bench.py may use it.

Shims:
  1. np.int / np.float aliases            (env/checkers.py:117,279 use the removed names)
  2. stub `gym`, `gym.spaces`, `gym.envs.registration`
                                          (multiagent/__init__.py:1, environment.py:1-5,
                                           multi_discrete.py:6 import them; only
                                           spaces.Discrete(n).n is read on our path)
  3. stub `imp.load_source`               (multiagent/scenarios/__init__.py:1,7)
Nothing under /root/reference is edited or copied.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("CM3_REFERENCE_ROOT", "/root/reference")
_MPE = os.path.join(REFERENCE_ROOT, "env", "multiagent-particle-envs")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "env", "checkers.py"))


def _install_numpy_aliases():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")

    class Env(object):
        pass

    class Space(object):
        pass

    gym.Env = Env
    gym.Space = Space

    spaces = types.ModuleType("gym.spaces")

    class Discrete(Space):
        def __init__(self, n):
            self.n = n

    class Box(Space):
        def __init__(self, low=None, high=None, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Tuple(Space):
        def __init__(self, spaces_):
            self.spaces = spaces_

    spaces.Discrete, spaces.Box, spaces.Tuple = Discrete, Box, Tuple
    gym.spaces = spaces

    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda *a, **k: None

    class EnvSpec(object):
        pass

    registration.EnvSpec = EnvSpec
    envs.registration = registration
    gym.envs = envs

    error = types.ModuleType("gym.error")
    gym.error = error

    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.envs": envs,
        "gym.envs.registration": registration, "gym.error": error,
    })


def _install_imp_stub():
    if "imp" in sys.modules:
        return
    imp = types.ModuleType("imp")

    def load_source(name, pathname):
        spec = importlib.util.spec_from_file_location(name or "_cm3_ref_scenario", pathname)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    imp.load_source = load_source
    sys.modules["imp"] = imp


def load_reference():
    """Returns (checkers_module, MultiAgentEnv, scenarios_module) of the live reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_numpy_aliases()
    _install_gym_stub()
    _install_imp_stub()
    for p in (REFERENCE_ROOT, _MPE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from env import checkers as ref_checkers  # namespace package, env/checkers.py
    from multiagent.environment import MultiAgentEnv
    import multiagent.scenarios as scenarios
    return ref_checkers, MultiAgentEnv, scenarios


def reference_config(name):
    """Load one of the reference's alg/*.json files (container only)."""
    import json
    with open(os.path.join(REFERENCE_ROOT, "alg", name)) as f:
        return json.load(f)

import importlib.util
import os.path as osp


def load(name):
    """scenarios.load("multi-goal_spread.py") - multiagent/scenarios/__init__.py:5-7
    (imp.load_source re-expressed with importlib)."""
    pathname = osp.join(osp.dirname(__file__), name)
    if not osp.isfile(pathname):
        raise FileNotFoundError("scenario %r is not provided by cm3_b200 (only multi-goal_spread.py: "
                                "the reference's other scenarios do not match its own step())" % name)
    spec = importlib.util.spec_from_file_location("cm3_b200_scenario_" + name.replace("-", "_").replace(".py", ""),
                                                  pathname)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

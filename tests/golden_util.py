"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-golden (GPU) tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RESET, STEP = 0, 1

CHECKERS_FIELDS = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "reward",
                   "local_rewards", "done")
PARTICLE_FIELDS = ("global_state", "obs_others", "obs_self", "reward", "reward_n", "done")


def fixtures(prefix):
    return sorted(os.path.basename(p)[:-4]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def checkers_ctor(fix):
    return dict(n_rows=int(fix["ctor_n_rows"]), n_columns=int(fix["ctor_n_columns"]),
                n_obs=int(fix["ctor_n_obs"]), agents_r=[int(x) for x in fix["ctor_agents_r"]],
                agents_c=[int(x) for x in fix["ctor_agents_c"]],
                n_agents=int(fix["ctor_n_agents"]), max_steps=int(fix["ctor_max_steps"]))


def goal_idx_of(goals):
    """np.where(goals[idx]==1)[0][0] (env/checkers.py:235) for one-hot rows."""
    return np.argmax(np.asarray(goals) == 1, axis=-1).astype(np.int32)

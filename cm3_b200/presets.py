"""Environment presets: the env-relevant values of the reference's JSON configs.

These mirror alg/config_checkers_stage{1,2}.json ("init" block), alg/config_particle_*.json and
alg/config.json ("max_steps": 33, "prob_random": 0.2) of 011235813/cm3.  A user's own copy of
those files can be passed to `from_json` unchanged.
"""
import json

MAX_STEPS = 33       # alg/config.json:61
PROB_RANDOM = 0.2    # alg/config.json:62
SEED = 12341         # alg/config.json:6
L_ACTION = 5         # alg/config.json:35, config_checkers_stage2.json:20

CHECKERS = {
    # alg/config_checkers_stage1.json:2-9
    "stage1": dict(n_rows=3, n_columns=8, n_obs=2, agents_r=[0], agents_c=[8], n_agents=1),
    # alg/config_checkers_stage2.json:2-9
    "stage2": dict(n_rows=3, n_columns=8, n_obs=2, agents_r=[0, 2], agents_c=[8, 8], n_agents=2),
}

PARTICLE = {
    # alg/config_particle_stage1.json
    "stage1": dict(n_agents=1, agents_x=[-1.0], agents_y=[-1.0], landmarks_x=[1.0],
                   landmarks_y=[1.0], initial_std=0),
    # alg/config_particle_stage2_antipodal.json
    "antipodal": dict(n_agents=4, agents_x=[-0.9, 0.9, -0.9, 0.9], agents_y=[-0.9, 0.9, 0.9, -0.9],
                      landmarks_x=[0.9, -0.9, 0.9, -0.9], landmarks_y=[0.9, -0.9, -0.9, 0.9],
                      initial_std=0),
    # alg/config_particle_stage2_cross.json
    "cross": dict(n_agents=4, agents_x=[-0.9, 0.9, 0.15, -0.15], agents_y=[-0.15, 0.15, -0.9, 0.9],
                  landmarks_x=[0.9, -0.9, 0.15, -0.15], landmarks_y=[-0.15, 0.15, 0.9, -0.9],
                  initial_std=0),
    # alg/config_particle_stage2_merge.json
    "merge": dict(n_agents=2, agents_x=[-0.9, -0.9], agents_y=[0.2, -0.2], landmarks_x=[0.9, 0.9],
                  landmarks_y=[-0.2, 0.2], initial_std=0.05),
}


def checkers_from_json(path):
    """Reads a config_checkers_stage*.json as the trainers do (train_offpolicy.py:121-127)."""
    with open(path) as f:
        cfg = json.load(f)
    out = dict(cfg["init"])
    out["n_agents"] = cfg["n_agents"]
    return out


def particle_from_json(path):
    """Reads a config_particle_*.json as the trainers do (train_onpolicy.py:113-118)."""
    with open(path) as f:
        return json.load(f)

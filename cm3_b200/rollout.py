"""Vectorised episode loop and transition batches.

The reference's trainers run `while not done:` over ONE env (alg/train_onpolicy.py:302-350,
alg/train_offpolicy.py:309-368): pick actions, `env.step`, append an object-array transition to a
replay buffer, and later turn a sampled batch into per-(time step, agent) rows with
`Alg.process_batch` (alg/alg_credit.py:445-499 for particle, alg/alg_credit_checkers.py:414-477 for
Checkers).  Here the same data flow runs for B envs at once and never leaves HBM:

* `TransitionCollector.collect(T, policy=None)` steps all envs T times (one fused launch when
  the actions are the trainers' pre-training policy, uniform random - train_onpolicy.py:307 - else
  one launch per step with `policy(obs) -> actions`), with in-kernel episode reset, and returns
  the transitions as `[T, B, ...]` device tensors under the reference's variable names.  "next"
  fields are views of the same rollout buffer shifted by one step: nothing is copied.
* `evaluate_episodes(env, policy)` is the reference's evaluation loop (alg/evaluate.py) over the
  batch: one episode per env, reward sums masked at each env's own first `done`.
* `process_batch(...)` re-expresses the reference's batch formatting on the device: one row per
  (time step, env, agent), global quantities repeated per agent, one-hot actions and the
  other-agents' action blocks, in the reference's return order.

Vectorised-env convention (a documented deviation): after a terminal step the
"next" observation is the first observation of the fresh episode; `done` marks the boundary, which
is where the reference's TD targets cut the bootstrap (alg_credit.py `(1 - done)` factors).
"""
import numpy as np
import torch

from .vec_checkers import VecCheckers


class TransitionCollector(object):
    def __init__(self, env, l_action=5, seed=12341):
        self.env, self.l_action, self.seed = env, int(l_action), int(seed)
        self.is_checkers = isinstance(env, VecCheckers)
        self.obs_fields = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v") if self.is_checkers \
            else ("global_state", "obs_others", "obs_self")
        self.t_global = 0
        self._cur = None         # observation of every env before the next step
        self._prev_actions = None
        self._buf, self._buf_T = None, 0

    # ------------------------------------------------------------------ episode control
    def reset(self, **kw):
        out = self.env.reset(**kw)
        self._cur = {k: out[k].clone() for k in self.obs_fields}
        self._prev_actions = torch.zeros(self.env.B, self.env.N, dtype=torch.int8, device=self.env.device)
        return self._cur

    def goals(self):
        """[B, N, l_goal]: one-hot goal per agent for Checkers (train_offpolicy.py:291-298), landmark
        position for particle (train_onpolicy.py:283-285)."""
        env = self.env
        if self.is_checkers:
            bits = (env.state["meta"].to(torch.int64) >> 24).unsqueeze(1) >> torch.arange(env.N, device=env.device)
            idx = (bits & 1).to(torch.int64)
            return torch.nn.functional.one_hot(idx, 2).to(env.dtype)
        return env.state["landmarks"].clone()

    def _buffers(self, T):
        if self._buf is None or self._buf_T != T:
            self._buf = self.env.alloc_outputs(T + 1)
            self._buf_T = T
        return self._buf

    # ------------------------------------------------------------------ collection
    def collect(self, T, policy=None):
        """T steps of every env.  Returns a dict of [T, B, ...] device tensors:
        observation fields, `<field>_next`, `actions`, `actions_prev`, `reward`, `local_rewards`
        (Checkers) / `reward_n` (particle), `done`, `goals`."""
        env = self.env
        if self._cur is None:
            raise RuntimeError("call reset() first")
        T = int(T)
        buf = self._buffers(T)
        for k in self.obs_fields:
            buf[k][0].copy_(self._cur[k])
        tail = {k: v[1:] for k, v in buf.items()}
        redrawn_goals = (not self.is_checkers) and env.prob_random > 0.0
        goals = None
        if policy is None and not redrawn_goals:
            out = env.rollout(T, actions=None, seed=self.seed, t0=self.t_global, auto_reset=True, out=tail,
                              record_actions=True)
            actions = out["actions"]
        else:
            actions = torch.empty(T, env.B, env.N, dtype=torch.int8, device=env.device)
            goals = torch.empty((T,) + tuple(self.goals().shape), dtype=env.dtype, device=env.device)
            for t in range(T):
                obs = {k: buf[k][t] for k in self.obs_fields}
                if policy is None:
                    a = torch.randint(0, self.l_action, (env.B, env.N), device=env.device, dtype=torch.int8)
                else:
                    a = policy(obs).to(torch.int8).reshape(env.B, env.N)
                actions[t] = a
                goals[t] = self.goals()
                env.rollout(1, actions=a.unsqueeze(0), seed=self.seed, t0=self.t_global + t, auto_reset=True,
                            out={k: v[t:t + 1] for k, v in tail.items()})
        self.t_global += T
        done = tail["done"]
        # actions_prev restarts from zeros with every episode (train_offpolicy.py:300)
        prev = torch.empty_like(actions)
        prev[0] = self._prev_actions
        if T > 1:
            prev[1:] = actions[:-1] * (1 - done[:-1].to(torch.int8)).unsqueeze(-1)
        self._prev_actions = actions[-1] * (1 - done[-1].to(torch.int8)).unsqueeze(-1)
        if goals is None:
            goals = self.goals().unsqueeze(0).expand((T,) + tuple(self.goals().shape))
        tr = {"actions": actions, "actions_prev": prev, "done": done, "reward": tail["reward"], "goals": goals}
        tr["local_rewards" if self.is_checkers else "reward_n"] = tail["local_rewards" if self.is_checkers else "reward_n"]
        for k in self.obs_fields:
            tr[k] = buf[k][:T]
            tr[k + "_next"] = buf[k][1:]
        self._cur = {k: buf[k][T].clone() for k in self.obs_fields}
        return tr

    # ------------------------------------------------------------------ batch formatting
    def process_actions(self, actions):
        """alg_credit.py:406-442 for [T, B, N] action indices -> (actions_1hot [T*B*N, l_action],
        actions_others_1hot [T*B*N, N-1, l_action]); row order (t, b, n), n fastest."""
        T, B, N = actions.shape
        a = actions.to(torch.int64).clamp(0, self.l_action - 1)
        onehot = torch.nn.functional.one_hot(a, self.l_action)             # [T, B, N, A]
        if N > 1:
            idx = torch.tensor([[j for j in range(N) if j != n] for n in range(N)], device=actions.device)
            others = onehot[:, :, idx, :]                                   # [T, B, N, N-1, A]
            others = others.reshape(T * B * N, N - 1, self.l_action).to(torch.float64)
        else:
            others = torch.zeros(T * B * N, 0, self.l_action, dtype=torch.float64, device=actions.device)
        return onehot.reshape(T * B * N, self.l_action), others

    def process_batch(self, tr):
        """The reference's Alg.process_batch on the device.  Returns the same tuple, in the same
        order, with the time axis replaced by (time, env): n_steps = T*B and every row is one agent
        of one env at one time step."""
        T, B, N = tr["actions"].shape
        S = T * B
        rep = lambda x: x.reshape((S,) + tuple(x.shape[2:])).repeat_interleave(N, dim=0)   # noqa: E731
        flat = lambda x: x.reshape((S * N,) + tuple(x.shape[3:]))                           # noqa: E731
        a1, ao = self.process_actions(tr["actions"])
        done = rep(tr["done"])
        goals = tr["goals"].reshape((S,) + tuple(tr["goals"].shape[2:]))
        if self.is_checkers:   # alg_credit_checkers.py:414-477
            ap = torch.nn.functional.one_hot(tr["actions_prev"].to(torch.int64).clamp(0, self.l_action - 1),
                                             self.l_action).reshape(S * N, self.l_action)
            return (S, rep(tr["grid"]), tr["vec"].reshape(S, N, 4), flat(tr["obs_others"]), flat(tr["obs_self_t"]),
                    flat(tr["obs_self_v"]), ap, a1, ao, tr["reward"].reshape(S), flat(tr["local_rewards"]),
                    rep(tr["grid_next"]), tr["vec_next"].reshape(S, N, 4), flat(tr["obs_others_next"]),
                    flat(tr["obs_self_t_next"]), flat(tr["obs_self_v_next"]), done, goals)
        # alg_credit.py:445-499
        return (S, tr["global_state"].reshape(S, N, 4), flat(tr["obs_others"]), flat(tr["obs_self"]), a1, ao,
                rep(tr["reward"]), flat(tr["reward_n"]), tr["global_state_next"].reshape(S, N, 4),
                flat(tr["obs_others_next"]), flat(tr["obs_self_next"]), done, goals)


def evaluate_episodes(env, policy, l_action=5, reset_kwargs=None):
    """Vectorised statement of the reference's evaluation loops (alg/evaluate.py:87-123 test_particle,
    :159-203 test_checkers): every one of the B envs plays ONE episode from a fresh reset under
    `policy(obs) -> actions [B, N]` (the caller's greedy actor; `obs` holds the observation fields plus
    `actions_prev` and `goals`), rewards are summed until the env's own first `done`, and the sums are
    averaged over the envs - B takes the place of n_eval.

    Returns (reward_local_mean [N], reward_global_mean, info) as device tensors; info has
    `episode_len` [B] and, like test_checkers' printout, `action_distribution` [N, l_action].
    No host synchronisation inside the loop: all envs are stepped max_steps times (an env that is
    done earlier keeps stepping, as the reference's env would, but is masked out of the sums)."""
    col = TransitionCollector(env, l_action=l_action)
    obs = dict(col.reset(**(reset_kwargs or {})))
    B, N, dev = env.B, env.N, env.device
    goals = col.goals()
    local_key = "local_rewards" if col.is_checkers else "reward_n"
    acc = torch.float64
    reward_local = torch.zeros(B, N, dtype=acc, device=dev)
    reward_global = torch.zeros(B, dtype=acc, device=dev)
    dist_action = torch.zeros(N, l_action, dtype=acc, device=dev)
    episode_len = torch.zeros(B, dtype=torch.int32, device=dev)
    alive = torch.ones(B, dtype=torch.bool, device=dev)
    prev = torch.zeros(B, N, dtype=torch.int8, device=dev)   # evaluate.py:181: actions_prev starts at zero
    for _ in range(int(env.max_steps)):
        obs["actions_prev"], obs["goals"] = prev, goals
        a = policy(obs).to(torch.int8).reshape(B, N)
        out = env.step(a)
        w = alive.to(acc)
        reward_local += out[local_key].to(acc) * w.unsqueeze(-1)
        reward_global += out["reward"].to(acc) * w
        onehot = torch.nn.functional.one_hot(a.to(torch.int64).clamp(0, l_action - 1), l_action).to(acc)
        dist_action += (onehot * w.view(B, 1, 1)).sum(dim=0)
        episode_len += alive.to(torch.int32)
        alive = alive & (out["done"] == 0)
        prev = a
        obs = {k: out[k] for k in col.obs_fields}
    info = {"episode_len": episode_len, "action_distribution": dist_action / dist_action.sum().clamp(min=1)}
    return reward_local.mean(dim=0), reward_global.mean(), info


def split_good_bad(tr, collisions_before, collisions_after):
    """Dual replay buffer split of the reference (replay_buffer_dual.py:13-63,
    train_onpolicy.py:356: an episode is "bad" when scenario.collisions != 0).  Given the per-env
    collision counters around a collected block, returns boolean env masks (good, bad)."""
    bad = (collisions_after - collisions_before) != 0
    return ~bad, bad


def numpy_process_actions(actions, l_action):
    """NumPy statement of the same formatting for ONE env's [time, agents] action array, written
    from the reference's docstring (alg_credit.py:406-421); used by the tests as the checker."""
    actions = np.asarray(actions)
    n_steps, n = actions.shape
    one = np.zeros((n_steps, n, l_action), dtype=int)
    for t in range(n_steps):
        for i in range(n):
            one[t, i, actions[t, i]] = 1
    others = np.zeros((n_steps * n, n - 1, l_action))
    for t in range(n_steps):
        for i in range(n):
            others[t * n + i] = np.stack([one[t, j] for j in range(n) if j != i]) if n > 1 else np.zeros((0, l_action))
    return one.reshape(n_steps * n, l_action), others

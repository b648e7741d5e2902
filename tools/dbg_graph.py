import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from cm3_b200 import VecCheckers, presets
CK2 = dict(presets.CHECKERS["stage2"], max_steps=33)
B, ring = 4096, 35
rng = np.random.default_rng(B)
actions = torch.from_numpy(rng.integers(0, 5, size=(70, B, 2)).astype(np.int8)).cuda()
def make():
    e = VecCheckers(B, **CK2); e.reset(goals=np.eye(2)); return e
fields = ("grid", "vec", "obs_others", "obs_self_t", "obs_self_v", "reward", "local_rewards", "done")
c = make()
slots = c.alloc_outputs(ring, fields=fields)
acts = [actions[t] for t in range(ring)]
outs = [c._outputs_struct({k: v[t] for k, v in slots.items()}) for t in range(ring)]
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    c.step_chained(acts[0], outs[0], seed=5, t0=0, auto_reset=True)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
print("sync after warm-up", c.state["sync"].view(-1)[:8].tolist() if "sync" in c.state else "no sync in state")
c.load_state_dict(make().state_dict())
print("sync after load", c.state["sync"].view(-1)[:8].tolist() if "sync" in c.state else "-")
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
if mode == "graph":
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for t in range(ring):
            c.step_chained(acts[t], outs[t], seed=5, t0=t, auto_reset=True)
    g.replay()
else:
    for t in range(ring):
        c.step_chained(acts[t], outs[t], seed=5, t0=t, auto_reset=True)
torch.cuda.synchronize()
d = make(); ref = d.alloc_outputs(ring, fields=fields)
for t in range(ring):
    d.rollout(1, actions=actions[t:t + 1], auto_reset=True, out={k: v[t:t + 1] for k, v in ref.items()})
torch.cuda.synchronize()
for f in fields:
    bad = (slots[f] != ref[f]).reshape(ring, B, -1).any(-1)
    print(f, "bad slots:", bad.any(1).nonzero().view(-1).tolist()[:12], "bad envs in first bad slot:", int(bad[bad.any(1).nonzero()[0,0]].sum()) if bad.any() else 0)

f = "vec"
bad = (slots[f] != ref[f]).reshape(ring, B, -1).any(-1)
t0 = int(bad.any(1).nonzero()[0, 0]) if bad.any() else -1
if t0 >= 0:
    envs = bad[t0].nonzero().view(-1).tolist()
    print("first bad slot", t0, "bad envs", envs[:40])
    for e in envs[:4]:
        print(" env", e, "tile", e // 16, "got", slots[f][t0, e].tolist(), "want", ref[f][t0, e].tolist(), "prev want", ref[f][t0 - 1, e].tolist(), "next want", ref[f][t0 + 1, e].tolist() if t0 + 1 < ring else None)
        print("   local_rewards got", slots["local_rewards"][t0, e].tolist(), "want", ref["local_rewards"][t0, e].tolist(), " actions", actions[t0, e].tolist())
sy = c._sync.view(-1, 2)
print("sync words: tickets min/max", int(sy[:, 0].min()), int(sy[:, 0].max()), "finished min/max", int(sy[:, 1].min()), int(sy[:, 1].max()), "n tiles", sy.shape[0])
bt = sorted(set(e // 16 for e in (bad[t0].nonzero().view(-1).tolist() if t0 >= 0 else [])))
print("bad tiles", bt[:50])
# within a bad tile: which envs are bad over all slots
if bt:
    tl = bt[0]
    print("tile", tl, "bad env offsets per slot:", [(t, bad[t, tl * 16:(tl + 1) * 16].nonzero().view(-1).tolist()) for t in range(t0, min(t0 + 4, ring))])
    st_c, st_d = c.state_dict(), d.state_dict()
    for k in st_c:
        if torch.is_tensor(st_c[k]) and st_c[k].shape == st_d[k].shape:
            diff = (st_c[k] != st_d[k]).reshape(B, -1).any(-1).nonzero().view(-1).tolist()
            print("final state", k, "differs in envs", diff[:20], "count", len(diff))

"""GPU, >= 2 devices: sharded stepping + rollout all-gather (NCCL and fused peer stores) against a
single-GPU run of the whole batch.  Skipped on 1-GPU boxes; run with `gpurun --gpus 2`."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_sharded_rollout_all_gather_is_shard_invariant():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    g = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(g),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-4000:])
    line = [l for l in res.stdout.splitlines() if l.startswith("MULTIGPU_REPORT ")][-1]
    rep = json.loads(line[len("MULTIGPU_REPORT "):])
    print(rep)
    assert rep["ck2_nccl"] == "ok" and rep["pm2_nccl"] == "ok", rep
    for k in ("ck2_peer", "pm2_peer", "ck2_peer_db", "pm2_peer_db"):
        assert rep[k] == "ok" or rep[k].startswith("unavailable"), rep

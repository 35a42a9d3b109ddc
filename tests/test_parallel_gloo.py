"""world_size-2 gloo test of the N>1 host path: weight-blob broadcast from rank 0, rank-strided game ids
(disjoint and complete), aggregate counters.  Runs on CPU."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import json, os, sys
sys.path.insert(0, os.path.join(%(root)r, "reversi-alpha-zero_b200"))
import numpy as np, torch, torch.distributed as dist
from reversi_zero_b200.agent import model as M
from reversi_zero_b200 import parallel as P
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mc = M.ModelConfig(cnn_filter_num=16, res_layer_num=1, value_fc_size=16)
blob = M.weights_to_blob(mc, M.build_random_weights(mc, 123)) if rank == 0 else None
t = P.broadcast_blob(mc, blob, "cpu")
ids = P.rank_game_ids(rank, world, 1000, slots=3, games_per_slot=2)
tot = P.sum_over_ranks([len(ids), float(rank + 1)], "cpu")
print(json.dumps(dict(rank=rank, digest=M.blob_digest(t.numpy()), ids=ids, tot=tot)))
dist.destroy_process_group()
'''


def test_two_rank_broadcast_and_sharding(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "w.py"
    script.write_text(SCRIPT % dict(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, text=True, env=env))
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0
        outs.append(json.loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o["rank"])
    sys.path.insert(0, os.path.join(ROOT, "reversi-alpha-zero_b200"))
    from reversi_zero_b200.agent import model as M
    mc = M.ModelConfig(cnn_filter_num=16, res_layer_num=1, value_fc_size=16)
    want = M.blob_digest(M.weights_to_blob(mc, M.build_random_weights(mc, 123)))
    assert outs[0]["digest"] == outs[1]["digest"] == want
    a, b = set(outs[0]["ids"]), set(outs[1]["ids"])
    assert not (a & b) and sorted(a | b) == list(range(1000, 1012))
    assert outs[0]["tot"] == outs[1]["tot"] == [12.0, 3.0]

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


WORKER_SCRIPT = r'''
import json, os, sys, time
sys.path.insert(0, os.path.join(%(root)r, "reversi-alpha-zero_b200"))
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch.distributed as dist
from reversi_zero_b200.agent import model as M
from reversi_zero_b200.config import Config
from reversi_zero_b200.worker.self_play import SelfPlayWorker
from test_host_logic import StandInEngine
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
cfg = Config(project_dir=%(tmp)r, data_dir=os.path.join(%(tmp)r, "data%%d" %% rank))
cfg.model.update(dict(cnn_filter_num=16, res_layer_num=1, value_fc_size=16))
cfg.resource.model_dir = os.path.join(%(tmp)r, "model")          # one shared model directory, rank 0 reads it
cfg.resource.model_best_blob_path = os.path.join(cfg.resource.model_dir, "model_best_weight.rzblob.npy")
cfg.resource.next_generation_model_dir = os.path.join(cfg.resource.model_dir, "next_generation")
cfg.resource.create_directories()
cfg.play_data.update(dict(nb_game_in_file=4, enable_ggf_data=False))

class FakeNet:
    blob_floats = M.blob_size(cfg.model)
    digest = "old"
    def load_blob(self, blob):
        self.digest = M.blob_digest(np.asarray(blob))

new_blob = M.weights_to_blob(cfg.model, M.build_random_weights(cfg.model, 77))
if rank == 0:
    np.save(cfg.resource.model_best_blob_path, new_blob)
dist.barrier()
w = SelfPlayWorker(cfg, net=FakeNet(), rank=rank, world_size=world)
w.engine = StandInEngine(per_run=2 if rank == 0 else 5)         # the ranks finish games at different rates
# only rank 0's clock asks for a weight check; rank 1 must follow it into the collective all the same
w.MODEL_CHECK_INTERVAL_SEC = 0 if rank == 0 else 10 ** 9
n = w.start(max_games=12)
runs = sum(1 for c in w.engine.calls if c[0] == "run")
print(json.dumps(dict(rank=rank, n=n, runs=runs, digest=w.net.digest, want=M.blob_digest(new_blob))))
dist.destroy_process_group()
'''


def test_two_rank_worker_control_points(tmp_path):
    """the weight-reload collective and the decision to leave the loop are taken at control points every rank
    reaches the same number of times (all-reduced flags), whatever the ranks' own clocks and harvest rates say."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "w2.py"
    script.write_text(WORKER_SCRIPT % dict(root=ROOT, tmp=str(tmp_path)))
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
    assert outs[0]["runs"] == outs[1]["runs"]                     # same number of control points: no unmatched collective
    assert outs[0]["n"] >= 12 and outs[1]["n"] >= 12 and outs[1]["n"] > outs[0]["n"]   # the fast rank waited for the slow one
    assert outs[0]["digest"] == outs[1]["digest"] == outs[0]["want"]                 # both switched to rank 0's new weights

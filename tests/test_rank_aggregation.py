"""N>1 path of bench.py on CPU: one process per rank, gloo, world_size 2.  The hot path shards as
independent replicas, so the only cross-rank step is the MAX-over-ranks timing and the
SUM of work; this test runs exactly that aggregation."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from tools.rank_agg import aggregate
dist.init_process_group("gloo")
r = dist.get_rank()
out = aggregate(ms_per_step=100.0 + 50.0 * r, solver_s=1.0 + r, work_flops=1e12 * (r + 1), device="cpu")
if r == 0:
    print(json.dumps(out))
dist.destroy_process_group()
"""


def test_gloo_world2_max_time_sum_work():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    import json
    res = json.loads(outs[0][0].strip().splitlines()[-1])
    assert res["ms_per_step"] == 150.0            # max over ranks
    assert res["solver_s"] == 2.0
    assert res["work_flops"] == 3e12              # summed over ranks
    assert abs(res["tflops"] - 3e12 / 2.0 / 1e12) < 1e-12

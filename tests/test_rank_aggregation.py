"""N>1 path of bench.py on CPU: one process per rank, gloo, world_size 2.  The data-path collectives of the sharded
apply are NCCL calls inside the library (covered on GPUs by tools/multi_gpu_check.py and by the
`parity` entry of bench.py); what bench.py itself does across ranks is the MAX-over-ranks timing and the SUM of the
per-rank work -- this test runs exactly that aggregation, plus the contiguous shard ranges the ranks own."""
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


def test_shard_ranges_partition_every_sector():
    """tnl_shard_range: the per-rank slices of a link sector are contiguous, disjoint and cover the sector; the
    remainder rotates with the sector number so that no rank collects all the extra rows."""
    import ctypes as C
    import tennetlib.jl_b200 as T
    lib = T.load(build_if_missing=True)
    for world in (2, 3, 8):
        extra = [0] * world
        for sector, dim in enumerate((1, 7, 64, 935, 1254)):
            got = []
            for r in range(world):
                s, n = C.c_int32(), C.c_int32()
                assert lib.tnl_shard_range(dim, world, sector, r, C.byref(s), C.byref(n)) == 0
                got.append((s.value, n.value))
                extra[r] += n.value - dim // world
            got.sort()
            pos = 0
            for s, n in got:
                assert s == pos and n >= 0
                pos += n
            assert pos == dim
        assert max(extra) - min(extra) <= 2

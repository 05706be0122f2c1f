"""Sharded H_eff apply / Lanczos across ranks vs the CPU oracle.  Launch:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
from oracle import blocksparse as ob, dmrg as od, krylov as ok, models as om

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = T.Context(local)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)

ok_all = True
if os.environ.get("TNL_COMM_BENCH"):
    for kind, name in ((0, "allreduce"), (1, "reduce_scatter"), (2, "allgather")):
        for n in (1, 1024, 1 << 20, 3600000):
            ms = ctx.comm_bench(n, 50, kind)
            if rank == 0:
                print(f"comm {name} n={n} doubles/rank: {ms * 1e3:.1f} us", flush=True)
for kind, N, chi, pos in (("S=1", 8, 40, 4), ("S=1/2", 10, 24, 1), ("S=1", 6, 300, 3)):
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(chi, 1.3, 4, step=2 if kind == "S=1" else 1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(3)))
    od.orthogonalize(mps, pos)
    env_o = od.StateEnvs(mps, H)
    env_o.set_nsite(2); env_o.position(pos)
    phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
    phi_o = phi_o.scale(1 / phi_o.norm())
    Hv_o = env_o.product(phi_o)
    e_o, v_o, info = ok.eigsolve_lanczos(env_o, phi_o)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=pos - 1, rlim=pos + 1)
    env_d.set_nsite(2)
    phi_d = env_d.make_phi(pos); env_d.position(pos)
    phi_d.scale_(1 / phi_d.norm())
    Hv_d = env_d.product(phi_d)
    err = np.abs(Hv_d.to_host().to_dense() - Hv_o.to_dense()).max() / np.abs(Hv_o.to_dense()).max()
    e_d, _ = T.eig_solver(env_d, phi_d)
    good = err < 1e-12 and abs(e_d - e_o) < 1e-10 * abs(e_o)
    ok_all &= good
    print(f"rank {rank}/{world} {kind} N={N} chi={chi} pos={pos}: apply err {err:.2e}  E dev {e_d:.12f} oracle {e_o:.12f} "
          f"local flops {env_d.apply_flops():.0f} {'OK' if good else 'FAIL'}", flush=True)
# full sweeps with the distributed truncation (charge groups spread over the ranks) against the golden fixture
import json
G = json.load(open("tests/golden/oracle_golden.json"))
for g in (G["dmrg"][0], G["dmrg"][2]):
    sites = om.siteinds(g["kind"], g["N"])
    H = om.heisenberg_mpo(sites)
    e, env, sw = T.dmrg2(ctx, om.neel_mps(sites), H, T.DMRGParams(**g["params"]), outputlevel=0)
    de = max(abs(a - b) for a, b in zip(sw.energy, g["energy"]))
    good = sw.maxchi == g["maxchi"] and de < 1e-7 and abs(sw.energy[-1] - g["energy"][-1]) < 1e-10 * abs(g["energy"][-1])
    ok_all &= good
    print(f"rank {rank}/{world} dmrg {g['name']}: E {sw.energy[-1]:.12f} fixture {g['energy'][-1]:.12f} max dE {de:.2e} "
          f"maxchi {sw.maxchi == g['maxchi']} {'OK' if good else 'FAIL'}", flush=True)
# ComplexF64 (planar) vectors: sharded vs replicated on the same data -- real-time exponentiate, then the apply and the
# Lanczos eigensolver on the complex result (plane-by-plane pack / reduce-scatter, two-word scalar all-reduces)
cpos = 7
sites = om.siteinds("S=1", 14)
H = om.heisenberg_mpo(sites)
qn, dm = om.gaussian_link_sectors(200, 1.3, 4, step=2)
mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(5)))
od.orthogonalize(mps, cpos)
res = {}
for mode in (True, False):
    ctx.comm_set_sharding(mode)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=cpos - 1, rlim=cpos + 1)
    env_d.set_nsite(2)
    phi_d = env_d.make_phi(cpos); env_d.position(cpos)
    phi_d.scale_(1 / phi_d.norm())
    ctx.reset_counters()
    _, psi_t = T.exp_solver(env_d, phi_d, -0.2j, solver_krylovdim=12)
    nops = env_d.last_solver_info["numops"]
    Hpsi = env_d.product(psi_t)
    e_c, gs = T.eig_solver(env_d, psi_t.copy())
    res[mode] = (psi_t.to_host().to_dense(), Hpsi.to_host().to_dense(), e_c, nops, env_d.apply_flops())
ctx.comm_set_sharding(True)
(a_t, a_h, a_e, a_n, a_c), (b_t, b_h, b_e, b_n, b_c) = res[True], res[False]
err_t = np.abs(a_t - b_t).max() / np.abs(b_t).max()
err_h = np.abs(a_h - b_h).max() / np.abs(b_h).max()
good = a_c < 0.75 * b_c and err_t < 1e-12 and err_h < 1e-12 and abs(a_e - b_e) < 1e-10 * abs(b_e) and a_n == b_n and np.iscomplexobj(a_t) and np.abs(a_t.imag).max() > 1e-3
ok_all &= good
print(f"rank {rank}/{world} complex sharded vs replicated: exp err {err_t:.2e} apply err {err_h:.2e} E {a_e:.12f} / {b_e:.12f} "
      f"numops {a_n}/{b_n} local apply flops {a_c:.0f} / {b_c:.0f} {'OK' if good else 'FAIL'}", flush=True)
# CouplingModel: the term ids are distributed over the ranks (one all-reduce per apply) -- product, eig_solver and a
# DMRG run against the oracle
from oracle import couplingmodel as oc
sites = om.siteinds("S=1", 8)
M = oc.heisenberg_coupling_model(sites, merge=False, field=0.3, j2=0.5)
qn, dm = om.gaussian_link_sectors(24, 1.3, 4, step=2)
mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(3)))
od.orthogonalize(mps, 4)
env_o = od.StateEnvs(mps, M)
env_o.set_nsite(2); env_o.position(4)
phi_o = ob.contract(env_o.psi[4], env_o.psi[5])
phi_o = phi_o.scale(1 / phi_o.norm())
ref = env_o.product(phi_o).permute(phi_o.inds).to_dense()
e_o, v_o, _ = ok.eigsolve_lanczos(env_o, phi_o)
env_d = T.StateEnvs(ctx, mps.t, M, llim=3, rlim=5)
env_d.set_nsite(2)
phi_d = env_d.make_phi(4); env_d.position(4)
phi_d.scale_(1 / phi_d.norm())
err = np.abs(env_d.product(phi_d).to_host().to_dense() - ref).max() / np.abs(ref).max()
full = None
ctx.comm_set_sharding(False)
env_r = T.StateEnvs(ctx, mps.t, M, llim=3, rlim=5)
env_r.set_nsite(2); env_r.make_phi(4); env_r.position(4)
env_r.product(phi_d)
full = env_r.apply_flops()
ctx.comm_set_sharding(True)
env_d.product(phi_d)
e_d, _ = T.eig_solver(env_d, phi_d)
good = err < 1e-12 and abs(e_d - e_o) < 1e-10 * abs(e_o) and env_d.apply_flops() < 0.8 * full
ok_all &= good
print(f"rank {rank}/{world} CouplingModel ids over ranks: apply err {err:.2e} E dev {e_d:.12f} oracle {e_o:.12f} "
      f"local flops {env_d.apply_flops():.0f} of {full:.0f} {'OK' if good else 'FAIL'}", flush=True)
t = torch.tensor([1.0 if ok_all else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU CHECK", "PASSED" if t.item() == 1.0 else "FAILED", flush=True)
ctx.comm_destroy()
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)

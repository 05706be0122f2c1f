"""Secondary bench workloads: BASELINE.json configs[2], [3], [4] at (or towards) their full size on one B200
(`python bench.py --workload j1j2_cylinder | hubbard_tdvp | ttn_tfi`).  The headline line (configs[1], S=1 chain at
chi = 4096) stays in bench.py; these lines use the same timing rules (warm-up steps, CUDA events on the library stream,
clocks sampled during the timed region) and the same JSON keys, without the CPU arm.

A STEP is one local update through the public API:
  j1j2_cylinder : two-site DMRG bond update (eig_solver)             S=1/2 J1-J2 on a width-6 cylinder, MPO bond dim ~ 30
  hubbard_tdvp  : two-site real-time TDVP bond update (exp_solver forward + one-site backward), ComplexF64,
                  U(1) x U(1) "Electron" sites, krylovdim 30 -- the workload where the Krylov vector kernels matter
  ttn_tfi       : update_position! of a top node of the binary tree (position! + eig_solver), no quantum numbers
"""
from __future__ import annotations

import json
import math
import os
import time

import numpy as np


def _gauss_dims(chi, cands, center, sigma):
    """chi distributed over candidate charges with a Gaussian weight around `center`"""
    w = np.array([math.exp(-sum((a - c) ** 2 / (2.0 * s * s) for a, c, s in zip(q, center, sigma))) for q in cands])
    d = np.floor(w / w.sum() * chi).astype(int)          # sectors whose share is below one state are not opened
    d[int(np.argmax(w))] += chi - int(d.sum())
    return {q: int(x) for q, x in zip(cands, d)}


def _random_state(T, ctx, sites, links, seed, cplx=False):
    psi = []
    for j in range(len(sites)):
        A = T.DeviceTensor.zeros(ctx, [links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)], nrow=2)
        A.fill_random(seed + j)
        if cplx:
            A.promote_()
        psi.append(A)
    return psi


def _links(pm, sites, total_q, chi, sigma):
    """QN link indices with a Gaussian profile around the mean charge of each cut, restricted to reachable charges"""
    N = len(sites)
    nq = len(sites[0].qns[0])
    probe = pm.random_mps_links_q(sites, total_q, lambda j, q: 1 << 30)       # reachable charges and multiplicities
    table = []
    for j in range(N + 1):
        cands = [tuple(q) for q in probe[j].qns]
        center = tuple(total_q[a] * j / N for a in range(nq))
        table.append(_gauss_dims(chi, cands, center, sigma))
    return pm.random_mps_links_q(sites, total_q, lambda j, q: table[j].get(tuple(q), 0))


def _finish(args, ctx, sysenv, name, cfg, steps_ms, apply_flops_total, numops_total, cnt, clocks, extra):
    phases = dict(sysenv.phase_ms) if hasattr(sysenv, "phase_ms") else {}
    solver_s = phases.get("solver", 0.0) * 1e-3
    ms = float(np.mean(steps_ms))
    tf = apply_flops_total / solver_s / 1e12 if solver_s > 0 else 0.0
    line = {"metric": "heff_apply_fp64_tflops", "value": tf, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": cfg.pop("dtype", "f64"), "data": "synthetic", "config": dict(workload=name, **cfg),
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "apply_gflop": apply_flops_total / max(1, numops_total) / 1e9, "applies_per_step": numops_total / args.steps,
            "krylov_vec": {"algorithmic_gb_per_step": cnt["vec_bytes"] / args.steps / 1e9},
            "gpu_launches": int(cnt["launches"]), "clocks": clocks}
    line.update(extra)
    print(json.dumps(line), flush=True)


def run(args, ClockSampler, fp64_peak):
    import torch
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import models as pm
    torch.cuda.set_device(0)
    ctx = T.Context(0)
    if args.workload == "ttn_tfi":
        return run_ttn(args, ctx, T, ClockSampler, fp64_peak)
    cplx = args.workload == "hubbard_tdvp"
    if args.workload == "j1j2_cylinder":
        Lx, Ly = args.nsites // 6 if args.nsites % 6 == 0 else 12, 6
        N = Lx * Ly
        sites = pm.siteinds("S=1/2", N)
        H = pm.heisenberg_bonds_mpo(sites, pm.j1j2_cylinder_bonds(Lx, Ly, 1.0, 0.5))
        links = _links(pm, sites, (0,), args.chi, (2.0,))
        name = "J1-J2 (J2 = 0.5) Heisenberg on a %dx%d cylinder, S=1/2 U(1), two-site DMRG chi=%d (BASELINE.json configs[2])" % (Lx, Ly, args.chi)
        solver, kw = T.eig_solver, dict(maxdim=args.chi, cutoff=args.cutoff, noise=0.0, normalize=True)
    else:
        N = args.nsites if args.nsites % 2 == 0 else 64
        sites = pm.electron_siteinds(N)
        H = pm.hubbard_mpo(sites, pm.ladder_bonds(N // 2, 2), t=1.0, U=4.0)
        links = _links(pm, sites, (N, 0), args.chi, (1.5, 1.5))
        name = "Fermi-Hubbard 2-leg ladder N=%d U(1)xU(1), two-site real-time TDVP chi=%d, ComplexF64 (BASELINE.json configs[3])" % (N, args.chi)
        solver, kw = T.exp_solver, dict(maxdim=args.chi, cutoff=1e-12, normalize=True, time_step=-0.5j * 0.05,
                                        solver_krylovdim=30)
    wmax = max(W.inds[3].dim for W in H[:-1])
    t0 = time.time()
    psi = _random_state(T, ctx, sites, links, 4242, cplx)
    sysenv = T.StateEnvs(ctx, psi, H, llim=0, rlim=N + 1, copy=False)
    del psi
    b0 = max(1, N // 2 - (args.warmup + args.steps) // 2)
    sysenv.orthogonalize(1)
    sysenv.orthogonalize(b0)
    sysenv.set_nsite(2)
    sysenv.position(b0)
    ctx.reserve(int(args.reserve_gb * (1 << 30)))
    ctx.sync()
    setup_s = time.time() - t0
    bond = b0
    for _ in range(args.warmup):
        T.update_position(sysenv, solver, bond, 2, "left", **kw)
        bond += 1
    ctx.sync()
    li = sysenv.site_tensor(bond).inds[0]
    sampler = ClockSampler(0)
    sampler.start()
    ctx.reset_counters()
    sysenv.profile = True
    sysenv.profile_kernels = True
    sysenv.phase_ms, sysenv.phase_log, sysenv.phase_kernel_ms = {}, [], {}
    ctx.profile_gemm(True)
    ctx.profile_read()
    fl, nops, steps_ms, energies = 0.0, 0, [], []
    fl0, nops0 = getattr(sysenv, "solver_flops_total", 0.0), getattr(sysenv, "solver_numops_total", 0)
    for _ in range(args.steps):
        ctx.timer_start(1)
        e, err, eigs = T.update_position(sysenv, solver, bond, 2, "left", **kw)
        steps_ms.append(ctx.timer_stop(1))
        energies.append(e)
        bond += 1
    # every solver call of the update counts (TDVP: two-site forward + one-site backward), not only the last one
    fl, nops = sysenv.solver_flops_total - fl0, sysenv.solver_numops_total - nops0
    clocks = sampler.stop()
    cnt = ctx.counters()
    ctx.profile_read()
    cat = {}
    for ph, d in sysenv.phase_kernel_ms.items():
        for k, v in d.items():
            cat[k] = cat.get(k, 0.0) + v
    prof = dict(sysenv.gemm_prof)
    ctx.profile_gemm(False)
    sysenv.profile = False
    peak, peak_src, burst = fp64_peak()
    gemm_tf = prof["flops"] / (prof["total_ms"] * 1e-3) / 1e12 if prof["total_ms"] > 0 else 0.0
    vec_ms = cat.get("vector", 0.0)
    hbm_peak = 6547.8
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        hbm_peak = float(mp.get("hbm_gbs", hbm_peak))
    except Exception:
        pass
    vec_gbs = cnt["vec_bytes"] / (vec_ms * 1e-3) / 1e9 if vec_ms > 0 else 0.0
    extra = {"energies": energies[-2:], "setup_s": setup_s,
             "timed_bond_sectors": {"bond": bond - args.steps, "nsect": li.nsect, "dim": li.dim, "largest": max(li.dims)},
             "device_ms_per_step_by_kernel_class": {k: v / args.steps for k, v in cat.items()},
             "roofline": {"bound": "tensor", "kernel": "tnl::gemm_tma_ws_kernel (FP64 DMMA grouped GEMM)", "achieved": gemm_tf,
                          "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak, "traffic": None, "peak_source": peak_src,
                          "launches": prof["launches"]},
             "roofline_krylov_vectors": {"bound": "hbm", "kernel": "Krylov vector kernels (dot / axpy / MGS / lincomb)",
                                         "achieved": vec_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": vec_gbs / hbm_peak,
                                         "device_ms_per_step": vec_ms / args.steps}}
    cfg = dict(nsites=N, chi=args.chi, mpo_bond_dim=wmax, dtype="c128" if cplx else "f64",
               step="one two-site bond update at a mid-chain bond through update_position",
               l2_policy="inputs larger than L2")
    _finish(args, ctx, sysenv, name, cfg, steps_ms, fl, nops, cnt, clocks, extra)


def run_ttn(args, ctx, T, ClockSampler, fp64_peak):
    from tennetlib.jl_b200 import ttn as dt
    N, chi = (args.nsites if args.nsites != 100 else 128), args.chi
    sites = dt.dense_siteinds(N)
    M = dt.tfi_coupling_model(sites, h=1.0)
    t0 = time.time()
    psi0 = dt.default_random_ttn(ctx, sites, chi, seed=1)
    sysenv = dt.StateEnvsTTN(psi0, M)
    path = dt.default_sweeppath(sysenv.psi)
    top = [n for n in path if n[0] == max(p[0] for p in path)]
    second = [n for n in path if n[0] == max(p[0] for p in path) - 1]
    visit = (top + second) * (1 + (args.warmup + args.steps) // max(1, len(top + second)))
    setup_s = time.time() - t0
    kw = dict(maxdim=chi, cutoff=-1.0, normalize=True)
    k = 0
    for _ in range(args.warmup):
        dt.update_position(sysenv, dt.eig_solver, visit[k], **kw)
        k += 1
    ctx.sync()
    sampler = ClockSampler(0)
    sampler.start()
    ctx.reset_counters()
    ctx.profile_gemm(True)
    ctx.profile_read()
    steps_ms, solver_ms, fl, nops, energies = [], 0.0, 0.0, 0, []
    for _ in range(args.steps):
        node = visit[k]
        k += 1
        ctx.timer_start(1)
        sysenv.position(node, **kw)
        ctx.timer_start(2)
        e, phi = dt.eig_solver(sysenv, sysenv.psi[node])
        solver_ms += ctx.timer_stop(2)
        sysenv.psi[node] = phi
        steps_ms.append(ctx.timer_stop(1))
        energies.append(e)
        info = sysenv.last_solver_info
        nops += info["numops"]
        fl += info["apply_flops"] * info["numops"]
    clocks = sampler.stop()
    cnt = ctx.counters()
    pr = ctx.profile_read()
    ctx.profile_gemm(False)
    peak, peak_src, burst = fp64_peak()
    gemm_tf = pr["flops"] / (pr["total_ms"] * 1e-3) / 1e12 if pr["total_ms"] > 0 else 0.0
    tf = fl / (solver_ms * 1e-3) / 1e12 if solver_ms > 0 else 0.0
    line = {"metric": "heff_apply_fp64_tflops", "value": tf, "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean(steps_ms)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "TTN binary-tree ground state, transverse-field Ising N=%d maxdim=%d (BASELINE.json "
                                   "configs[4])" % (N, chi), "nsites": N, "chi": chi,
                       "step": "update_position! (position! + eig_solver krylovdim 5) of a node of the two top layers",
                       "nodes": len(path), "l2_policy": "node tensors larger than L2 (chi^3 doubles)"},
            "energies": energies[-2:], "solver_ms_per_step": solver_ms / args.steps,
            "apply_gflop": fl / max(1, nops) / 1e9, "applies_per_step": nops / args.steps, "setup_s": setup_s,
            "device_ms_per_step_by_kernel_class": {k2: v / args.steps for k2, v in pr["category_ms"].items()},
            "roofline": {"bound": "tensor", "kernel": "tnl::gemm_tma_ws_kernel (FP64 DMMA grouped GEMM)", "achieved": gemm_tf,
                         "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak, "traffic": None, "peak_source": peak_src,
                         "launches": pr["launches"]},
            "gpu_launches": int(cnt["launches"]), "clocks": clocks}
    print(json.dumps(line), flush=True)

#!/usr/bin/env python
"""bench.py -- DMRG hot path at chi=4096 on B200 (BASELINE.json metric: "DMRG sweep time (s) and H_eff-apply
FP64 TFLOP/s at chi=4096").

Workload (config[1] of BASELINE.json): S=1 Heisenberg chain, N=100, U(1) QN, two-site DMRG, maxdim 4096.
The MPS is a random QN MPS with the Gaussian link-sector profile of SURVEY.md section 8d (sigma = 1.3 in Sz),
generated on the device, right-canonicalised with QR, centre moved to mid-chain, all environments built.

A STEP is one pass of the hot path = one two-site bond update through the public API
(`update_position(sysenv, eig_solver, bond, 2, "left")`): phi = A_j A_{j+1}, environment update, Lanczos
eig_solver (<= 7 H_eff applies + Krylov vector ops), noise-free truncation (replacebond!).

  value / metric : H_eff-apply FP64 TFLOP/s = algorithmic apply flops (reference contraction order, 2mnk per
                   block GEMM) / device time of the eig_solver phase (applies + Krylov vector kernels)
  ms_per_step    : one full bond update;   sweep_time_s_est = 198 * ms_per_step (upper bound: edge bonds are cheaper)
  sweep_time_s   : MEASURED: one real `fullsweep!` (198 bond updates, left to right and back) of the same state, timed
                   after the steps (CUDA events on the library stream)
  parity         : N > 1 only, before the timed region: one H_eff v and one eig_solver computed sharded and unsharded
                   (replicated, no collectives) on the same vector; relative deviations
  truncation_decaying_spectrum : replacebond! of a two-site tensor whose Schmidt values decay over 24 decades (a
                   converged state, not the flat spectrum of the random MPS), per SVD driver
  e2e            : same metric through the C ABI with HOST buffers: host phi -> tnl_tensor_import (H2D) ->
                   tnl_eigsolve_lanczos -> tnl_tensor_export (D2H), copies inside the timed region
  roofline       : grouped DGEMM kernel (dominant), achieved algorithmic TFLOP/s per launch (CUDA events on the
                   library stream) vs the cuBLAS DGEMM 8192^3 rate measured in this very run (torch.matmul fp64)
  cpu_baseline   : oracle (NumPy restatement of the ITensors CPU path) timed on the host cores for one H_eff
                   apply at the same bond ("port")
`--impl reference` times that oracle apply alone on synthetic tensors of the same sector structure.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
# (CUDA_MODULE_LOADING=EAGER would remove the one-off lazy-loading hiccups that `phase_ms_max_over_steps` exposes, but
# on a cold box it costs minutes of start-up: left to the caller's environment.)

FP64_PEAK_TFLOPS_FALLBACK = 35.46     # cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_peaks_fp64.json)
SIGMA_SZ = 1.3
QMAX = 6


def fp64_peak(measure_s: float = 1.5):
    """cuBLAS DGEMM 8192^3 on this GPU, in this run: best single call (burst) and the rate of a back-to-back loop of
    `measure_s` seconds (sustained, the denominator for kernels timed inside a long step).  MEASURED_PEAKS.json has
    no FP64 entry, so the bench measures its own denominator instead of reading a committed number."""
    try:
        import torch
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device="cuda")
        b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty_like(a)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        reps = max(3, int(measure_s * best * 1e12 / (2.0 * n ** 3)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        sus = reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a, b, c
        torch.cuda.empty_cache()
        return sus, "cuBLAS DGEMM 8192^3 measured in this run (torch.matmul fp64): sustained %.2f / burst %.2f TFLOP/s" % (sus, best), best
    except Exception as exc:                                   # keep the line alive
        return FP64_PEAK_TFLOPS_FALLBACK, "fallback %.2f (in-run cuBLAS measurement failed: %r)" % (FP64_PEAK_TFLOPS_FALLBACK, exc), None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        """Start of the timed region: only samples from here on count.  The process itself is started during the
        warm-up -- the first nvidia-smi query initialises NVML and can hold the driver for hundreds of milliseconds,
        which must not land inside the first timed step."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def hbm_rooflines(cnt, prof, steps):
    """HBM-bound kernel classes of the step (this rank): algorithmic bytes (read + written once)
    over the CUDA-event time of the class, against the measured copy bandwidth of MEASURED_PEAKS.json (driver-written;
    fallback: 6548 GB/s measured on this pool in round 1, profiles/r01_peaks_fp64.json)."""
    peak, src = 6547.8, "profiles/r01_peaks_fp64.json (copy, measured on this pool)"
    try:
        mp = json.load(open(os.path.join(HERE, "MEASURED_PEAKS.json")))
        peak, src = float(mp["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    out = {"peak": peak, "unit": "GB/s", "peak_source": src}
    for name, key, cat in (("transform", "transform_bytes", "transform"), ("krylov_vectors", "vec_bytes", "vector")):
        ms = prof["category_ms"].get(cat, 0.0)
        gbs = cnt[key] / 1e9 / (ms * 1e-3) if ms > 0 else 0.0
        out[name] = {"bound": "hbm", "achieved": gbs, "frac": gbs / peak, "algorithmic_gb_per_step": cnt[key] / steps / 1e9,
                     "device_ms_per_step": ms / steps}
    return out


# ------------------------------------------------------------------------------------------------
def workload_config(args):
    return {"workload": "S=1 Heisenberg chain N=%d U(1) QN two-site DMRG chi=%d (BASELINE.json configs[1])" % (args.nsites, args.chi),
            "nsites": args.nsites, "chi": args.chi, "site": "S=1", "mpo_bond_dim": 5,
            "link_sector_profile": "gaussian sigma=%.1f (Sz units), |Sz|<=%d" % (SIGMA_SZ, QMAX),
            "step": "one two-site bond update at a mid-chain bond (phi, env update, Lanczos eig_solver "
                    "krylovdim=5 maxiter=2 tol=1e-14, truncation maxdim=chi cutoff=%g via %s)" % (args.cutoff, args.decomp),
            "l2_policy": "inputs larger than L2 (per-apply working set > 3 GB at chi=4096)",
            "cutoff": args.cutoff, "decomp": args.decomp, "svd_alg": args.svd_alg}


def oracle_apply_inputs(sector_qns, sector_dims, rng):
    """Synthetic mid-chain bond for the CPU arm: phi, L, R with every symmetry-allowed block N(0,1)
     and the exact S=1 Heisenberg MPO tensors."""
    from oracle import blocksparse as ob, models as om
    sites = om.siteinds("S=1", 4)
    H = om.heisenberg_mpo(sites)
    W1, W2 = H[1], H[2]
    link = ob.Index(sector_qns, sector_dims, dir=+1, tags="Link")
    l, r = link.sim(), link.sim()
    s1, s2 = W1.inds[2], W2.inds[2]
    phi = ob.BSTensor.random([l.copy(dir=+1), s1.copy(dir=+1), s2.copy(dir=+1), r.copy(dir=-1)], rng)
    phi = phi.scale(1.0 / phi.norm())
    wl, wr = W1.inds[0], W2.inds[3]
    L = ob.BSTensor.random([l.prime().copy(dir=+1), wl.copy(dir=-1), l.copy(dir=-1)], rng)
    R = ob.BSTensor.random([r.prime().copy(dir=-1), wr.copy(dir=+1), r.copy(dir=+1)], rng)
    return phi, L, W1, W2, R


def oracle_apply(phi, L, W1, W2, R):
    from oracle.blocksparse import contract
    return contract(contract(contract(contract(phi, L), W1), W2), R).noprime()


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class all_host_threads:
    """Give BLAS every host core for the CPU arm.  Under `torch.distributed.run` OMP_NUM_THREADS=1 is exported to
    the ranks, which would leave the reference arm on one BLAS thread (N>1 ratios inflated 4.6x)."""

    def __enter__(self):
        self.ctl = None
        try:
            from threadpoolctl import threadpool_limits
            self.ctl = threadpool_limits(limits=host_cores())
        except Exception:
            pass
        return self

    def __exit__(self, *exc):
        if self.ctl is not None:
            self.ctl.restore_original_limits()
        return False


def bench_sectors(chi: int):
    """Left-link sectors of the first timed bond of the GPU arm (recorded from a GPU run of this workload, committed
    as profiles/r02_bench_sectors.json) so that both arms time the same block structure; the Gaussian profile the
    state is initialised with when no record exists for this chi."""
    try:
        rec = json.load(open(os.path.join(HERE, "profiles", "r02_bench_sectors.json")))
        ent = rec.get(str(chi))
        if ent:
            return [tuple(q) for q in ent["qns"]], list(ent["dims"]), "profiles/r02_bench_sectors.json (timed bond of the GPU arm)"
    except Exception:
        pass
    from tennetlib.jl_b200 import models as pm          # host-only helper (sector profile); no CUDA involved
    qns, dims = pm.gaussian_link_sectors(chi, SIGMA_SZ, QMAX, 0, 2)
    return qns, dims, "gaussian profile of the initial state"


def time_oracle_apply(sector_qns, sector_dims, reps: int, seed: int = 20262):
    from oracle import blocksparse as ob
    rng = np.random.default_rng(seed)
    t0 = time.time()
    phi, L, W1, W2, R = oracle_apply_inputs(sector_qns, sector_dims, rng)
    gen_s = time.time() - t0
    times, flops = [], 0
    with all_host_threads():
        for _ in range(reps):
            ob.reset_flops()
            t0 = time.time()
            oracle_apply(phi, L, W1, W2, R)
            times.append(time.time() - t0)
            flops = ob.get_flops()
        threads = blas_threads()
    return times, flops, gen_s, threads


def run_reference(args):
    """CPU arm: the oracle port of the reference CPU path (Julia is not installed)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qns, dims, src = bench_sectors(args.chi)
    reps = args.warmup + args.steps
    times, flops, gen_s, cores = time_oracle_apply(qns, dims, reps)
    timed = times[args.warmup:]
    tf = flops / np.mean(timed) / 1e12
    line = {"impl": "reference", "metric": "heff_apply_fp64_tflops", "value": tf, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(timed)),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "host_cores": host_cores(), "kind": "port",
                             "sample": "one H_eff apply per step (reference contraction order, NumPy/BLAS per block pair) "
                                       "on a synthetic mid-chain bond; link sectors: " + src},
            "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "apply_gflop": flops / 1e9, "note": "oracle port of the ITensors CPU path, not Julia (reference cannot run here)"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import models as pm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = T.Context(local)
    sharded = world > 1 and args.multi == "sharded"
    if sharded:
        # library-level NCCL communicator: rank 0 creates the id, torch.distributed only carries the 128 bytes
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    # untimed pre-warm: a tiny DMRG touches every kernel and cuSOLVER path once (first process on a fresh box
    # otherwise pays library page-in inside the first steps)
    ws = pm.siteinds("S=1", 8)
    T.dmrg2(ctx, pm.neel_mps(ws), pm.heisenberg_mpo(ws), T.DMRGParams(nsweeps=[2], maxdim=[30], cutoff=1e-14, noise=[1e-3]),
            outputlevel=0)
    N, chi = args.nsites, args.chi
    sites = pm.siteinds("S=1", N)
    H = pm.heisenberg_mpo(sites)
    qns, dims = pm.gaussian_link_sectors(chi, SIGMA_SZ, QMAX, 0, 2)
    links = pm.random_mps_links(sites, qns, dims, 0)
    t_setup = time.time()
    psi = []
    for j in range(N):
        A = T.DeviceTensor.zeros(ctx, [links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)], nrow=2)
        A.fill_random(20262 + (0 if sharded else 1000 * rank) + j)   # sharded: identical replicated state
        psi.append(A)
    sysenv = T.StateEnvs(ctx, psi, H, llim=0, rlim=N + 1, copy=False)
    del psi
    b0 = max(1, N // 2 - (args.warmup + args.steps) // 2)
    sysenv.orthogonalize(1)
    sysenv.orthogonalize(b0)
    sysenv.set_nsite(2)
    sysenv.position(b0)
    ctx.reserve(int(args.reserve_gb * (1 << 30)))      # pool head-room for the per-bond allocations of the sweep
    ctx.sync()
    t_setup = time.time() - t_setup

    kw = dict(maxdim=chi, cutoff=args.cutoff, noise=0.0, normalize=True, svd_alg=args.svd_alg,
              which_decomp=None if args.decomp == "auto" else args.decomp)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- parity of the sharded path, inside the run that produces the numbers:
    # H_eff v and one eig_solver on the same vector, sharded (reduce-scatter / all-gather) vs unsharded (every rank
    # computes the whole replicated problem, no collectives)
    parity = None
    if sharded:
        phi0 = sysenv.make_phi(b0)
        sysenv.position(b0)
        phi0.scale_(1.0 / phi0.norm())
        hv_s = sysenv.product(phi0)
        v_s = phi0.copy()
        e_s, v_s = T.eig_solver(sysenv, v_s)
        ctx.comm_set_sharding(False)
        sysenv.position(b0)
        hv_u = sysenv.product(phi0)
        v_u = phi0.copy()
        e_u, v_u = T.eig_solver(sysenv, v_u)
        ctx.comm_set_sharding(True)
        den = hv_u.norm()
        hv_s.axpy_(hv_u, -1.0)
        ov = abs(v_s.dot(v_u))
        parity = {"apply_rel_err": hv_s.norm() / den, "dE_rel": abs(e_s - e_u) / abs(e_u), "E_sharded": e_s,
                  "E_unsharded": e_u, "ritz_vector_overlap": ov, "bond": b0}
        del phi0, hv_s, hv_u, v_s, v_u

    energies = []
    bond = b0
    sampler = ClockSampler(local)
    sampler.start()                  # NVML start-up happens during the warm-up; samples count from mark() on
    for _ in range(args.warmup):
        e, err, eigs = T.update_position(sysenv, T.eig_solver, bond, 2, "left", **kw)
        bond += 1
    barrier()
    li0 = sysenv.site_tensor(bond).inds[0]
    timed_sectors = {"bond": bond, "qns": [list(q) for q in li0.qns], "dims": list(li0.dims)}
    sampler.mark()
    ctx.reset_counters()
    sysenv.profile = True
    sysenv.profile_kernels = True
    sysenv.phase_ms = {}
    sysenv.phase_log = []
    sysenv.phase_kernel_ms = {}
    sysenv.collective_detail = {}
    ctx.profile_gemm(True)
    ctx.profile_read()
    apply_flops_total, numops_total = 0.0, 0
    if args.profile_region:
        torch.cuda.profiler.start()     # ncu --profile-from-start off: only the timed region is captured
    ctx.timer_start(1)              # CUDA events on the library's stream (slot 0 is used by the phase timers)
    for _ in range(args.steps):
        e, err, eigs = T.update_position(sysenv, T.eig_solver, bond, 2, "left", **kw)
        energies.append(e)
        numops_total += sysenv.last_solver_info["numops"]
        apply_flops_total += sysenv.last_solver_info["apply_flops"] * sysenv.last_solver_info["numops"]
        bond += 1
    total_s = ctx.timer_stop(1) * 1e-3
    if args.profile_region:
        torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    cnt = ctx.counters()
    ctx.profile_read()
    prof = dict(sysenv.gemm_prof)
    prof["category_ms"] = {}
    for ph, d in sysenv.phase_kernel_ms.items():
        for k, v in d.items():
            prof["category_ms"][k] = prof["category_ms"].get(k, 0.0) + v
    ctx.profile_gemm(False)
    sysenv.profile = False
    sysenv.profile_kernels = False
    phases = dict(sysenv.phase_ms)

    sysenv.set_nsite(2)
    bond_e2e = bond
    li = sysenv.site_tensor(bond_e2e).inds[0]          # left link of the e2e / cpu_baseline bond (before later legs)
    phi = sysenv.make_phi(bond_e2e)
    sysenv.position(bond_e2e)
    _ = sysenv.product(phi)
    flops_per_apply = sysenv.apply_flops()
    solver_s = phases.get("solver", 0.0) * 1e-3
    tf_apply = apply_flops_total / solver_s / 1e12 if solver_s > 0 else 0.0

    # ---- e2e through the C ABI with HOST buffers: what the Julia shim does around one eig_solver call.  The host side
    # holds the two-site tensor as one flat NDTensors vector (ITensor storage already is that vector: no host-side
    # repacking belongs to the boundary); it lives in pinned memory.  Timed per call: tnl_tensor_import (H2D copy +
    # relayout into the charge-fused layout) -> tnl_eigsolve_lanczos -> tnl_tensor_export (relayout + D2H copy).
    import ctypes as C
    from tennetlib.jl_b200._lib import check as _check
    from tennetlib.jl_b200.tensor import _index_array, flatten_blocks
    phi_host = phi.to_host()
    coords, offsets, flat, nb = flatten_blocks(phi_host)
    nbytes = int(flat.nbytes)
    rank_phi = len(phi_host.inds)
    arr, nq, _keep = _index_array(phi_host.inds)
    pin_in = torch.empty(flat.size, dtype=torch.float64).pin_memory()
    pin_in.numpy()[:] = flat
    pin_out = torch.empty(flat.size, dtype=torch.float64).pin_memory()
    out_coords = np.zeros((max(nb, 1), rank_phi), dtype=np.int32)
    out_offsets = np.zeros(max(nb, 1), dtype=np.int64)

    def e2e_call():
        h = C.c_void_p()
        _check(ctx.lib.tnl_tensor_import(ctx.h, rank_phi, nq, arr, nb, coords.ctypes.data, offsets.ctypes.data,
                                         pin_in.data_ptr(), 1, C.byref(h)), ctx.h)
        d = T.DeviceTensor(ctx, h, phi_host.inds)
        T.eig_solver(sysenv, d)
        _check(ctx.lib.tnl_tensor_export(d.h, out_coords.ctypes.data, out_offsets.ctypes.data, pin_out.data_ptr()), ctx.h)
        return sysenv.last_solver_info["numops"]

    def e2e_call_dict():
        """same boundary through the Python HostTensor (dict of blocks): adds host-side repacking on both sides"""
        d = T.DeviceTensor.from_host(ctx, phi_host, nrow=1)
        T.eig_solver(sysenv, d)
        d.to_host()
        return sysenv.last_solver_info["numops"]

    e2e_ops = 0
    e2e_path = "tnl_tensor_import -> tnl_eigsolve_lanczos -> tnl_tensor_export (flat host vectors in pinned memory)"
    try:
        e2e_call()
        e2e_norm = float(np.linalg.norm(pin_out.numpy()))      # the Ritz vector comes back normalised
        if not (0.99 < e2e_norm < 1.01):
            raise RuntimeError("flat e2e path returned a vector of norm %g" % e2e_norm)
    except Exception as exc:                                    # keep the bench line alive: measured through HostTensor
        sys.stderr.write("e2e flat path failed (%r); using the HostTensor path\n" % (exc,))
        e2e_call, e2e_norm = e2e_call_dict, float("nan")
        e2e_path = "HostTensor dict -> tnl_tensor_import -> tnl_eigsolve_lanczos -> tnl_tensor_export -> HostTensor dict"
        e2e_call()
    barrier()
    t1 = time.time()
    for _ in range(max(1, min(args.steps, 3))):
        e2e_ops += e2e_call()
    ctx.sync()
    e2e_s = time.time() - t1
    tf_e2e = flops_per_apply * e2e_ops / e2e_s / 1e12

    # ---- truncation of a DECAYING spectrum (the random MPS has a flat spectrum, so the Gram
    # driver's guard never trips in the timed steps).  phi' = Q diag(w) V with Q a left isometry (QR of the centre),
    # V the right-orthonormal neighbour and w_k^2 = 10^(-24 k / chi) pooled over the link sectors: its Schmidt values
    # are exactly w.  Timed: replacebond! alone, per SVD driver.
    trunc_decay = None
    if args.decaying:
        b = bond_e2e
        # gauge moves there and back make the link between b and b+1 rank-consistent (a link of the random initial
        # MPS may carry sectors larger than either side can support); V = right isometry, Q = left isometry
        sysenv.orthogonalize(b + 1)
        sysenv.orthogonalize(b)
        Vt = sysenv.site_tensor(b + 1).copy()
        sysenv.orthogonalize(b + 1)                        # psi[b] = Q (left isometry), centre on b+1
        m = sysenv.site_tensor(b).inds[2]
        nlink = sum(m.dims)
        ranks = np.random.default_rng(7).permutation(nlink)
        w, at = [], 0
        for dsec in m.dims:
            r = np.sort(ranks[at:at + dsec]); at += dsec
            w.append(np.sqrt(10.0 ** (-24.0 * r / nlink)))
        w = np.concatenate(w)
        w /= np.linalg.norm(w)
        trunc_decay = {"decades": 24, "link_dim": int(nlink), "cutoff": args.cutoff, "maxdim": chi, "ms": {}, "kept": {},
                       "truncerr": {}}
        Qt = sysenv.site_tensor(b).copy()
        for alg in ("divide_and_conquer", "polar"):
            ms_l = []
            for rep in range(2):
                Vw = Vt.copy().scale_index_(0, w)
                sysenv.set_site_tensor(b, Qt.copy())
                sysenv.set_site_tensor(b + 1, Vw)
                sysenv.llim, sysenv.rlim = b, b + 2
                sysenv.set_nsite(2)
                phid = sysenv.make_phi(b)
                sysenv.position(b)
                barrier()
                ctx.timer_start(2)
                terr_d, eigs_d = sysenv.replacebond(b, phid, maxdim=chi, mindim=1, cutoff=args.cutoff, noise=0.0,
                                                    ortho="left", normalize=True, svd_alg=alg, which_decomp="svd")
                ms_l.append(ctx.timer_stop(2))
            trunc_decay["ms"][alg] = ms_l[-1]
            trunc_decay["kept"][alg] = int(len(eigs_d))
            trunc_decay["truncerr"][alg] = float(terr_d)
        # put the original (normalised) state back: centre on b+1
        sysenv.set_site_tensor(b, Qt)
        sysenv.set_site_tensor(b + 1, Vt)
        sysenv.llim, sysenv.rlim = b, b + 2
        del Vt, Qt

    # ---- one MEASURED full sweep (the metric's "DMRG sweep time"): fullsweep! = 198 bond updates from site 1 to N
    # and back, same kwargs as the steps; CUDA events on the library stream + host wall clock
    sweep = None
    if args.sweep:
        sysenv.orthogonalize(1)
        swd = T.SweepData()
        barrier()
        t_w = time.time()
        ctx.timer_start(2)
        T.fullsweep(sysenv, T.eig_solver, 2, swd, outputlevel=0, **kw)
        sweep_ms = ctx.timer_stop(2)
        barrier()
        sweep = {"sweep_time_s": sweep_ms * 1e-3, "wall_s": time.time() - t_w, "bond_updates": 2 * (N - 1),
                 "energy": swd.energy[-1], "maxlinkdim": swd.maxchi[-1], "maxtruncerr": swd.maxtruncerr[-1]}
        if world > 1:
            tt = torch.tensor([sweep["sweep_time_s"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sweep["sweep_time_s"] = tt.item()

    # ---- aggregate over ranks (sharded: every rank holds a slice of the same step; replicas: independent copies):
    # time = max over ranks, work = sum over ranks
    from tools.rank_agg import aggregate
    ms_per_step = 1e3 * total_s / args.steps
    agg = aggregate(ms_per_step, solver_s, apply_flops_total)
    ms_per_step, tf_apply = agg["ms_per_step"], agg["tflops"]
    tf_e2e = aggregate(0.0, e2e_s, flops_per_apply * e2e_ops)["tflops"]

    # per-rank device time by kernel class: a slow rank shows up as waiting time inside the collectives of the others
    per_rank = None
    if world > 1:
        names = ("gemm", "transform", "vector", "collective")
        mine = torch.tensor([prof["category_ms"].get(k, 0.0) / args.steps for k in names], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{k: float(v) for k, v in zip(names, t.tolist())} for t in allr]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src, peak_burst = fp64_peak()
    gemm_tf = prof["flops"] / (prof["total_ms"] * 1e-3) / 1e12 if prof["total_ms"] > 0 else 0.0
    traffic, traffic_src = None, None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tj = json.load(open(os.path.join(HERE, "profiles", "gemm_traffic.json")))
        traffic, traffic_src = float(tj["dram_bytes_per_launch_mean"]), tj["source"]
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "tnl::gemm_tma_ws_kernel (FP64 DMMA grouped GEMM: TMA producer warpgroup + 8 DMMA warps, 128x128x16 tiles, 6 stages; 64x64 cp.async kernel for small sectors)",
                "achieved": gemm_tf, "peak": peak, "unit": "TFLOP/s", "frac": gemm_tf / peak, "traffic": traffic,
                "traffic_unit": "bytes per launch (DRAM read + write)", "traffic_source": traffic_src,
                "peak_source": peak_src, "peak_burst": peak_burst, "launches": prof["launches"],
                "avg_launch_ms": prof["total_ms"] / max(1, prof["launches"]),
                "algorithmic_gflop_per_launch": prof["flops"] / max(1, prof["launches"]) / 1e9,
                "best_launch_tflops": prof["max_tflops"]}

    cpu = None
    if not args.no_cpu_baseline:
        # bounded sample: ONE oracle apply on a synthetic bond with the same sector profile
        times, fl, gen_s, cores = time_oracle_apply(list(li.qns), list(li.dims), 1)
        cpu = {"value": fl / times[0] / 1e12, "unit": "TFLOP/s", "cores": cores, "host_cores": host_cores(), "kind": "port",
               "sample": "one H_eff apply (%.1f GFLOP algorithmic) by the NumPy oracle on a synthetic bond with the "
                         "left-link sectors of the timed bond; %.1f s" % (fl / 1e9, times[0])}

    line = {"metric": "heff_apply_fp64_tflops", "value": tf_apply, "unit": "TFLOP/s", "n_gpus": world,
            "energies": energies[-2:], "parity": parity,
            "sweep_time_s": sweep["sweep_time_s"] if sweep else None,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            # one sweep of one fixed state at every N: total work does not grow with N
            "scaling": "weak" if (world > 1 and not sharded) else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": dict(workload_config(args),
                           parallelism=("H_eff apply + Krylov vectors sharded over the right link across %d GPUs "
                                        "(T3 R^T GEMM fused with the reduce-scatter over peer memory -- NCCL "
                                        "reduce-scatter if peer memory is unavailable --, scalar all-reduces, one "
                                        "all-gather per eig_solver); truncation groups distributed; env update sharded "
                                        "over a link index + all-gather" % world)
                           if sharded else ("replicas x%d" % world if world > 1 else "single GPU"),
                           first_timed_bond=b0 + args.warmup),
            "sweep": sweep, "truncation_decaying_spectrum": trunc_decay, "timed_bond_sectors": timed_sectors,
            "allreduce_gb_per_step": cnt["_"] / args.steps / 1e9,
            "sweep_time_s_est": 198 * ms_per_step * 1e-3,
            "phase_ms_per_step": {k: v / args.steps for k, v in phases.items()},
            "phase_ms_max_over_steps": {k: max(v for n, v in sysenv.phase_log if n == k) for k in phases},
            "step_ms_list": [round(sum(v for _, v in sysenv.phase_log[4 * i:4 * i + 4]), 2) for i in range(len(sysenv.phase_log) // 4)],
            "apply_gflop": apply_flops_total / max(1, numops_total) / 1e9, "applies_per_step": numops_total / args.steps,
            "setup_s": t_setup,
            "krylov_vec": {"algorithmic_gb_per_step": cnt["vec_bytes"] / args.steps / 1e9},
            "transform": {"algorithmic_gb_per_step": cnt["transform_bytes"] / args.steps / 1e9},
            "roofline_hbm": hbm_rooflines(cnt, prof, args.steps),
            "device_ms_per_step_by_kernel_class": {k: v / args.steps for k, v in prof["category_ms"].items()},
            "device_ms_per_step_by_phase_and_kernel_class": {ph: {k: v / args.steps for k, v in d.items()}
                                                             for ph, d in sysenv.phase_kernel_ms.items()},
            "device_ms_per_step_by_rank": per_rank,
            "collectives_rank0_per_step": {ph: {k: {"ms": v[0] / args.steps, "calls": v[1] / args.steps} for k, v in d.items()}
                                           for ph, d in sysenv.collective_detail.items() if d},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": tf_e2e, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "call": e2e_path, "result_norm": e2e_norm},
            "gpu_launches": int(cnt["launches"]), "clocks": clocks}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chi", type=int, default=4096)
    ap.add_argument("--nsites", type=int, default=100)
    ap.add_argument("--cutoff", type=float, default=1e-15)
    ap.add_argument("--decomp", default="auto", choices=["auto", "svd", "eigen"])
    ap.add_argument("--svd-alg", dest="svd_alg", default="divide_and_conquer",
                    choices=["divide_and_conquer", "qr_iteration", "polar", "gram"],
                    help="SVD driver of the truncation: the product default `divide_and_conquer` (Gram eigenproblem on the "
                         "grouped DGEMM + deflated refinement; `gram` is an alias), gesvd `qr_iteration`, gesvdp `polar`")
    ap.add_argument("--multi", default="sharded", choices=["sharded", "replicas"],
                    help="N>1: shard the H_eff apply of ONE sweep over the GPUs (strong scaling) or run independent replicas")
    ap.add_argument("--reserve-gb", dest="reserve_gb", type=float, default=40.0,
                    help="device memory pool head-room reserved before the timed region")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the measured full sweep (198 bond updates)")
    ap.add_argument("--no-decaying", dest="decaying", action="store_false",
                    help="skip the truncation timing on a decaying Schmidt spectrum")
    ap.add_argument("--profile-region", action="store_true", help="cudaProfilerStart/Stop around the timed region (for ncu)")
    ap.add_argument("--workload", default="heisenberg_s1", choices=["heisenberg_s1", "j1j2_cylinder", "hubbard_tdvp", "ttn_tfi"],
                    help="heisenberg_s1 = the headline line (BASELINE.json configs[1]); the others are configs[2..4] on one "
                         "GPU (tools/bench_workloads.py), own arm only")
    args = ap.parse_args()
    if args.workload != "heisenberg_s1" and args.impl == "ours":
        from tools import bench_workloads
        bench_workloads.run(args, ClockSampler, fp64_peak)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

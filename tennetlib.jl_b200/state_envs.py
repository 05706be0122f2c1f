"""`StateEnvs` on the GPU -- host-side mirror of /root/reference/src/mps/state_envs.jl:18-27,54-60,
352-378 for PH = ProjMPO.  The MPS, the MPO and all environments live in HBM behind a `tnl_env_t`."""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np

from ._lib import check
from .couplingmodel import canonical_terms, is_coupling_model
from .tensor import Context, DeviceTensor, HostTensor, _index_array, flatten_blocks


class StateEnvs:
    """StateEnvs(psi::MPS, H::MPO) or StateEnvs(psi::MPS, Hs::Vector{MPO}) (sum of MPOs).  `psi`: sequence of site tensors A_j(l, s, r) (host tensors or
    DeviceTensors); `H`: sequence of MPO tensors W_j(wl, s', s, wr) (host).  As in the reference the state is
    copied on construction (state_envs.jl:59)."""

    def __init__(self, ctx: Context, psi: Sequence, H: Sequence, llim: int = 0, rlim: int | None = None,
                 copy: bool = True, Ms: Sequence | None = None, weight: float = -1.0):
        self.ctx = ctx
        self.profile = False           # when True, per-phase device times are accumulated in phase_ms
        self.phase_ms = {}
        self.phase_log = []            # (phase, ms) in call order: shows one-off hiccups that an average hides
        self.profile_kernels = False   # additionally split the device time of each phase by kernel class
        self.phase_kernel_ms = {}
        self.collective_detail = {}     # phase -> kind -> [device ms, calls] (multi-GPU, when profile_kernels)
        self.gemm_prof = dict(total_ms=0.0, launches=0, flops=0.0, max_tflops=0.0)
        self.last_solver_info = {}
        self.N = len(psi)
        # StateEnvs(psi, H::MPO) -> ProjMPO ; StateEnvs(psi, Hs::Vector{MPO}) -> ProjMPOSum2 ;
        # StateEnvs(psi, H::CouplingModel) -> ProjCouplingModel (state_envs.jl:54-79)
        self.is_coupling_model = is_coupling_model(H)
        if self.is_coupling_model:
            Hs, cm_terms = [], canonical_terms(H)
            if len(cm_terms) != self.N:
                raise ValueError("MPS and CouplingModel lengths differ")
        else:
            Hs = list(H) if len(H) and isinstance(H[0], (list, tuple)) else [H]
        if any(len(Hk) != self.N for Hk in Hs):
            raise ValueError("MPS and MPO lengths differ")
        self.nterms = len(Hs)
        self.H_host = Hs[0] if len(Hs) == 1 else None      # krylov_extend! works on the MPO itself (sweep.jl:446)
        h = C.c_void_p()
        check(ctx.lib.tnl_env_create(ctx.h, self.N, C.byref(h)), ctx.h)
        self.h = h
        for k, Hk in enumerate(Hs):
            for j, W in enumerate(Hk):
                arr, nq, keep = _index_array(W.inds)
                coords, offsets, data, nb = flatten_blocks(W)
                check(ctx.lib.tnl_env_set_site_op_term(self.h, k, j + 1, nq, arr, nb, coords.ctypes.data,
                                                       offsets.ctypes.data, data.ctypes.data), ctx.h)
        if self.is_coupling_model:
            for j, terms in enumerate(cm_terms):
                for tid, W in terms.items():
                    arr, nq, keep = _index_array(W.inds)
                    coords, offsets, data, nb = flatten_blocks(W)
                    check(ctx.lib.tnl_env_cm_set_term(self.h, j + 1, int(tid), int(W.has_wl), int(W.has_wr), nq, arr, nb,
                                                      coords.ctypes.data,
                                                      offsets.ctypes.data, data.ctypes.data), ctx.h)
        for j, A in enumerate(psi):
            dt = (A.copy() if copy else A) if isinstance(A, DeviceTensor) else DeviceTensor.from_host(ctx, A, nrow=2)
            check(ctx.lib.tnl_env_set_state(self.h, j + 1, dt.h), ctx.h)
        # StateEnvs(psi, H, Ms; weight) (state_envs.jl:86-150): penalised states for excited-state DMRG; the same
        # rank-1 terms on top of an MPO (ProjMPO_MPS2), a sum of MPOs (ProjMPOSum_MPS) or a CouplingModel
        # (ProjCouplingModel_MPS)
        self.has_penalty = bool(Ms)
        if Ms:
            if weight <= 0.0:
                raise ValueError(f"`weight` parameter should be > 0.0 (value passed was `weight={weight}`)")
            for M in Ms:
                if len(M) != self.N:
                    raise ValueError("penalised MPS has the wrong length")
                dts = [m if isinstance(m, DeviceTensor) else DeviceTensor.from_host(ctx, m, nrow=2) for m in M]
                arr = (C.c_void_p * self.N)(*[d.h for d in dts])
                check(ctx.lib.tnl_env_add_penalty(self.h, float(weight), self.N, arr), ctx.h)
        # orthogonality limits of the MPS (ITensorMPS llim / rlim)
        self.llim = llim
        self.rlim = self.N + 1 if rlim is None else rlim
        self._nsite = 2

    def updateH(self, H, Ms: Sequence | None = None, weight: float = -1.0, recalcEnv: bool = True):
        """`updateH!` (src/mps/state_envs.jl:181-330).  recalcEnv = True: a fresh projected Hamiltonian over the same
        (shared, not copied) device state -- all environments are rebuilt on demand.  recalcEnv = False (single MPO
        only, as in the reference): only the site operators are replaced, every cached environment is kept."""
        if recalcEnv:
            new = StateEnvs(self.ctx, [self.site_tensor(j) for j in range(1, self.N + 1)], H, llim=self.llim,
                            rlim=self.rlim, copy=False, Ms=Ms, weight=weight)
            new.set_nsite(self._nsite)
            self.h, new.h = new.h, self.h                      # the old tnl_env_t goes away with `new`
            for k in ("nterms", "H_host", "is_coupling_model", "has_penalty"):
                setattr(self, k, getattr(new, k))
            return self
        if self.nterms != 1 or self.is_coupling_model or self.has_penalty or Ms:
            raise RuntimeError(f"`updateH!()` :: Not implemented for `recalcEnv={recalcEnv}` with this `StateEnvs` !!")
        if len(H) != self.N:
            raise ValueError("MPS and MPO lengths differ")
        if (not self.isortho()) or self.orthocenter() != 1:
            self.orthogonalize1()
        for j, W in enumerate(H):
            arr, nq, keep = _index_array(W.inds)
            coords, offsets, data, nb = flatten_blocks(W)
            check(self.ctx.lib.tnl_env_update_site_op(self.h, j + 1, nq, arr, nb, coords.ctypes.data, offsets.ctypes.data,
                                                      data.ctypes.data), self.ctx.h)
        self.H_host = H
        return self

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.tnl_env_destroy(self.h)
        except Exception:
            pass
        self.h = None

    def __len__(self):
        return self.N

    def phase(self, name: str):
        """Context manager: CUDA-event time of a phase on the library stream (only when self.profile)."""
        env = self

        class _P:
            def __enter__(self_p):
                if env.profile:
                    env.ctx.timer_start()

            def __exit__(self_p, *exc):
                if env.profile and exc[0] is None:
                    dt = env.ctx.timer_stop()
                    env.phase_ms[name] = env.phase_ms.get(name, 0.0) + dt
                    env.phase_log.append((name, dt))
                    if env.profile_kernels:          # device time by kernel class inside this phase
                        pr = env.ctx.profile_read()
                        acc = env.phase_kernel_ms.setdefault(name, {})
                        for k, v in pr["category_ms"].items():
                            acc[k] = acc.get(k, 0.0) + v
                        cd = env.collective_detail.setdefault(name, {})
                        for k, v in pr["collective_ms"].items():
                            if pr["collective_calls"][k]:
                                e = cd.setdefault(k, [0.0, 0])
                                e[0] += v; e[1] += pr["collective_calls"][k]
                        g = env.gemm_prof
                        g["total_ms"] += pr["total_ms"]; g["launches"] += pr["launches"]; g["flops"] += pr["flops"]
                        g["max_tflops"] = max(g["max_tflops"], pr["max_tflops"])
                return False
        return _P()

    # ---- MPS bookkeeping
    def isortho(self) -> bool:
        return self.llim + 2 == self.rlim

    def orthocenter(self) -> int:
        if not self.isortho():
            raise RuntimeError("MPS has no well-defined orthogonality centre")
        return self.llim + 1

    def orthogonalize(self, j: int):
        """orthogonalize!(psi, j) (sweep.jl:100-102): QR gauge moves on the device."""
        if self.llim < j - 1:
            check(self.ctx.lib.tnl_env_move_center(self.h, self.llim + 1, j), self.ctx.h)
        if self.rlim > j + 1:
            check(self.ctx.lib.tnl_env_move_center(self.h, self.rlim - 1, j), self.ctx.h)
        self.llim, self.rlim = j - 1, j + 1

    def orthogonalize1(self):
        self.orthogonalize(1)

    def site_tensor(self, j: int) -> DeviceTensor:
        h = C.c_void_p()
        check(self.ctx.lib.tnl_env_get_state(self.h, j, C.byref(h)), self.ctx.h)
        return DeviceTensor(self.ctx, h)

    def set_site_tensor(self, j: int, t: DeviceTensor):
        """sysenv.psi[j] = t (shares the device tensor; environments that contain site j are invalidated)."""
        check(self.ctx.lib.tnl_env_set_state(self.h, j, t.h), self.ctx.h)

    def svd_split(self, pos: int, phi: DeviceTensor, *, maxdim, mindim, cutoff, ortho, normalize,
                  svd_alg="divide_and_conquer", absorb: bool = True):
        """One-site tail (update_site.jl:158-186): psi[pos] = U, psi[posnext] = (S*V)*psi[posnext].
        absorb=False (TDVP): returns (truncerr, eigs, S*V) and leaves psi[posnext] alone -- see absorb_bond."""
        cap = 1 << 16
        eigs = np.zeros(cap)
        truncerr = C.c_double()
        neigs = C.c_int64()
        carry = C.c_void_p()
        md = 0 if maxdim is None or maxdim >= (1 << 62) else int(maxdim)
        alg = {"divide_and_conquer": 0, "recursive": 0, "polar": 1, "gram": 2, "qr_iteration": 3}[svd_alg]
        check(self.ctx.lib.tnl_svd_split(self.h, pos, phi.h, 1 if ortho == "left" else 0, md, int(mindim), float(cutoff),
                                         1 if normalize else 0, alg, C.byref(truncerr), eigs.ctypes.data, cap,
                                         C.byref(neigs), None if absorb else C.byref(carry)), self.ctx.h)
        if ortho == "left":
            self.llim, self.rlim = pos, pos + 2                  # setleftlim!(psi, pos)
        else:
            self.llim, self.rlim = pos - 2, pos                  # setrightlim!(psi, pos)
        out = (truncerr.value, eigs[:min(neigs.value, cap)].copy())
        return out if absorb else out + (DeviceTensor(self.ctx, carry),)

    def absorb_bond(self, pos: int, ortho: str, carry: DeviceTensor):
        """psi[posnext] = carry * psi[posnext] (update_site.jl:186)."""
        check(self.ctx.lib.tnl_env_absorb_bond(self.h, pos, 1 if ortho == "left" else 0, carry.h), self.ctx.h)

    def getpsi(self) -> List[HostTensor]:
        """getpsi (state_envs.jl:36): host copy of the MPS."""
        return [self.site_tensor(j).to_host() for j in range(1, self.N + 1)]

    def linkdim(self, bond: int) -> int:
        return self.site_tensor(bond).inds[2].dim

    def linkdims(self) -> List[int]:
        return [self.site_tensor(j).inds[2].dim for j in range(1, self.N)]

    # ---- environment interface
    def nsite(self) -> int:
        return self._nsite

    def set_nsite(self, n: int):
        check(self.ctx.lib.tnl_env_set_nsite(self.h, n), self.ctx.h)
        self._nsite = n
        return self

    def position(self, pos: int):
        check(self.ctx.lib.tnl_env_position(self.h, pos), self.ctx.h)
        return self

    def make_phi(self, pos: int) -> DeviceTensor:
        h = C.c_void_p()
        check(self.ctx.lib.tnl_env_make_phi(self.h, pos, C.byref(h)), self.ctx.h)
        return DeviceTensor(self.ctx, h)

    def product(self, v: DeviceTensor) -> DeviceTensor:
        h = C.c_void_p()
        check(self.ctx.lib.tnl_heff_apply(self.h, v.h, C.byref(h)), self.ctx.h)
        return DeviceTensor(self.ctx, h, v._inds)

    __call__ = product

    def apply_flops(self) -> float:
        out = C.c_double()
        check(self.ctx.lib.tnl_env_apply_flops(self.h, C.byref(out)), self.ctx.h)
        return out.value

    def expectation(self, phi: DeviceTensor) -> float:
        out = C.c_double()
        check(self.ctx.lib.tnl_expectation(self.h, phi.h, C.byref(out)), self.ctx.h)
        return out.value

    def replacebond(self, pos: int, phi: DeviceTensor, *, maxdim, mindim, cutoff, noise, ortho, normalize,
                    which_decomp=None, svd_alg="divide_and_conquer"):
        """noiseterm + replacebond! (update_site.jl:59-76).  Returns (truncerr, eigs)."""
        N = self.N
        cap = 1 << 16
        eigs = np.zeros(cap)
        truncerr = C.c_double()
        neigs = C.c_int64()
        which = {None: 0, "svd": 1, "eigen": 2}[which_decomp]
        which |= {"divide_and_conquer": 0, "recursive": 0, "polar": 1, "gram": 2, "qr_iteration": 3}[svd_alg] << 4
        md = 0 if maxdim is None or maxdim >= (1 << 62) else int(maxdim)
        check(self.ctx.lib.tnl_replacebond(self.h, pos, phi.h, 1 if ortho == "left" else 0, md, int(mindim),
                                           float(cutoff), float(noise), 1 if normalize else 0, which,
                                           C.byref(truncerr), eigs.ctypes.data, cap, C.byref(neigs)), self.ctx.h)
        if ortho == "left":
            if self.llim == pos - 1:
                self.llim += 1
            if self.rlim == pos + 1:
                self.rlim += 1
        else:
            if self.llim == pos:
                self.llim -= 1
            if self.rlim == pos + 2:
                self.rlim -= 1
        return truncerr.value, eigs[:min(neigs.value, cap)].copy()

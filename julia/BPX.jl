# BPX.jl -- thin `ccall` glue between ITensorNetworksNext.jl and libbpx.so (include/bpx.h).
#
# NOT RUNNABLE IN THIS IMAGE (no Julia toolchain, SURVEY.md F9): everything behavioural lives behind the C ABI
# and is tested from Python through the very same entry points (itensornetworksnext.jl_b200/_lib.py).  This file
# shows the reference-side binding a maintainer adds; it touches no reference source.
#
# Plug-in points used (all in /root/reference):
#   * `MessageUpdateAlgorithm` strategy interface ........ src/beliefpropagation/beliefpropagation.jl:214-220
#   * instance pass-through of `select_algorithm` ........ src/select_algorithm.jl:41-48
#   * nested `AI.step!` (one outer step = one sweep) ..... src/AlgorithmsInterfaceExtensions/AlgorithmsInterfaceExtensions.jl:27-32
#   * `AIE.iterate_diff` used by `StopWhenConverged` ...... src/beliefpropagation/beliefpropagation.jl:261-267
module BPX

using ITensorNetworksNext: ITensorNetworksNext, NormNetwork, MessageCache, MessageUpdateAlgorithm,
    BeliefPropagationProblem, BeliefPropagationAlgorithm, BeliefPropagationSweepAlgorithm,
    kettensor, braname, linknames, sitenames
import ITensorNetworksNext: message_update!
using ITensorNetworksNext.AlgorithmsInterfaceExtensions: AlgorithmsInterfaceExtensions as AIE
import AlgorithmsInterface as AI
using Graphs: vertices, src, dst, neighbors
using NamedGraphs: NamedEdge
using ITensorBase: ITensorBase, ITensor, dimnames, unnamed

const libbpx = get(ENV, "LIBBPX", "libbpx.so")
const BPX_F64, BPX_C64 = Cint(0), Cint(1)
const BPX_MODE_NORM = Cint(0)

struct BPXError <: Exception
    status::Cint
    msg::String
end
function check(ctx::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:bpx_last_error, libbpx), Cstring, (Ptr{Cvoid},), ctx))
    return throw(BPXError(rc, msg))
end

"""
    B200MessageUpdate(; normalize = true, devices = [0])

Message-update strategy that runs whole synchronous BP sweeps on one or several B200s of this process through libbpx.
Pass it as `beliefpropagation(nn, messages; message_update_algorithm = B200MessageUpdate(devices = 0:7), ...)`: ONE
context over the device list (`bpx_create_multi`), the library partitions the vertices and exchanges cut-edge messages
over NVLink inside the sweep kernels -- no MPI, no second Julia process.  The context is created lazily on the first
sweep and kept in the (mutable) strategy object.

Two ways to hold the iterate (SURVEY.md 8 b2):
  (A) ordinary ITensor messages on the host: works with the stock `StopWhenConverged`; every outer iteration is one
      `bpx_sweep_host` call (packed iterate in, packed iterate + fused residual out);
  (B) `device_iterate(alg, nn, messages)`: the messages handed to `beliefpropagation` are `DeviceMessageRef`s (edge id +
      context).  `copy(iterate)` copies refs, `AIE.iterate_diff` returns the residual fused into the last sweep's kernels,
      a sweep is one `bpx_sweep` call without any message traffic, and `materialize(cache)` / `ITensor(ref)` download on
      demand.  `solve_resident!(alg, nn, messages; maxiter, tol)` runs the whole `(; maxiter, tol)` loop in ONE call with
      the convergence test on the device.
"""
mutable struct B200MessageUpdate <: MessageUpdateAlgorithm
    normalize::Bool
    devices::Vector{Cint}
    ctx::Ptr{Cvoid}
    edge_ids::Dict{Any, Int}      # NamedEdge -> directed edge id of the C ABI
    vertex_ids::Dict{Any, Int}
    bra_ket::Vector{Tuple{Any, Any}}  # per directed edge: (bra name, ket name) of the message axes
    dirty::Bool                   # host cache newer than the device copy
    last_residual::Float64
    host_in::Vector               # page-locked packed iterates (bpx_host_register): bpx_sweep_host streams through them
    host_out::Vector
end
B200MessageUpdate(; normalize = true, devices = [0]) =
    B200MessageUpdate(normalize, Cint.(collect(devices)), C_NULL, Dict(), Dict(), Tuple{Any, Any}[], true, Inf, Float64[], Float64[])

# Two packed host iterates, page-locked once: with such buffers `bpx_sweep_host` overlaps the upload with the sweep
# kernel and lets the kernel store the new messages straight into `host_out` (include/bpx.h).
function host_buffers!(alg::B200MessageUpdate, ::Type{E}) where {E}
    total = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), alg.ctx, length(alg.edge_ids))
    if length(alg.host_in) != total || eltype(alg.host_in) != E
        alg.host_in, alg.host_out = Vector{E}(undef, total), Vector{E}(undef, total)
        for buf in (alg.host_in, alg.host_out)
            check(alg.ctx, ccall((:bpx_host_register, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t), alg.ctx, buf, sizeof(buf)))
        end
    end
    return alg
end

# ---- lowering to the canonical layout (include/bpx.h; SURVEY.md §8 b2 iv) -------------------------------
function upload!(alg::B200MessageUpdate, nn::NormNetwork, cache::MessageCache)
    vs = collect(vertices(nn))
    alg.vertex_ids = Dict(v => i - 1 for (i, v) in enumerate(vs))
    srcs, dsts, slots = Int64[], Int64[], Int32[]
    for v in vs, (k, w) in enumerate(neighbors(nn, v))
        alg.edge_ids[NamedEdge(v => w)] = length(srcs)
        push!(srcs, alg.vertex_ids[v]); push!(dsts, alg.vertex_ids[w]); push!(slots, k - 1)
    end
    ctxref = Ref{Ptr{Cvoid}}(C_NULL)
    # one context over the whole device list (a single device is the list of one)
    rc = ccall((:bpx_create_multi, libbpx), Cint, (Ptr{Cint}, Cint, Ref{Ptr{Cvoid}}), alg.devices, length(alg.devices), ctxref)
    rc == 0 || throw(BPXError(rc, unsafe_string(ccall((:bpx_last_error, libbpx), Cstring, (Ptr{Cvoid},), C_NULL))))
    ctx = alg.ctx = ctxref[]
    check(ctx, ccall((:bpx_set_graph, libbpx), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}),
        ctx, length(vs), length(srcs), srcs, dsts, slots))
    T = eltype(unnamed(kettensor(nn, first(vs))))
    dtype = T <: Complex ? BPX_C64 : BPX_F64
    E = T <: Complex ? ComplexF64 : Float64
    phys, link = Int32[], Int32[]
    sites = E[]
    alg.bra_ket = Tuple{Any, Any}[]
    for v in vs
        A = kettensor(nn, v)                                   # normnetwork.jl:77
        snames = collect(sitenames(nn.ket, v))
        lnames = [only(linknames(nn.ket, NamedEdge(v => w))) for w in neighbors(nn, v)]
        arr = Array{E}(unnamed(A, (snames..., lnames...)))      # permute to [sites..., links in neighbour order]
        push!(phys, prod(size(arr)[1:length(snames)]; init = 1))
        for (w, ln) in zip(neighbors(nn, v), lnames)
            push!(link, size(arr, length(snames) + findfirst(==(ln), lnames)))
            push!(alg.bra_ket, (braname(nn, ln), ln))
        end
        append!(sites, vec(arr))
    end
    check(ctx, ccall((:bpx_set_dims, libbpx), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Int32}),
        ctx, dtype, BPX_MODE_NORM, phys, link))
    GC.@preserve sites check(ctx, ccall((:bpx_set_site_tensors, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx, sites))
    push_messages!(alg, cache, E)
    return alg
end

function push_messages!(alg::B200MessageUpdate, cache::MessageCache, ::Type{E}) where {E}
    packed = E[]
    for (e, id) in sort(collect(alg.edge_ids); by = last)
        bra, ket = alg.bra_ket[id + 1]
        append!(packed, vec(Array{E}(unnamed(cache[e], (bra, ket)))))   # canonicalise to [bra, ket]
    end
    GC.@preserve packed check(alg.ctx, ccall((:bpx_set_messages, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), alg.ctx, packed))
    alg.dirty = false
    return alg
end

# packed host iterate <-> MessageCache (same layout as push_messages! / pull_messages!)
function pack_messages!(packed::Vector, alg::B200MessageUpdate, cache::MessageCache)
    for (e, id) in alg.edge_ids
        off = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), alg.ctx, id)
        bra, ket = alg.bra_ket[id + 1]
        m = vec(Array{eltype(packed)}(unnamed(cache[e], (bra, ket))))
        copyto!(packed, off + 1, m, 1, length(m))
    end
    return packed
end
function unpack_messages!(cache::MessageCache, alg::B200MessageUpdate, packed::Vector)
    for (e, id) in alg.edge_ids
        off = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), alg.ctx, id)
        bra, ket = alg.bra_ket[id + 1]
        χ = size(unnamed(cache[e], (bra, ket)), 1)
        cache[e] = ITensor(reshape(packed[(off + 1):(off + χ * χ)], χ, χ), (bra, ket))  # messagecache.jl:92-96
    end
    return cache
end

function pull_messages!(alg::B200MessageUpdate, cache::MessageCache)
    ne = length(alg.edge_ids)
    total = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), alg.ctx, ne)
    E = eltype(unnamed(first(values(cache.messages))))
    E = E <: Complex ? ComplexF64 : Float64
    packed = Vector{E}(undef, total)
    GC.@preserve packed check(alg.ctx, ccall((:bpx_get_messages, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), alg.ctx, packed))
    for (e, id) in alg.edge_ids
        off = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), alg.ctx, id)
        bra, ket = alg.bra_ket[id + 1]
        old = cache[e]
        χ = size(unnamed(old, (bra, ket)), 1)
        cache[e] = ITensor(reshape(packed[(off + 1):(off + χ * χ)], χ, χ), (bra, ket))  # messagecache.jl:92-96
    end
    return cache
end

# ---- one outer iteration = one synchronous sweep in ONE ccall --------------------------------------------
# More specific than the NestedAlgorithm method at AIE.jl:27, so dispatch picks it whenever the sweep's
# strategy is a B200MessageUpdate; `beliefpropagation()` itself is unchanged.
function AI.step!(
        problem::BeliefPropagationProblem,
        algorithm::BeliefPropagationAlgorithm{<:Any, <:BeliefPropagationSweepAlgorithm{<:B200MessageUpdate}},
        state::AI.State
    )
    alg = algorithm.subalgorithm.message_update_algorithm
    cache = state.iterate
    alg.ctx == C_NULL && upload!(alg, problem.factors, cache)
    # Variant (A) of SURVEY.md §8 b2: plain ITensor messages that live on the host, so that the stock
    # `StopWhenConverged` (AIE.jl:84-109) keeps working: ONE `bpx_sweep_host` call per outer iteration (packed
    # iterate in, packed iterate + fused residual out).  Use `B200Converged` with `bpx_sweep` to keep them resident.
    E = eltype(unnamed(first(values(cache.messages)))) <: Complex ? ComplexF64 : Float64
    host_buffers!(alg, E)
    pack_messages!(alg.host_in, alg, cache)
    res = Ref{Cdouble}(Inf)
    GC.@preserve alg check(alg.ctx, ccall((:bpx_sweep_host, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Cdouble}),
        alg.ctx, alg.host_in, alg.host_out, alg.normalize, res))
    alg.last_residual = res[]
    unpack_messages!(cache, alg, alg.host_out)
    return state
end

# ---- variant (B): device-resident iterate --------------------------------------------------------------------
"""A message that lives on the device: what `MessageCache` holds when the iterate is resident (valtype of the messages
handed to `beliefpropagation`, messagecache.jl:33-49)."""
struct DeviceMessageRef
    alg::B200MessageUpdate
    edge_id::Int
end

"download one message as an ordinary ITensor (axes (bra, ket) like messagecache.jl:211-218)"
function ITensorBase.ITensor(r::DeviceMessageRef)
    ctx = r.alg.ctx
    n = ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), ctx, r.edge_id + 1) -
        ccall((:bpx_message_offset, libbpx), Int64, (Ptr{Cvoid}, Int64), ctx, r.edge_id)
    E = r.alg.host_in isa Vector{ComplexF64} ? ComplexF64 : Float64
    buf = Vector{E}(undef, n)
    GC.@preserve buf check(ctx, ccall((:bpx_get_message, libbpx), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}), ctx, r.edge_id, buf))
    χ = isqrt(n)
    return ITensor(reshape(buf, χ, χ), r.alg.bra_ket[r.edge_id + 1])
end

"""
    device_iterate(alg, nn, messages) -> messages of `DeviceMessageRef`

Upload the network and the initial messages once and return the resident iterate to pass to `beliefpropagation`.
"""
function device_iterate(alg::B200MessageUpdate, nn::NormNetwork, messages)
    cache = MessageCache(messages)
    alg.ctx == C_NULL ? upload!(alg, nn, cache) : push_messages!(alg, cache, eltype(alg.host_in) <: Complex ? ComplexF64 : Float64)
    return Dict(e => DeviceMessageRef(alg, id) for (e, id) in alg.edge_ids)
end

# `copy(iterate)` runs every outer iteration (AIE.jl:74, 96): refs are immutable handles, the cache itself is the copy
Base.copy(c::MessageCache{DeviceMessageRef}) = c
# ... and the comparison of two iterates is the residual the sweep kernels fused into their epilogues (bp.jl:261-267)
function AIE.iterate_diff(a::MessageCache{DeviceMessageRef}, ::MessageCache{DeviceMessageRef})
    return first(values(a.messages)).alg.last_residual
end

"all messages as ordinary ITensors: one `bpx_get_messages`"
function materialize(c::MessageCache{DeviceMessageRef})
    alg = first(values(c.messages)).alg
    return MessageCache(Dict(e => ITensor(r) for (e, r) in pairs(c.messages)))   # (bulk variant: pull_messages!)
end

# one outer iteration on a resident iterate: one sweep, no message traffic; the residual comes back with the call
function AI.step!(
        problem::BeliefPropagationProblem,
        algorithm::BeliefPropagationAlgorithm{<:Any, <:BeliefPropagationSweepAlgorithm{<:B200MessageUpdate}},
        state::AI.State{<:MessageCache{DeviceMessageRef}}
    )
    alg = algorithm.subalgorithm.message_update_algorithm
    res, done = Ref{Cdouble}(Inf), Ref{Cint}(0)
    check(alg.ctx, ccall((:bpx_sweep, libbpx), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cint, Ref{Cdouble}, Ref{Cint}),
        alg.ctx, 1, 0.0, alg.normalize, res, done))
    alg.last_residual = res[]
    return state
end

"""
    solve_resident!(alg, nn, messages; maxiter, tol) -> (cache, sweeps, residual)

The `(; maxiter, tol)` form of `beliefpropagation` (beliefpropagation.jl:46-54) in ONE call: `StopAfterIteration(maxiter) |
StopWhenConverged(tol)` is evaluated on the device (sweeps are enqueued in batches, a converged run turns the remaining
launches into no-ops), so the host neither copies nor compares iterates.
"""
function solve_resident!(alg::B200MessageUpdate, nn::NormNetwork, messages; maxiter::Int, tol::Float64)
    refs = device_iterate(alg, nn, messages)
    res, done = Ref{Cdouble}(Inf), Ref{Cint}(0)
    check(alg.ctx, ccall((:bpx_sweep, libbpx), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cint, Ref{Cdouble}, Ref{Cint}),
        alg.ctx, maxiter, tol, alg.normalize, res, done))
    alg.last_residual = res[]
    return materialize(MessageCache(refs)), Int(done[]), res[]
end

# Per-edge entry kept for API completeness (`message_update!(alg, cache, factors, edge)`, beliefpropagation.jl:242):
# a one-edge sequential "sweep" on the device.
function message_update!(alg::B200MessageUpdate, cache, factors, edge)
    alg.ctx == C_NULL && upload!(alg, factors, cache)
    seq = Int64[alg.edge_ids[edge]]
    check(alg.ctx, ccall((:bpx_sweep_sequence, libbpx), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Int64, Cint, Cdouble, Cint, Ptr{Cdouble}, Ptr{Cint}),
        alg.ctx, seq, 1, 1, 0.0, alg.normalize, C_NULL, C_NULL))
    return pull_messages!(alg, cache)
end

"""
    B200Converged(tol, alg)

Stopping criterion that consumes the residual fused into the sweep kernel's epilogue instead of a second pass
over two host-side caches (`iterate_diff`, beliefpropagation.jl:261-267).  Explicit criteria are accepted as-is
by `beliefpropagation` (beliefpropagation.jl:16).
"""
struct B200Converged <: AI.StoppingCriterion
    tol::Float64
    alg::B200MessageUpdate
end
mutable struct B200ConvergedState <: AI.StoppingCriterionState
    delta::Float64
    at_iteration::Int
end
AI.initialize_state(::AI.Problem, ::AI.Algorithm, ::B200Converged; kwargs...) = B200ConvergedState(Inf, -1)
AI.initialize_state!(::AI.Problem, ::AI.Algorithm, ::B200Converged, st::B200ConvergedState) = (st.delta = Inf; st)
function AI.is_finished!(::AI.Problem, ::AI.Algorithm, state::AI.State, c::B200Converged, st::B200ConvergedState)
    state.iteration == 0 && return false
    st.delta = c.alg.last_residual
    st.delta < c.tol || return false
    st.at_iteration = state.iteration
    return true
end
AI.is_finished(::AI.Problem, ::AI.Algorithm, ::AI.State, c::B200Converged, st::B200ConvergedState) = st.delta < c.tol

# beliefs on the device (messagecache.jl:139-201)
function vertex_scalars(alg::B200MessageUpdate, ::Type{E} = Float64) where {E}
    out = Vector{E}(undef, length(alg.vertex_ids))
    GC.@preserve out check(alg.ctx, ccall((:bpx_vertex_scalars, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), alg.ctx, out))
    return out
end
function edge_scalars(alg::B200MessageUpdate, ::Type{E} = Float64) where {E}
    out = Vector{E}(undef, length(alg.edge_ids) ÷ 2)
    GC.@preserve out check(alg.ctx, ccall((:bpx_edge_scalars, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), alg.ctx, out))
    return out
end

"""
    bethe_free_energy(alg::B200MessageUpdate)
`bethe_free_energy(factors, messages)` (messagecache.jl:185-201) of the resident iterate, reduced on the device: a
`Float64`, or a `ComplexF64` when the reference's promotion rule applies (complex element type or a negative term).
"""
function bethe_free_energy(alg::B200MessageUpdate)
    out = zeros(Float64, 2)
    promoted = Ref{Cint}(0)
    GC.@preserve out check(alg.ctx, ccall((:bpx_bethe_free_energy, libbpx), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cint}),
        alg.ctx, out, promoted))
    return promoted[] != 0 ? complex(out[1], out[2]) : out[1]
end

# ---- gate application on the device (src/apply/apply_operators.jl:150-283) ---------------------------------------
# Plug-in point: "abstract type ApplyOperatorAlgorithm" (:150) with `apply_operator!(algorithm, dest, operator, state, env)`
# (:185-194) and `initialize_output` (:204-208); an instance passed as `apply_operator(op, state, env; alg = X)` reaches
# `apply_operator(algorithm::ApplyOperatorAlgorithm, ...)` (:173-176) untouched (select_algorithm.jl:41-48).
import ITensorNetworksNext: apply_operator!, initialize_output
using ITensorNetworksNext: ApplyOperatorAlgorithm, normnetwork
using ITensorBase: domainnames

"""
    B200ApplyGate(; trunc = nothing, normalize = false, bp = B200MessageUpdate())

BP simple-update gate application (`BPApplyGate`, apply_operators.jl:180-283) on the B200 that `bp` is bound to: state
and environment are uploaded once (`upload!`), every gate -- or, through `apply_layer!`, every layer of vertex-disjoint
gates -- is one `ccall`, and tensors / messages are pulled back when the caller asks for them.
"""
Base.@kwdef mutable struct B200ApplyGate <: ApplyOperatorAlgorithm
    trunc::Union{Nothing, Int} = nothing
    normalize::Bool = false
    bp::B200MessageUpdate = B200MessageUpdate()
    resident::Tuple{UInt, UInt} = (UInt(0), UInt(0))   # objectid of the (state, env) the device copy mirrors
end

initialize_output(::typeof(apply_operator!), ::B200ApplyGate, operator, state, env) = copy(state), copy(env)

# operator -> [o1, o2, i1, i2] (column-major; 1 = src, 2 = dst of the gate edge), outputs first like ITensorBase operators
function lowered_operator(op, state, vs, ::Type{E}) where {E}
    ins = [only(intersect(domainnames(op), sitenames(state, v))) for v in vs]
    outs = [n for n in dimnames(op) if !(n in domainnames(op))]   # codomain names, same order as the domain
    return vec(Array{E}(unnamed(op, (outs..., ins...))))
end

function apply_operator!(alg::B200ApplyGate, dest, op, state, env)
    bp = alg.bp
    # state + env resident from here on -- for THESE objects: a different (or host-modified, hence re-created) state / env
    # is uploaded again instead of silently acting on the stale device copy
    if bp.ctx == C_NULL || alg.resident != (objectid(state), objectid(env))
        bp.ctx == C_NULL || close(bp)
        upload!(bp, normnetwork(state), env)
        alg.resident = (objectid(state), objectid(env))
    end
    E = eltype(unnamed(first(values(env.messages)))) <: Complex ? ComplexF64 : Float64
    vs = [v for v in vertices(state) if !isempty(intersect(domainnames(op), sitenames(state, v)))]
    isempty(vs) && throw(ArgumentError("operator shares no indices with the tensor network"))
    packed = lowered_operator(op, state, vs, E)
    if length(vs) == 1
        ids = Int64[bp.vertex_ids[only(vs)]]
        GC.@preserve ids packed check(bp.ctx, ccall((:bpx_apply_one_site_gates, libbpx), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint), bp.ctx, 1, ids, packed, alg.normalize))
    elseif length(vs) == 2
        ids = Int64[bp.edge_ids[NamedEdge(vs[1] => vs[2])]]
        χe = size(unnamed(env[NamedEdge(vs[1] => vs[2])]), 1)       # the C ABI keeps the link dimension (include/bpx.h)
        rank_bound = min(prod(size(unnamed(state[vs[1]]))), prod(size(unnamed(state[vs[2]])))) ÷ χe
        if something(alg.trunc, rank_bound) > χe && rank_bound > χe
            throw(ArgumentError("this gate would grow the bond from $χe (reference: up to $(something(alg.trunc, rank_bound))): " *
                "re-declare the link dimension (zero-padded tensors + bpx_set_dims) before the call, as apply.py `_apply_batch` does"))
        end
        sv = zeros(Float64, χe)
        GC.@preserve ids packed sv check(bp.ctx, ccall((:bpx_apply_two_site_gates, libbpx), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}),
            bp.ctx, 1, ids, packed, something(alg.trunc, 0), alg.normalize, sv))
        pull_messages!(bp, env)                                    # the two diag(S) messages of the gate edge (:273-277)
    else
        throw(ArgumentError("$(length(vs))-site gate decomposition not implemented"))
    end
    for v in vs                                                    # `dest[v] = ...` (:243, :270-271)
        old = state[v]
        names = (sitenames(state, v)..., (only(linknames(state, NamedEdge(v => w))) for w in neighbors(state, v))...)
        buf = Vector{E}(undef, length(unnamed(old)))
        GC.@preserve buf check(bp.ctx, ccall((:bpx_get_site_tensor, libbpx), Cint, (Ptr{Cvoid}, Int64, Ptr{Cvoid}),
            bp.ctx, bp.vertex_ids[v], buf))
        dest[v] = ITensor(reshape(buf, size(unnamed(old, names))), names)
    end
    return dest
end
# A whole circuit layer (vertex-disjoint gates) is ONE call: pass all edge ids / packed operators at once
# (`bpx_apply_two_site_gates(ctx, n_gates, edges, ops, ...)`); see itensornetworksnext.jl_b200/apply.py `_apply_batch`.

function Base.close(alg::B200MessageUpdate)
    alg.ctx == C_NULL || ccall((:bpx_destroy, libbpx), Cint, (Ptr{Cvoid},), alg.ctx)
    alg.ctx = C_NULL
    return nothing
end

end # module

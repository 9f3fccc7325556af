# HMMCuda.jl -- the `ccall` shim that makes libhmmcuda.so a drop-in for the hot-path methods of
# HMMSpikeSorter.jl.  It is a SUBMODULE: `include("HMMCuda.jl")` goes into src/HMMSpikeSorter.jl right after
# `include("types.jl")` (it needs `StateMatrix`), and each hot-path method of the package gets a one-line guard as
# its first statement (INTEGRATION.md lists the six lines), e.g.
#
#     function viterbi(y::AbstractArray{Float64,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64)
#         HMMCuda.enabled() && return HMMCuda.viterbi(y, lA, μ, σ)
#         ...                                   # the reference's body, untouched
#
# Nothing is overwritten: Julia >= 1.10 refuses method overwriting during precompilation, so the earlier form of
# this file (same-signature methods included at the end of the module) could not be precompiled.  Every caller
# (src/fit.jl:7,23,45,50,55; README.md:33-35; test/runtests.jl) keeps working unmodified, and
# `HMMSPIKESORTER_BACKEND=julia` (or a missing library / device) leaves the reference's own code in charge.
#
# NOTE: Julia is not installed in the environment this library is built and tested in, so this file has not been
# executed there.  It is kept mechanical on purpose: every call passes exactly the arrays the Python ctypes mirror
# (hmmspikesorter.jl_b200/__init__.py) passes to the same symbols, and that mirror is what the test-suite exercises.
#
# Memory layout facts relied upon (src/types.jl:1-9):
#   lA.states       :: Matrix{Int16}  [N x nstates], column-major, 1-based ring phases
#   lA.transitions  :: Vector{Tuple{Int64,Int64,Float64}}  isbits => contiguous 24-byte records
#   μ               :: Matrix{Float64} [K x N], column-major
module HMMCuda

import ..StateMatrix

const libhmmcuda = get(ENV, "LIBHMMCUDA", "libhmmcuda.so")

const HMM_EINVAL = 1

# ---- backend switch ------------------------------------------------------------------------------------------
# "cuda": always (errors if the library or a device is missing); "julia": never; "auto" (default): when
# libhmmcuda.so loads and reports at least one device.  Decided once, at first use.
const _enabled = Ref{Union{Nothing,Bool}}(nothing)
function enabled()
    e = _enabled[]
    e === nothing || return e
    want = lowercase(get(ENV, "HMMSPIKESORTER_BACKEND", "auto"))
    ok = false
    if want != "julia"
        try
            ok = ccall((:hmm_device_count, libhmmcuda), Cint, ()) > 0
        catch err
            want == "cuda" && rethrow(err)
        end
        want == "cuda" && !ok && error("HMMSPIKESORTER_BACKEND=cuda but libhmmcuda reports no CUDA device")
    end
    _enabled[] = ok
    ok
end
enable!(on::Bool=true) = (_enabled[] = on)

struct HmmInfo
    engine::Int32; n_chunks::Int32; fwd_repaired::Int32; bwd_repaired::Int32
    kernel_launches::Int64; device_ms::Float64; kernel_ms::Float64; top_kernel_ms::Float64
end

function _check(rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:hmm_last_error, libhmmcuda), Cstring, ()))
    rc == HMM_EINVAL ? throw(ArgumentError(msg)) : error("libhmmcuda: $msg")
end

_vec(y::Vector{Float64}) = y
_vec(y::SubArray{Float64,1,Vector{Float64},<:Tuple{UnitRange},true}) = y   # contiguous view (src/fit.jl:23)
_vec(y::AbstractVector{Float64}) = collect(y)                               # anything else is collected

# ---- viterbi(y, lA, μ, σ) -> (x, ll)                      replaces src/viterbi.jl:44-98
function viterbi(y::AbstractArray{Float64,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64)
    yv = _vec(y); T = length(yv)
    x = Vector{Int16}(undef, T); ll = Ref{Float64}(0.0)
    tr = lA.transitions
    GC.@preserve yv lA μ x begin
        _check(ccall((:hmm_viterbi_f64, libhmmcuda), Cint,
            (Ptr{Float64}, Int64, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Float64}, Float64,
             Ptr{Int16}, Ref{Float64}, Ptr{Int16}, Ptr{Float64}),
            pointer(yv), T, lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, σ,
            x, ll, C_NULL, C_NULL))
    end
    return x, ll[]
end

# ---- (x, T2, T1) form of README.md:34 / src/example.jl:32 -- trellis materialised on request
function viterbi_trellis(y::AbstractArray{Float64,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64)
    yv = _vec(y); T = length(yv)
    x = Vector{Int16}(undef, T); ll = Ref{Float64}(0.0)
    T1 = Matrix{Float64}(undef, lA.nstates, T); T2 = Matrix{Int16}(undef, lA.nstates, T)
    tr = lA.transitions
    GC.@preserve yv lA μ x T1 T2 begin
        _check(ccall((:hmm_viterbi_f64, libhmmcuda), Cint,
            (Ptr{Float64}, Int64, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Float64}, Float64,
             Ptr{Int16}, Ref{Float64}, Ptr{Int16}, Ptr{Float64}),
            pointer(yv), T, lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, σ, x, ll, T2, T1))
    end
    return x, T2, T1
end

# ---- forward / backward                                   replace src/baumwelch.jl:25-51, 73-98
for (fn, sym) in ((:forward, :hmm_forward_f64), (:backward, :hmm_backward_f64))
    @eval function $fn(V::Array{Float64,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64)
        T = length(V); out = Matrix{Float64}(undef, lA.nstates, T); tr = lA.transitions
        GC.@preserve V lA μ out begin
            _check(ccall(($(QuoteNode(sym)), libhmmcuda), Cint,
                (Ptr{Float64}, Int64, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Float64}, Float64, Ptr{Float64}),
                V, T, lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, σ, out))
        end
        out
    end
end

_nxi(lA::StateMatrix) = count(q -> q[1] == 1, lA.transitions)

# ---- update(α, β, lA, μ, σ, x) -> (lA_new, μ, σ)          replaces src/baumwelch.jl:205-309
# μ is overwritten in place exactly like the reference's fill!(μ, 0.0) (:268).
function update(α::Array{Float64,2}, β::Array{Float64,2}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64, x::Array{Float64,1})
    lp = Vector{Float64}(undef, max(_nxi(lA) - 1, 1)); pp = Vector{Float64}(undef, lA.nstates)
    s = Ref{Float64}(σ); tr = lA.transitions
    GC.@preserve α β lA μ x lp pp begin
        _check(ccall((:hmm_update_f64, libhmmcuda), Cint,
            (Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64,
             Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            α, β, length(x), lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, s, x, lp, pp))
    end
    # the UNCHANGED Julia constructor rebuilds the transitions (src/baumwelch.jl:265, src/types.jl:148-151)
    lA_new = StateMatrix(lA.states .- one(Int16), pp, lA.K, lp[1:_nxi(lA)-1]; allow_overlaps=lA.resolve_overlaps)
    lA_new, μ, s[]
end

# ---- one fused E/M step: train_model(X, lA, μ0, σ0)       replaces src/baumwelch.jl:362-370
function train_model(X::Array{Float64,1}, state_matrix::StateMatrix, μ0::Array{Float64,2}, σ0::Float64; verbose=0)
    lA = state_matrix
    lp = Vector{Float64}(undef, max(_nxi(lA) - 1, 1)); pp = Vector{Float64}(undef, lA.nstates)
    s = Ref{Float64}(σ0); ll = Ref{Float64}(0.0); tr = lA.transitions
    GC.@preserve X lA μ0 lp pp begin
        _check(ccall((:hmm_em_step_f64, libhmmcuda), Cint,
            (Ptr{Float64}, Int64, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64,
             Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}),
            X, length(X), lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ0, s, lp, pp, ll))
    end
    lA_new = StateMatrix(lA.states .- one(Int16), pp, lA.K, lp[1:_nxi(lA)-1]; allow_overlaps=lA.resolve_overlaps)
    lA_new, μ0, s[]
end
# The outer loop train_model(X, sm, μ, σ, nsteps, callback) (src/baumwelch.jl:324-354) is
# unchanged Julia: it calls the method above once per iteration, so callback(μ), yield() and
# the merge/prune phase behave exactly as before.  To keep X resident in HBM across the
# iterations, wrap the loop in hmm_train_create / hmm_train_em_step / hmm_train_destroy.

# ---- reconstruct_signal(x, lA, μ, σ)                      replaces src/reconstruction.jl:1-9
function reconstruct_signal(x::Array{T,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64) where T <: Integer
    xi = T === Int16 ? x : begin
        all(1 .<= x .<= lA.nstates) || throw(BoundsError(lA.states, (1, x)))
        convert(Vector{Int16}, x)
    end
    Y = Vector{Float64}(undef, length(xi))
    GC.@preserve xi lA μ Y begin
        _check(ccall((:hmm_reconstruct_f64, libhmmcuda), Cint,
            (Ptr{Int16}, Int64, Ptr{Int16}, Int32, Int32, Ptr{Float64}, Int32, Ptr{Float64}),
            xi, length(xi), lA.states, lA.N, lA.nstates, μ, lA.K, Y))
    end
    Y
end

# ---- optional: keep X resident in HBM across the E/M iterations -----------------------------------------
# train_model(X, sm, μ, σ, nsteps, callback) (src/baumwelch.jl:324-354) uploads X on every iteration when it
# calls the one-step method above.  `with_resident(X) do step ... end` uploads it once; `step(lA, μ, σ)` is then
# a drop-in for the one-step train_model(X, lA, μ, σ) inside the unchanged loop body.
function with_resident(f::Function, X::Array{Float64,1})
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X _check(ccall((:hmm_train_create, libhmmcuda), Cint, (Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                                X, length(X), ctx))
    step = function (lA::StateMatrix, μ::Array{Float64,2}, σ::Float64)
        lp = Vector{Float64}(undef, max(_nxi(lA) - 1, 1)); pp = Vector{Float64}(undef, lA.nstates)
        s = Ref{Float64}(σ); ll = Ref{Float64}(0.0); tr = lA.transitions
        GC.@preserve lA μ lp pp begin
            _check(ccall((:hmm_train_em_step, libhmmcuda), Cint,
                (Ptr{Cvoid}, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64,
                 Ptr{Float64}, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Float64}, Ptr{Cvoid}),
                ctx[], lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, s, lp, pp, ll, C_NULL))
        end
        lA_new = StateMatrix(lA.states .- one(Int16), pp, lA.K, lp[1:_nxi(lA)-1]; allow_overlaps=lA.resolve_overlaps)
        lA_new, μ, s[]
    end
    try
        return f(step)
    finally
        ccall((:hmm_train_destroy, libhmmcuda), Cint, (Ptr{Cvoid},), ctx[])
    end
end

# ---- n E/M steps in ONE library call (no callback between them) ------------------------------------------------
# The second loop of train_model (src/baumwelch.jl:351-353, `for i in 1:div(nsteps,2)`) and any caller-side loop with
# the default callback: X is uploaded once, every step's lp becomes the next step's transition weights inside the
# library (hmm_train_run), and the StateMatrix is rebuilt once at the end with the unchanged Julia constructor.
# Stops early if lp degenerates (a weight stops being finite) and finishes the remaining steps one by one.
function train_steps(X::Array{Float64,1}, lA::StateMatrix, μ::Array{Float64,2}, σ::Float64, nsteps::Integer)
    nsteps <= 0 && return lA, μ, σ
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve X _check(ccall((:hmm_train_create, libhmmcuda), Cint, (Ptr{Float64}, Int64, Ref{Ptr{Cvoid}}),
                                X, length(X), ctx))
    try
        left = Int(nsteps)
        while left > 0 && !isempty(lA)
            nlp = max(_nxi(lA) - 1, 1)
            lp = Vector{Float64}(undef, nlp); pp = Vector{Float64}(undef, lA.nstates)
            s = Ref{Float64}(σ); done = Ref{Int32}(0); tr = copy(lA.transitions)
            GC.@preserve lA μ lp pp tr begin
                _check(ccall((:hmm_train_run, libhmmcuda), Cint,
                    (Ptr{Cvoid}, Ptr{Int16}, Int32, Int32, Int32, Ptr{Cvoid}, Int64, Ptr{Float64}, Ref{Float64},
                     Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ref{Int32}, Ptr{Cvoid}),
                    ctx[], lA.states, lA.N, lA.K, lA.nstates, pointer(tr), length(tr), μ, s, lp, nlp, pp, C_NULL,
                    left, done, C_NULL))
            end
            σ = s[]
            lA = StateMatrix(lA.states .- one(Int16), pp, lA.K, lp[1:_nxi(lA)-1]; allow_overlaps=lA.resolve_overlaps)
            left -= max(Int(done[]), 1)
        end
    finally
        ccall((:hmm_train_destroy, libhmmcuda), Cint, (Ptr{Cvoid},), ctx[])
    end
    lA, μ, σ
end

# ---- optional: pinned host arrays ------------------------------------------------------------------------
# A Julia Array is pageable: long decodes then go through the library's pinned staging threads (≈2.9 Gsamples/s).
# Recordings that live in a pinned buffer reach the PCIe rate (≈5.6 Gsamples/s).  `pinned_vector(Float64, T)`
# returns an ordinary Vector backed by hmm_host_alloc memory; release it with `free_pinned(v)` when done.
function pinned_vector(::Type{T}, n::Integer) where T
    p = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:hmm_host_alloc, libhmmcuda), Cint, (Ref{Ptr{Cvoid}}, UInt64), p, UInt64(n) * sizeof(T)))
    unsafe_wrap(Array, convert(Ptr{T}, p[]), n; own=false)
end
free_pinned(v::Array) = (ccall((:hmm_host_free, libhmmcuda), Cint, (Ptr{Cvoid},), pointer(v)); nothing)

# ---- fit(HMMSpikingModel, templates, X, chunksize)        replaces the chunk loop of src/fit.jl:11-42
# The reference decodes 100k-sample chunks independently and stitches them at noise samples (approximate, and its
# `gc()` call no longer exists in Julia >= 1.0).  The GPU decode is exact and time-parallel inside one call, so the
# chunked method simply decodes the whole vector; `chunksize` is accepted and ignored.  Guard line for src/fit.jl:12:
#     HMMCuda.enabled() && return HMMSpikingModel(templates, HMMCuda.viterbi_whole(templates, X)..., X)
viterbi_whole(templates, X::AbstractVector{Float64}) = viterbi(X, templates.state_matrix, templates.μ, templates.σ)

# ---- unroll_mlseq(mlseq, state_matrix)                    replaces src/extraction.jl:4-13
function unroll_mlseq(mlseq::AbstractVector{Int16}, lA::StateMatrix)
    x = mlseq isa Vector{Int16} ? mlseq : collect(mlseq)
    out = Matrix{Int16}(undef, lA.N, length(x))
    GC.@preserve x lA out begin
        _check(ccall((:hmm_unroll_mlseq_i16, libhmmcuda), Cint,
            (Ptr{Int16}, Int64, Ptr{Int16}, Int32, Int32, Ptr{Int16}),
            x, length(x), lA.states, lA.N, lA.nstates, out))
    end
    out
end

# ---- devices ---------------------------------------------------------------------------------------------------
# set_devices([0,1,2,3]): subsequent viterbi / batch calls spread their work over these GPUs inside the library
# (channels of a batch round-robin; one long recording as time shards with peer-memory boundary exchange).
function set_devices(devs::AbstractVector{<:Integer})
    d = convert(Vector{Cint}, devs)
    GC.@preserve d _check(ccall((:hmm_set_devices, libhmmcuda), Cint, (Ptr{Cint}, Cint), d, length(d)))
end

end # module HMMCuda

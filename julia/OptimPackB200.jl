#
# OptimPackB200.jl --
#
# Julia host for libopkb200.so: a drop-in for the large-scale optimiser hot
# path of OptimPackNextGen.jl (`vmlmb`, `vmlmb!`, `spg`, `spg!`, `BoundedSet`,
# the line searches) with every n-vector resident on a B200.
#
# STATUS: syntax-reviewed only.  No Julia runtime exists in the build
# environment of this repository, so this file has never been executed; the
# tested twin of the same logic is the Python package next to the library
# (optimpacknextgen.jl_b200/vmlmb.py, spg.py).  Keep it thin: every vector
# operation is one `ccall`, the scalar stage machine is the reference's.
#
# Two ways to use it (INTEGRATION.md):
#
#  (A) plugin: `DeviceVector{T}` implements the vector-space contract of
#      doc/algebra.md:84-116 (LazyAlgebra `v*` methods + SimpleBounds), so the
#      UNMODIFIED `OptimPackNextGen.vmlmb!` / `spg!` run on device vectors, one
#      kernel per call;
#  (B) fused: `OptimPackB200.vmlmb` / `spg` keep the reference's keywords and
#      line-search objects and run three fused kernels per iteration.
#
module OptimPackB200

export DeviceVector, BoundedSet, Rosenbrock, Quadratic, Smooth2D, vmlmb, vmlmb!, vmlmb_native!, spg, spg!,
    spg_native!,
    conjgrad, conjgrad!, DiagonalOperator, Smooth2DOperator

using Libdl
using Printf
import LazyAlgebra
import LazyAlgebra: vcreate, vcopy!, vcopy, vdot, vnorm1, vnorm2, vnorminf, vscale!, vupdate!,
    vcombine!, vproduct!, vfill!, vswap!
import OptimPackNextGen
import OptimPackNextGen.SimpleBounds: project_variables!, project_direction!, step_limits,
    get_free_variables!
using OptimPackNextGen.LineSearches   # the reference's own line-search objects

const Float = Cdouble
const libopkb200 = get(ENV, "OPKB200_LIBRARY", "libopkb200.so")

# ---------------------------------------------------------------------------
# Status -> exception (src/wrappers.jl:110-155 style)
function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:opkb_last_error, libopkb200), Cstring, ()))
    rc == -5 && throw(ArgumentError("invalid bounds"))
    rc == -6 && throw(ArgumentError("invalid orientation"))
    error("libopkb200 error $rc: $msg")
end

eltype_code(::Type{Float32}) = Cint(0)   # OPKB_FLOAT
eltype_code(::Type{Float64}) = Cint(1)   # OPKB_DOUBLE

# ---------------------------------------------------------------------------
# Context and device vectors
mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer = -1)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:opkb_context_create, libopkb200), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device))
        ctx = new(ref[])
        finalizer(c -> ccall((:opkb_context_destroy, libopkb200), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        return ctx
    end
end

const DEFAULT_CONTEXT = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT_CONTEXT[] === nothing && (DEFAULT_CONTEXT[] = Context());
                     DEFAULT_CONTEXT[]::Context)

# Multi-GPU: rank 0 calls `unique_id()`, the host broadcasts it (MPI.jl, ...).
function unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:opkb_comm_unique_id, libopkb200), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(ctx::Context, id::Vector{UInt8}, rank::Integer, nranks::Integer) =
    check(ccall((:opkb_comm_init, libopkb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint),
                ctx.ptr, id, rank, nranks))

"""
    DeviceVector{T}(ctx, n)      # vcreate
    DeviceVector(ctx, a::Array)  # upload

An n-vector in the HBM of the context's GPU (this rank's shard).  It is an
`AbstractVector{T}` so that the reference's method signatures apply; scalar
indexing is deliberately not implemented.
"""
mutable struct DeviceVector{T<:Union{Float32,Float64}} <: AbstractVector{T}
    ptr::Ptr{Cvoid}
    ctx::Context
    len::Int
    owned::Bool
    function DeviceVector{T}(ctx::Context, n::Integer) where {T}
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:opkb_vector_create, libopkb200), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Ref{Ptr{Cvoid}}), ctx.ptr, eltype_code(T), n, ref))
        v = new{T}(ref[], ctx, n, true)
        finalizer(x -> x.owned && ccall((:opkb_vector_destroy, libopkb200), Cint, (Ptr{Cvoid},), x.ptr), v)
        return v
    end
    # borrowed handle (workspace vectors)
    DeviceVector{T}(ctx::Context, ptr::Ptr{Cvoid}, n::Integer) where {T} = new{T}(ptr, ctx, n, false)
end

Base.size(v::DeviceVector) = (v.len,)
Base.length(v::DeviceVector) = v.len
Base.getindex(::DeviceVector, i...) = error("scalar indexing of a DeviceVector is not supported")
Base.setindex!(::DeviceVector, x, i...) = error("scalar indexing of a DeviceVector is not supported")
Base.show(io::IO, v::DeviceVector{T}) where {T} = print(io, "DeviceVector{$T}(n=$(v.len))")
Base.show(io::IO, ::MIME"text/plain", v::DeviceVector) = show(io, v)

function upload!(v::DeviceVector{T}, a::Array{T}) where {T}
    length(a) == length(v) || throw(DimensionMismatch("upload!"))
    GC.@preserve a check(ccall((:opkb_vector_upload, libopkb200), Cint,
                               (Ptr{Cvoid}, Ptr{T}, Int64, Int64), v.ptr, a, 0, length(a)))
    return v
end
function download!(a::Array{T}, v::DeviceVector{T}) where {T}
    length(a) == length(v) || throw(DimensionMismatch("download!"))
    GC.@preserve a check(ccall((:opkb_vector_download, libopkb200), Cint,
                               (Ptr{Cvoid}, Ptr{T}, Int64, Int64), v.ptr, a, 0, length(a)))
    return a
end
DeviceVector(ctx::Context, a::Array{T}) where {T<:Union{Float32,Float64}} =
    upload!(DeviceVector{T}(ctx, length(a)), vec(a))
DeviceVector(a::Array) = DeviceVector(default_context(), a)
Base.Array(v::DeviceVector{T}) where {T} = download!(Vector{T}(undef, length(v)), v)

# ---------------------------------------------------------------------------
# (A) The vector-space contract (doc/algebra.md:84-116) on DeviceVector.
vcreate(x::DeviceVector{T}) where {T} = DeviceVector{T}(x.ctx, length(x))
function vcopy!(dst::DeviceVector{T}, src::DeviceVector{T}) where {T}
    check(ccall((:opkb_vcopy, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, src.ptr))
    return dst
end
vcopy(x::DeviceVector) = vcopy!(vcreate(x), x)
function vfill!(x::DeviceVector, alpha::Real)
    check(ccall((:opkb_vector_fill, libopkb200), Cint, (Ptr{Cvoid}, Cdouble), x.ptr, alpha))
    return x
end
function vdot(x::DeviceVector{T}, y::DeviceVector{T}) where {T}
    r = Ref{Cdouble}(0)
    check(ccall((:opkb_vdot, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}),
                x.ctx.ptr, x.ptr, y.ptr, r))
    return r[]
end
function vnorm2(x::DeviceVector)
    r = Ref{Cdouble}(0)
    check(ccall((:opkb_vnorm2, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}),
                x.ctx.ptr, x.ptr, r))
    return r[]
end
function vnorminf(x::DeviceVector)
    r = Ref{Cdouble}(0)
    check(ccall((:opkb_vnorminf, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}),
                x.ctx.ptr, x.ptr, r))
    return r[]
end
function vscale!(x::DeviceVector, alpha::Real)
    check(ccall((:opkb_vscale, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
                x.ctx.ptr, x.ptr, alpha))
    return x
end
function vupdate!(y::DeviceVector{T}, alpha::Real, x::DeviceVector{T}) where {T}
    check(ccall((:opkb_vupdate, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                y.ctx.ptr, y.ptr, alpha, x.ptr))
    return y
end
function vcombine!(dst::DeviceVector{T}, alpha::Real, x::DeviceVector{T},
                   beta::Real, y::DeviceVector{T}) where {T}
    check(ccall((:opkb_vcombine, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, alpha, x.ptr, beta, y.ptr))
    return dst
end
# the rest of the vector-space contract (doc/algebra.md:31, :50, :60-67, :116)
function vswap!(x::DeviceVector{T}, y::DeviceVector{T}) where {T}
    check(ccall((:opkb_vswap, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), x.ctx.ptr, x.ptr, y.ptr))
    return nothing
end
function vnorm1(x::DeviceVector)
    r = Ref{Cdouble}(0)
    check(ccall((:opkb_vnorm1, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}), x.ctx.ptr, x.ptr, r))
    return r[]
end
function vcombine!(dst::DeviceVector{T}, alpha::Real, x::DeviceVector{T}, beta::Real, y::DeviceVector{T},
                   gamma::Real, z::DeviceVector{T}) where {T}
    check(ccall((:opkb_vcombine3, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, alpha, x.ptr, beta, y.ptr, gamma, z.ptr))
    return dst
end

function vproduct!(dst::DeviceVector{T}, x::DeviceVector{T}, y::DeviceVector{T}) where {T}
    check(ccall((:opkb_vproduct, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, x.ptr, y.ptr))
    return dst
end

# The free set is a device mask, never an index list: `FreeMask` plays the role
# of `sel::Vector{Int}` (src/quasinewton.jl:324, src/bounds.jl:701-714).
struct FreeMask{T}
    w::DeviceVector{T}   # free variables are those with w[i] != 0
end
"""
    get_free_variables(gp::DeviceVector) -> FreeMask

A private copy of `gp` is kept because `_vmlmb!` overwrites the vector it took
the free set from (src/quasinewton.jl:528-530).
"""
get_free_variables(gp::DeviceVector) = FreeMask(vcopy(gp))
function vdot(sel::FreeMask{T}, x::DeviceVector{T}, y::DeviceVector{T}) where {T}
    r = Ref{Cdouble}(0)
    check(ccall((:opkb_vdot_masked, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}),
                x.ctx.ptr, sel.w.ptr, x.ptr, y.ptr, r))
    return r[]
end
function vupdate!(y::DeviceVector{T}, sel::FreeMask{T}, alpha::Real, x::DeviceVector{T}) where {T}
    check(ccall((:opkb_vupdate_masked, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
                y.ctx.ptr, y.ptr, sel.w.ptr, alpha, x.ptr))
    return y
end
function vproduct!(dst::DeviceVector{T}, sel::FreeMask{T}, x::DeviceVector{T}, y::DeviceVector{T}) where {T}
    check(ccall((:opkb_vproduct_masked, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, sel.w.ptr, x.ptr, y.ptr))
    return dst
end

# SimpleBounds on device vectors (src/bounds.jl:137-213, :321-409, :431-624).
const DeviceBound{T} = Union{Nothing,Real,DeviceVector{T}}
bound_args(::Nothing, default) = (C_NULL, Cdouble(default))
bound_args(b::Real, default) = (C_NULL, Cdouble(b))
bound_args(b::DeviceVector, default) = (b.ptr, Cdouble(default))
orient_code(o) = (o === (+) || (o isa Real && o > 0)) ? Cint(1) :
                 (o === (-) || (o isa Real && o < 0)) ? Cint(-1) :
                 throw(ArgumentError("invalid orientation"))

function project_variables!(dst::DeviceVector{T}, src::DeviceVector{T},
                            lo::DeviceBound{T}, hi::DeviceBound{T}) where {T}
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    check(ccall((:opkb_project_variables, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble),
                dst.ctx.ptr, dst.ptr, src.ptr, lp, lv, hp, hv))
    return dst
end
function project_direction!(dst::DeviceVector{T}, x::DeviceVector{T}, lo::DeviceBound{T},
                            hi::DeviceBound{T}, orient, d::DeviceVector{T}) where {T}
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    check(ccall((:opkb_project_direction, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Cint,
                 Ptr{Cvoid}),
                dst.ctx.ptr, dst.ptr, x.ptr, lp, lv, hp, hv, orient_code(orient), d.ptr))
    return dst
end
function step_limits(x::DeviceVector{T}, lo::DeviceBound{T}, hi::DeviceBound{T}, orient,
                     d::DeviceVector{T}) where {T}
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    smin, smax = Ref{Cdouble}(0), Ref{Cdouble}(0)
    check(ccall((:opkb_step_limits, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cvoid},
                 Ref{Cdouble}, Ref{Cdouble}),
                x.ctx.ptr, x.ptr, lp, lv, hp, hv, orient_code(orient), d.ptr, smin, smax))
    return (smin[], smax[])
end
# With these methods `OptimPackNextGen.spg!(fg!, prj!, x::DeviceVector, m)` runs
# unmodified.  `OptimPackNextGen.vmlmb!` additionally needs its private
# `vdot`/`vnorm2` shims (src/quasinewton.jl:45-69) and `sel = Array{Int}`
# (:324) generalised to `FreeMask`; INTEGRATION.md lists that three-line patch.

# ---------------------------------------------------------------------------
# (B) Fused path.
"""
    BoundedSet(lower, upper)

The box of the `lower=` / `upper=` keywords (src/quasinewton.jl:228-229); bounds
are scalars, host arrays (uploaded once) or `DeviceVector`s.
"""
struct BoundedSet{L,U}
    lower::L
    upper::U
end
BoundedSet(; lower = -Inf, upper = +Inf) = BoundedSet(lower, upper)

# Built-in objectives with fused f + gradient kernels.
struct Rosenbrock end                               # test/rosenbrock.jl:16-29
struct Quadratic{V}                                 # f = 1/2 sum a_i (x_i - c_i)^2
    a::V
    c::V
end
objective_code(::Rosenbrock) = (Cint(1), C_NULL, C_NULL)
objective_code(q::Quadratic{<:DeviceVector}) = (Cint(2), q.a.ptr, q.c.ptr)
objective_code(::Any) = (Cint(0), C_NULL, C_NULL)   # user fg!(x, g) on DeviceVectors

"""
    Smooth2D(w, y, mu, cols)

a non-separable `fg!(x, g)`: `f = 1/2 sum(w.*(x-y).^2) + mu/2 * sum over neighbour pairs
of (x_i - x_j)^2` on a rows x cols grid (`opkb_smooth2d_fg`).  With several ranks each
one holds a block of whole rows; the boundary rows are exchanged between neighbouring
ranks (ncclSend/ncclRecv) inside the call, overlapped with the inner rows.
"""
struct Smooth2D{T}
    w::DeviceVector{T}
    y::DeviceVector{T}
    mu::Float
    cols::Int
    up::DeviceVector{T}     # halo rows (unused on one rank)
    dn::DeviceVector{T}
end
Smooth2D(w::DeviceVector{T}, y::DeviceVector{T}, mu::Real, cols::Integer) where {T} =
    Smooth2D{T}(w, y, Float(mu), Int(cols), DeviceVector{T}(w.ctx, cols), DeviceVector{T}(w.ctx, cols))
function (obj::Smooth2D{T})(x::DeviceVector{T}, g::DeviceVector{T}) where {T}
    f = Ref{Cdouble}(0)
    check(ccall((:opkb_smooth2d_fg, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ptr{Cvoid}, Ref{Cdouble}),
                x.ctx.ptr, obj.w.ptr, obj.y.ptr, obj.mu, obj.cols, x.ptr, obj.up.ptr, obj.dn.ptr, g.ptr, f))
    return f[]
end

to_device(ctx, b::Real, ::Type{T}) where {T} = Float(b)
to_device(ctx, b::DeviceVector{T}, ::Type{T}) where {T} = b
to_device(ctx, b::Array, ::Type{T}) where {T} = DeviceVector(ctx, convert(Array{T}, b))

mutable struct Workspace{T}
    ptr::Ptr{Cvoid}
    ctx::Context
    n::Int
    mem::Int
    method::Int
    keep::Vector{Any}      # bounds / objective data that must outlive the workspace
end
function make_workspace(::Type{T}, ctx::Context, n::Int, mem::Int, method::Int, lo, hi) where {T}
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:opkb_vmlmb_create, libopkb200), Cint,
                (Ptr{Cvoid}, Cint, Int64, Cint, Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble,
                 Ref{Ptr{Cvoid}}),
                ctx.ptr, eltype_code(T), n, mem, method, lp, lv, hp, hv, ref))
    ws = Workspace{T}(ref[], ctx, n, mem, method, Any[lo, hi])
    finalizer(w -> ccall((:opkb_vmlmb_destroy, libopkb200), Cint, (Ptr{Cvoid},), w.ptr), ws)
    return ws
end
"""
    carry_stats(ws) -> (hits, misses)

how many Gram blocks of this workspace were assembled without a sweep over the pairs
(the incremental Gram of the fused direction + trial kernel, DESIGN.md section 4) and how
many sweeps ran although a carry was armed.  Diagnostics only: `vmlmb!` needs no change.
"""
function carry_stats(ws::Workspace)
    h, m = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:opkb_vmlmb_carry_stats, libopkb200), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}),
                ws.ptr, h, m))
    return (h[], m[])
end

"""
    chain_stats(ws) -> (heads, adopted, cancelled)

device-resident decisions of the native driver (`opkb200.h`, `opkb_vmlmb_chain_stats`): fused
launches that started a chain, launches adopted while already in flight, launches enqueued ahead
that the device cancelled.  Diagnostics only.
"""
function chain_stats(ws::Workspace)
    a, b, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:opkb_vmlmb_chain_stats, libopkb200), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}),
                ws.ptr, a, b, c))
    return (a[], b[], c[])
end

wsvector(ws::Workspace{T}, which::Char, slot::Integer = 0) where {T} =
    DeviceVector{T}(ws.ctx, ccall((:opkb_vmlmb_vector, libopkb200), Ptr{Cvoid},
                                  (Ptr{Cvoid}, Cint, Cint), ws.ptr, Cint(which), slot), ws.n)

sufficient_descent(gd, ε, gnorm, dnorm) = ε > 0 ? (gd ≤ -ε*dnorm*gnorm) : gd < 0

# Two-loop recursion in coefficient space from the Gram block of
# `opkb_vmlmb_gram` (pairs j = 1 oldest .. m newest).  G is (m+1) x 2m column
# major here, i.e. G[c, r] = row r, column c of the C row-major block.
function solve_two_loop(G::Matrix{Float}, m::Int, method::Int, rho_store::Vector{Float},
                        slots::Vector{Cint}, gamma_prev::Float)
    sv(a) = G[1, 2a-1]; yv(a) = G[1, 2a]
    SY(a, b) = G[1+b, 2a-1]
    YY(a, b) = b ≥ a ? G[1+b, 2a] : G[1+a, 2b]
    gamma_store = gamma_prev
    if method == 2
        rho = Float[SY(j, j) for j in 1:m]
    else
        rho_store[slots[m]+1] = SY(m, m)
        if rho_store[slots[m]+1] > 0
            gamma_store = rho_store[slots[m]+1]/YY(m, m)
        end
        rho = Float[rho_store[slots[j]+1] for j in 1:m]
    end
    alpha = fill(NaN, m); c = zeros(Float, m); gamma = Float(0)
    for k in m:-1:1
        if rho[k] > 0
            acc = sv(k)
            for j in m:-1:k+1
                rho[j] > 0 && (acc -= alpha[j]*SY(k, j))
            end
            alpha[k] = acc/rho[k]
            gamma == 0 && (gamma = rho[k]/YY(k, k))
        end
    end
    method < 2 && (gamma = gamma_store)
    change = any(r -> r > 0, rho) && gamma != 0
    if change
        for k in 1:m
            if rho[k] > 0
                acc = yv(k)
                for j in m:-1:1
                    rho[j] > 0 && (acc -= alpha[j]*YY(k, j))
                end
                acc *= gamma
                for j in 1:k-1
                    rho[j] > 0 && (acc += c[j]*SY(j, k))
                end
                c[k] = alpha[k] - acc/rho[k]
            end
        end
    end
    return alpha, c, gamma, gamma_store, change
end

"""
    vmlmb(fg!, x0; mem, lower, upper, blmvm, fmin, maxiter, maxeval, xtol, ftol,
          gtol, epsilon, verb, printer, output, lnsrch) -> x

Same keywords and defaults as `OptimPackNextGen.vmlmb` (src/quasinewton.jl:226-242).
`fg!` is `Rosenbrock()`, `Quadratic(a, c)` or any `fg!(x::DeviceVector, g::DeviceVector)`;
`x0` is a host `Array` (a host `Array` is returned) or a `DeviceVector`.
"""
vmlmb(fg!, x0::Array; kwds...) = Array(vmlmb!(fg!, DeviceVector(default_context(), x0); kwds...))
vmlmb(fg!, x0::DeviceVector; kwds...) = vmlmb!(fg!, vcopy(x0); kwds...)

function vmlmb!(fg!, x::DeviceVector{T};
                mem::Integer = min(5, length(x)),
                lower = -Inf, upper = +Inf,
                autodiff::Bool = false,
                blmvm::Bool = false,
                fmin::Real = -Inf,
                maxiter::Integer = typemax(Int),
                maxeval::Integer = typemax(Int),
                xtol::NTuple{2,Real} = (0.0, 1e-7),
                ftol::NTuple{2,Real} = (0.0, 1e-8),
                gtol::NTuple{2,Real} = (0.0, 1e-6),
                epsilon::Real = 0.0,
                verb::Integer = false,
                printer::Function = OptimPackNextGen.QuasiNewton.print_iteration,
                output::IO = stdout,
                lnsrch::Union{LineSearch{Float},Nothing} = nothing,
                nglobal::Integer = length(x)) where {T}
    # (src/quasinewton.jl:230: the keyword exists for drop-in calls; automatic differentiation of a host
    # closure has no meaning on device vectors -- SURVEY section 2, out of scope)
    autodiff && throw(ArgumentError("autodiff=true is not supported on device vectors: provide fg!(x, g)"))
    if lower isa BoundedSet
        lower, upper = lower.lower, lower.upper
    end
    if T === Float32 || min(mem, nglobal) > MEM_MAX_GRAM
        # The host loop below solves the two-loop recursion in coefficient space from one Gram
        # sweep: Float64 and at most 20 pairs.  Float32 data (whose parity-grade form keeps the
        # roundings of the intermediate vector) and larger memories (the reference accepts any
        # `mem`, src/quasinewton.jl:227, :328) run the step-by-step recursion of the native driver.
        return vmlmb_native!(fg!, x; mem=mem, lower=lower, upper=upper, blmvm=blmvm, fmin=fmin,
                             maxiter=maxiter, maxeval=maxeval, xtol=xtol, ftol=ftol, gtol=gtol,
                             epsilon=epsilon, verb=verb, printer=printer, output=output, lnsrch=lnsrch,
                             nglobal=nglobal)
    end
    ctx = x.ctx
    lo = to_device(ctx, lower, T); hi = to_device(ctx, upper, T)
    bounds = 0
    (lo isa DeviceVector || lo > -Inf) && (bounds |= 1)
    (hi isa DeviceVector || hi < +Inf) && (bounds |= 2)
    method = (bounds == 0 ? 0 : blmvm ? 1 : 2)                       # :273
    ls = lnsrch !== nothing ? lnsrch :                                # :276-284
         method == 0 ? MoreThuenteLineSearch(Float; ftol=1e-3, gtol=0.9, xtol=0.1) :
                       MoreToraldoLineSearch(Float; ftol=1e-3, gamma=(0.1,0.5))
    mem = Int(min(mem, nglobal))
    n = length(x)
    ws = make_workspace(T, ctx, n, mem, method, lo, hi)
    kind, pa, pc = objective_code(fg!)
    fg! isa Quadratic && push!(ws.keep, fg!)
    check(ccall((:opkb_vmlmb_set_objective, libopkb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                ws.ptr, kind, pa, pc))
    xw = wsvector(ws, 'x'); gw = wsvector(ws, 'g')
    vcopy!(xw, x)

    xatol, xrtol = Float.(xtol); fatol, frtol = Float.(ftol); gatol, grtol = Float.(gtol)
    fmin = Float(fmin); epsilon = Float(epsilon)
    fminset = (!isnan(fmin) && fmin > -Inf)
    STPMIN, STPMAX = Float(1e-20), Float(1e+20)
    mark = 0; m = 0; iter = 0; eval = 0; rejects = 0; stage = 0
    f = f0 = gd = gd0 = stp = gnorm = gtest = Float(0); gamma = Float(1)
    beststp = bestf = bestgnorm = Float(0); smin = smax = Float(Inf)
    reason = ""
    rho = zeros(Float, mem)
    out8 = zeros(Float, 8); out4 = zeros(Float, 4)
    saved_for = -1; need_steepest = false

    while true
        # P1: x = prj(S[mark] - stp*d); f = fg!(x, g); p = projected gradient; 6 reductions
        initial = Cint(eval == 0)
        if kind != 0
            check(ccall((:opkb_vmlmb_trial, libopkb200), Cint,
                        (Ptr{Cvoid}, Cint, Cdouble, Cint, Ptr{Cdouble}), ws.ptr, mark, stp, initial, out8))
        else
            check(ccall((:opkb_vmlmb_trial_begin, libopkb200), Cint,
                        (Ptr{Cvoid}, Cint, Cdouble, Cint), ws.ptr, mark, stp, initial))
            fx = Float(fg!(xw, gw))
            check(ccall((:opkb_vmlmb_trial_end, libopkb200), Cint,
                        (Ptr{Cvoid}, Cint, Cdouble, Cint, Ptr{Cdouble}), ws.ptr, mark, fx, initial, out8))
        end
        f = out8[1]; eval += 1
        if eval == 1 || f < bestf                                     # :398-419
            gnorm = sqrt(out8[2]); beststp = stp; bestf = f; bestgnorm = gnorm
            eval == 1 && (gtest = max(gatol, grtol*gnorm))
            if gnorm ≤ gtest
                stage = 3
                reason = gnorm == 0 ? "a stationary point has been found" :
                         method > 0 ? "projected gradient norm sufficiently small" :
                                      "gradient norm sufficiently small"
                eval > 1 && (iter += 1)
            end
        end
        if fminset && f < fmin
            stage = 4; reason = "f < fmin"
        end
        if stage == 1                                                 # :425-474
            if usederivatives(ls)
                gd = method > 0 ? out8[3]/stp : -out8[4]
            end
            task = iterate!(ls, stp, f, gd)
            if task == :SEARCH
                stp = getstep(ls)
            elseif task == :CONVERGENCE
                iter += 1
                xtest = max(xatol, zero(Float))
                xrtol > 0 && (xtest = max(xtest, xrtol*sqrt(out8[5])))
                if sqrt(out8[6]) ≤ xtest
                    stage = 3; reason = "X test satisfied"
                else
                    delta = max(abs(f - f0), stp*abs(gd0))
                    if delta ≤ fatol
                        stage = 3; reason = "fatol test satisfied"
                    elseif delta ≤ frtol*abs(f0)
                        stage = 3; reason = "frtol test satisfied"
                    elseif iter ≥ maxiter
                        stage = 4; reason = "too many iterations"
                    elseif eval ≥ maxeval
                        stage = 4; reason = "too many evaluations"
                    else
                        stage = 2
                    end
                end
            else
                stage = 4; reason = getreason(ls)
            end
        end
        if stage < 2 && eval ≥ maxeval
            stage = 4; reason = "too many evaluations"
        end
        if stage != 1                                                 # :480-596
            if stp != beststp
                if stage ≥ 3
                    stp = beststp; f = bestf; gnorm = bestgnorm
                    check(ccall((:opkb_vmlmb_revert, libopkb200), Cint, (Ptr{Cvoid}, Cint, Cdouble),
                                ws.ptr, mark, stp))
                else
                    gnorm = sqrt(out8[2])
                end
            end
            if verb > 0 && iter % verb == 0
                printer(output, iter, eval, rejects, f, gnorm, stp)
            end
            stage ≥ 3 && break
            if stage == 2
                # P3: pair update + Gram sweep; host solve; P4: direction
                m = min(m + 1, mem)
                slots = Cint[mod(mark - (m - 1) + j, mem) for j in 0:m-1]
                G = zeros(Float, m + 1, 2m)
                check(ccall((:opkb_vmlmb_gram, libopkb200), Cint,
                            (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}), ws.ptr, mark, m, slots, G))
                alpha, c, gam, gamma, change = solve_two_loop(G, m, method, rho, slots, gamma)
                saved_for = -1
                reject = true
                if change
                    mark_next = mark < mem - 1 ? mark + 1 : 0
                    check(ccall((:opkb_vmlmb_direction, libopkb200), Cint,
                                (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint,
                                 Ptr{Cdouble}), ws.ptr, m, slots, alpha, c, gam, mark_next, out4))
                    gd = -out4[1]
                    smin, smax = out4[3], out4[4]
                    saved_for = mark_next; need_steepest = false
                    reject = !sufficient_descent(gd, epsilon, gnorm, sqrt(out4[2]))
                end
                if reject
                    stage = 0; rejects += 1
                end
                mark = mark < mem - 1 ? mark + 1 : 0
            end
            if stage == 0
                need_steepest = true; gd = -gnorm^2
            end
            f0 = f; gd0 = gd
            if need_steepest || saved_for != mark                     # P4': d = p, save x, g, limits
                check(ccall((:opkb_vmlmb_steepest, libopkb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}),
                            ws.ptr, mark, out4))
                smin, smax = out4[3], out4[4]; need_steepest = false
            end
            saved_for = -1
            stp = stage == 2 ? one(Float) :
                  (fminset && fmin < f0) ? 2*(fmin - f0)/gd0 : 1/gnorm
            if method > 0
                stp = min(stp, smax); stpmin = STPMIN*stp; stpmax = min(STPMAX*stp, smax)
            else
                stpmin = STPMIN*stp; stpmax = STPMAX*stp
            end
            task = start!(ls, f0, gd0, stp; stpmin=stpmin, stpmax=stpmax)
            if task != :SEARCH
                stage = 4; reason = getreason(ls); break
            end
            stage = 1
        end
    end
    if verb > 0
        printstyled(output, "# ", (stage > 3 ? "WARNING: " : "CONVERGENCE: "), reason, "\n";
                    color = (stage > 3 ? :red : :green))
    end
    vcopy!(x, xw)
    return x
end

# ---------------------------------------------------------------------------
# (C) The same solver with the driver loop inside the library (opkb_solver_*):
# reverse communication in the style of `opk_iterate_vmlmb`
# (src/wrappers.jl:1831-1859) -- one ccall per objective evaluation, or a single
# ccall for a built-in objective.
struct NativeOptions          # opkb_vmlmb_options of include/opkb200.h (same layout)
    blmvm::Cint
    fmin::Cdouble
    maxiter::Int64
    maxeval::Int64
    xatol::Cdouble; xrtol::Cdouble; fatol::Cdouble; frtol::Cdouble; gatol::Cdouble; grtol::Cdouble
    epsilon::Cdouble
    lnsrch::Cint              # 0 default, 1 Armijo, 2 More & Toraldo, 3 More & Thuente
    ls_ftol::Cdouble; ls_gtol::Cdouble; ls_xtol::Cdouble; ls_gamma1::Cdouble; ls_gamma2::Cdouble
    ls_strict::Cint
    twoloop::Cint
    speculate::Cint
end

const TASK_COMPUTE_FG = Cint(1)
const MEM_MAX_GRAM = 20       # pairs the Gram-form kernels hold (OPKB_MEM_MAX); beyond: sequential two-loop

# options of the reference's line-search objects -> opkb_vmlmb_options (OPKB_LS_*)
function native_linesearch(ls)
    ls === nothing && return (Cint(0), 1e-3, 0.9, 0.1, 0.1, 0.5, Cint(0))
    if ls isa MoreThuenteLineSearch
        return (Cint(3), Float(ls.ftol), Float(ls.gtol), Float(ls.xtol), 0.1, 0.5, Cint(0))
    elseif ls isa MoreToraldoLineSearch
        return (Cint(2), Float(ls.ftol), 0.9, 0.1, Float(ls.gamma1), Float(ls.gamma2), Cint(ls.strict))
    elseif ls isa ArmijoLineSearch
        return (Cint(1), Float(ls.ftol), 0.9, 0.1, 0.1, 0.5, Cint(0))
    end
    throw(ArgumentError("the native driver knows the three line searches of the reference"))
end

# Device-callback objective (opkb_solver_set_fg_callback): the reverse-communication loop runs
# inside the library and calls back into Julia; style of src/cobyla.jl:184-193 (`@cfunction` +
# `unsafe_pointer_to_objref`).  The raw pointers are those of the workspace vectors `x`, `g`
# (whose HANDLES are stable while their data pointers move): `fg!` is called on the handles.
mutable struct CallbackState{F,V}
    fg::F
    x::V
    g::V
    err::Any
end

function fg_trampoline(user::Ptr{Cvoid}, n::Int64, xp::Ptr{Cvoid}, gp::Ptr{Cvoid}, stream::Ptr{Cvoid},
                       fout::Ptr{Cdouble})::Cint
    st = unsafe_pointer_to_objref(user)
    try
        unsafe_store!(fout, Float(st.fg(st.x, st.g)))
        return Cint(0)
    catch err
        st.err = err
        return Cint(1)
    end
end

function solver_status(solver)
    iter = Ref{Int64}(0); eval = Ref{Int64}(0); rej = Ref{Int64}(0)
    f = Ref{Cdouble}(0); gnorm = Ref{Cdouble}(0); stp = Ref{Cdouble}(0)
    stage = Ref{Cint}(0); newx = Ref{Cint}(0)
    check(ccall((:opkb_solver_status, libopkb200), Cint,
                (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble},
                 Ref{Cint}, Ref{Cint}), solver, iter, eval, rej, f, gnorm, stp, stage, newx))
    return (iter[], eval[], rej[], f[], gnorm[], stp[], stage[], newx[] != 0)
end

"""
    vmlmb_native!(fg!, x; kwds...) -> x

`vmlmb!` with the driver loop inside libopkb200 (`opkb_solver_*`): the keywords of
`vmlmb!` (src/quasinewton.jl:227-242) plus

* `twoloop = :default | :gram | :sequential` -- coefficient-space two-loop from one Gram
  sweep (Float64 default, `mem ≤ 20`) or the reference's step-by-step recursion
  (parity-grade; default for Float32 data and for `mem > 20`);
* `callback = true` -- a user `fg!` is installed as a device callback
  (`opkb_solver_set_fg_callback`): one `ccall` for the whole run.

`printer` is called at every iteration boundary when `verb > 0` (the run then
advances one iteration per `ccall`).
"""
function vmlmb_native!(fg!, x::DeviceVector{T}; mem::Integer = min(5, length(x)),
                       lower = -Inf, upper = +Inf, autodiff::Bool = false, blmvm::Bool = false, fmin::Real = -Inf,
                       maxiter::Integer = typemax(Int), maxeval::Integer = typemax(Int),
                       xtol::NTuple{2,Real} = (0.0, 1e-7), ftol::NTuple{2,Real} = (0.0, 1e-8),
                       gtol::NTuple{2,Real} = (0.0, 1e-6), epsilon::Real = 0.0,
                       verb::Integer = false,
                       printer::Function = OptimPackNextGen.QuasiNewton.print_iteration,
                       output::IO = stdout,
                       lnsrch::Union{LineSearch{Float},Nothing} = nothing,
                       twoloop::Symbol = :default, callback::Bool = false, chain::Bool = false,
                       nglobal::Integer = length(x)) where {T}
    autodiff && throw(ArgumentError("autodiff=true is not supported on device vectors: provide fg!(x, g)"))
    if lower isa BoundedSet
        lower, upper = lower.lower, lower.upper
    end
    ctx = x.ctx
    lo = to_device(ctx, lower, T); hi = to_device(ctx, upper, T)
    bounded = (lo isa DeviceVector || lo > -Inf) || (hi isa DeviceVector || hi < +Inf)
    method = !bounded ? 0 : blmvm ? 1 : 2
    ws = make_workspace(T, ctx, length(x), Int(min(mem, nglobal)), method, lo, hi)
    kind, pa, pc = objective_code(fg!)
    fg! isa Quadratic && push!(ws.keep, fg!)
    check(ccall((:opkb_vmlmb_set_objective, libopkb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                ws.ptr, kind, pa, pc))
    xw = wsvector(ws, 'x'); gw = wsvector(ws, 'g')
    vcopy!(xw, x)
    lscode, lsftol, lsgtol, lsxtol, lsg1, lsg2, lsstrict = native_linesearch(lnsrch)
    tl = twoloop === :default ? Cint(-1) : twoloop === :gram ? Cint(0) : twoloop === :sequential ? Cint(1) :
         throw(ArgumentError("twoloop must be :default, :gram or :sequential"))
    opt = Ref(NativeOptions(blmvm, fmin, maxiter, maxeval, xtol[1], xtol[2], ftol[1], ftol[2],
                            gtol[1], gtol[2], epsilon, lscode, lsftol, lsgtol, lsxtol, lsg1, lsg2, lsstrict,
                            tl, 1))
    sref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:opkb_solver_create, libopkb200), Cint, (Ptr{Cvoid}, Ref{NativeOptions}, Ref{Ptr{Cvoid}}),
                ws.ptr, opt, sref))
    solver = sref[]
    # device-resident decisions of the run loop (same results; opt-in, DESIGN.md "Chained launches")
    chain && check(ccall((:opkb_solver_set_chain, libopkb200), Cint, (Ptr{Cvoid}, Cint), solver, 1))
    task = Ref{Cint}(0)
    report() = begin
        iter, eval, rej, f, gnorm, stp, stage, newx = solver_status(solver)
        (newx && verb > 0 && iter % verb == 0) && printer(output, iter, eval, rej, f, gnorm, stp)
        stage
    end
    state = CallbackState(fg!, xw, gw, nothing)
    try
        if kind != 0 || callback
            if kind == 0      # user objective as a device callback: the loop stays inside the library
                cfg = @cfunction(fg_trampoline, Cint,
                                 (Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}))
                check(ccall((:opkb_solver_set_fg_callback, libopkb200), Cint,
                            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint), solver, cfg,
                            pointer_from_objref(state), Cint(0)))
            end
            GC.@preserve state begin
                step = verb > 0 ? Int64(1) : typemax(Int64) ÷ 2
                while true    # the whole run in one call, or one iteration per call when printing
                    rc = ccall((:opkb_solver_run, libopkb200), Cint, (Ptr{Cvoid}, Int64, Ref{Cint}),
                               solver, step, task)
                    state.err === nothing || throw(state.err)
                    check(rc)
                    report()
                    task[] == TASK_COMPUTE_FG || break
                end
            end
        else                  # reverse communication, one ccall per evaluation
            check(ccall((:opkb_solver_start, libopkb200), Cint, (Ptr{Cvoid}, Ref{Cint}), solver, task))
            while task[] == TASK_COMPUTE_FG
                fx = Float(fg!(xw, gw))
                check(ccall((:opkb_solver_iterate, libopkb200), Cint, (Ptr{Cvoid}, Cdouble, Ref{Cint}),
                            solver, fx, task))
                report()
            end
        end
        if verb > 0
            stage = solver_status(solver)[7]
            reason = unsafe_string(ccall((:opkb_solver_reason, libopkb200), Cstring, (Ptr{Cvoid},), solver))
            printstyled(output, "# ", (stage > 3 ? "WARNING: " : "CONVERGENCE: "), reason, "\n";
                        color = (stage > 3 ? :red : :green))
        end
        vcopy!(x, xw)
    finally
        ccall((:opkb_solver_destroy, libopkb200), Cint, (Ptr{Cvoid},), solver)
    end
    return x
end

# SPG with the box projection and a built-in objective: one kernel per function
# evaluation (opkb_spg_start / _step / _backtrack); any other `prj!` / `fg!` runs
# through `OptimPackNextGen.spg!` on DeviceVectors, path (A).
spg(fg!, prj!, x0::Array, m::Integer; kwds...) =
    Array(spg!(fg!, prj!, DeviceVector(default_context(), x0), m; kwds...))
spg(fg!, prj!, x0::DeviceVector, m::Integer; kwds...) = spg!(fg!, prj!, vcopy(x0), m; kwds...)

function spg!(fg!, prj!, x::DeviceVector{T}, m::Integer; kwds...) where {T}
    if !(prj! isa BoundedSet && (fg! isa Rosenbrock || fg! isa Quadratic))
        return OptimPackNextGen.spg!(fg!, prj!, x, m; kwds...)
    end
    return fused_spg!(fg!, prj!, x, Int(m); kwds...)
end

function fused_spg!(fg!, box::BoundedSet, x::DeviceVector{T}, m::Int;
                    autodiff::Bool = false,
                    ws::OptimPackNextGen.SPG.Info = OptimPackNextGen.SPG.Info(),
                    maxit::Integer = typemax(Int), maxfc::Integer = typemax(Int),
                    eps1::Real = 1e-6, eps2::Real = 1e-6, eta::Real = 1.0,
                    printer::Function = OptimPackNextGen.SPG.default_printer,
                    verb::Integer = false, io::IO = stdout) where {T}
    autodiff && throw(ArgumentError("autodiff=true is not supported on device vectors: provide fg!(x, g)"))
    ctx = x.ctx
    lo = to_device(ctx, box.lower, T); hi = to_device(ctx, box.upper, T)
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:opkb_spg_create, libopkb200), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ref{Ptr{Cvoid}}),
                ctx.ptr, eltype_code(T), length(x), lp, lv, hp, hv, ref))
    h = ref[]
    spgvec(which::Char) = DeviceVector{T}(ctx, ccall((:opkb_spg_vector, libopkb200), Ptr{Cvoid},
                                                     (Ptr{Cvoid}, Cint), h, Cint(which)), length(x))
    try
        kind, pa, pc = objective_code(fg!)
        check(ccall((:opkb_spg_set_objective, libopkb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                    h, kind, pa, pc))
        vcopy!(spgvec('x'), x)
        lmin, lmax, ftol, amin, amax = 1e-30, 1e+30, 1e-4, 0.1, 0.9   # src/spg.jl:201-205
        out = zeros(Float, 8)
        check(ccall((:opkb_spg_start, libopkb200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), h, eta, out))
        f = out[1]; fbest = f; iter = 0; fcnt = 1; pcnt = 1; status = 0
        lastfv = fill(-Inf, m); pgtwon = pginfn = Inf
        while true
            if m > 1
                lastfv[(iter % m) + 1] = f
                fmin, fmax = extrema(lastfv); fconst = !(fmin < fmax)
            else
                fmin = fmax = f; fconst = false
            end
            pgtwon = sqrt(out[3]); pginfn = out[4]; pcnt += 1
            if verb > 0 && iter % verb == 0
                ws.f = f; ws.fbest = fbest; ws.pgtwon = pgtwon; ws.pginfn = pginfn
                ws.iter = iter; ws.fcnt = fcnt; ws.pcnt = pcnt; ws.status = status
                printer(io, ws)
            end
            pginfn ≤ eps1 && (status = 1; break)
            pgtwon ≤ eps2 && (status = 2; break)
            fconst && (status = 3; break)
            iter ≥ maxit && (status = -1; break)
            fcnt ≥ maxfc && (status = -2; break)
            lambda = iter == 0 ? clamp(1/pginfn, lmin, lmax) :
                     out[5] > 0 ? clamp(out[6]/out[5], lmin, lmax) : lmax
            f0 = f
            check(ccall((:opkb_spg_step, libopkb200), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}),
                        h, lambda, eta, out))
            f = out[1]; delta = out[2]; pcnt += 1; stp = 1.0
            while true
                fcnt += 1
                if f < fbest
                    fbest = f
                    check(ccall((:opkb_spg_mark_best, libopkb200), Cint, (Ptr{Cvoid},), h))
                end
                f ≤ fmax + stp*ftol*delta && break
                fcnt ≥ maxfc && (status = -2; break)
                q = -delta*(stp*stp); r = (f - f0 - stp*delta)*2
                stp = (r > 0 && amin*r ≤ q ≤ amax*stp*r) ? q/r : stp/2
                check(ccall((:opkb_spg_backtrack, libopkb200), Cint,
                            (Ptr{Cvoid}, Cdouble, Cdouble, Cdouble, Ptr{Cdouble}), h, lambda, stp, eta, out))
                f = out[1]
            end
            status != 0 && break
            iter += 1
        end
        ws.f = fbest; ws.fbest = fbest; ws.pgtwon = pgtwon; ws.pginfn = pginfn
        ws.iter = iter; ws.fcnt = fcnt; ws.pcnt = pcnt; ws.status = status
        vcopy!(x, spgvec(fbest < f ? 'b' : 'x'))
    finally
        ccall((:opkb_spg_destroy, libopkb200), Cint, (Ptr{Cvoid},), h)
    end
    return x
end

# ---------------------------------------------------------------------------
# spg! with the loop inside the library (`opkb_spg_solver_*`, Tier 3): built-in objective + box.
function spg_native!(fg!, box::BoundedSet, x::DeviceVector{T}, m::Integer;
                     ws::OptimPackNextGen.SPG.Info = OptimPackNextGen.SPG.Info(),
                     maxit::Integer = typemax(Int), maxfc::Integer = typemax(Int),
                     eps1::Real = 1e-6, eps2::Real = 1e-6, eta::Real = 1.0) where {T}
    ctx = x.ctx
    lo = to_device(ctx, box.lower, T); hi = to_device(ctx, box.upper, T)
    (lp, lv), (hp, hv) = bound_args(lo, -Inf), bound_args(hi, +Inf)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:opkb_spg_create, libopkb200), Cint,
                (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cdouble, Ref{Ptr{Cvoid}}),
                ctx.ptr, eltype_code(T), length(x), lp, lv, hp, hv, ref))
    h = ref[]
    sref = Ref{Ptr{Cvoid}}(C_NULL)
    spgvec(which::Char) = DeviceVector{T}(ctx, ccall((:opkb_spg_vector, libopkb200), Ptr{Cvoid},
                                                     (Ptr{Cvoid}, Cint), h, Cint(which)), length(x))
    try
        kind, pa, pc = objective_code(fg!)
        check(ccall((:opkb_spg_set_objective, libopkb200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                    h, kind, pa, pc))
        vcopy!(spgvec('x'), x)
        check(ccall((:opkb_spg_solver_create, libopkb200), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Int64, Cdouble, Cdouble, Cdouble, Ref{Ptr{Cvoid}}),
                    h, m, maxit, maxfc, eps1, eps2, eta, sref))
        alive = Ref{Cint}(1)
        check(ccall((:opkb_spg_solver_run, libopkb200), Cint, (Ptr{Cvoid}, Int64, Ref{Cint}),
                    sref[], typemax(Int64), alive))
        f = Ref{Cdouble}(0); fb = Ref{Cdouble}(0); pi = Ref{Cdouble}(0); p2 = Ref{Cdouble}(0)
        it = Ref{Int64}(0); fc = Ref{Int64}(0); pc_ = Ref{Int64}(0); st = Ref{Cint}(0); bx = Ref{Cint}(1)
        check(ccall((:opkb_spg_solver_info, libopkb200), Cint,
                    (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}, Ref{Int64}, Ref{Int64},
                     Ref{Int64}, Ref{Cint}, Ref{Cint}), sref[], f, fb, pi, p2, it, fc, pc_, st, bx))
        ws.f = fb[]; ws.fbest = fb[]; ws.pginfn = pi[]; ws.pgtwon = p2[]
        ws.iter = it[]; ws.fcnt = fc[]; ws.pcnt = pc_[]; ws.status = st[]
        vcopy!(x, spgvec(bx[] != 0 ? 'x' : 'b'))
    finally
        sref[] != C_NULL && ccall((:opkb_spg_solver_destroy, libopkb200), Cint, (Ptr{Cvoid},), sref[])
        ccall((:opkb_spg_destroy, libopkb200), Cint, (Ptr{Cvoid},), h)
    end
    return x
end

# ---------------------------------------------------------------------------
# conjgrad / conjgrad! (re-exported by the reference from LazyAlgebra:
# src/OptimPackNextGen.jl:18-19, :36; call site test/conjgrad-tests.jl:30).  With plan (A)
# LazyAlgebra's own `conjgrad!` already runs on `DeviceVector`s, one kernel per call.  This
# is plan (B): the same loop with the update fused (`opkb_cg_update`: x += alpha p,
# r -= alpha q, r.r and x.x in one pass) and the curvature p.q riding on the operator
# when the operator is built in.

"`A = diag(a)` (`vproduct!`): direction + product + curvature in one kernel."
struct DiagonalOperator{T}
    a::DeviceVector{T}
end
(A::DiagonalOperator{T})(dst::DeviceVector{T}, src::DeviceVector{T}) where {T} = vproduct!(dst, A.a, src)

"`A = diag(w) + mu*L` on a rows x cols grid: the Hessian of `Smooth2D` (`opkb_smooth2d_apply`)."
struct Smooth2DOperator{T}
    w::DeviceVector{T}
    mu::Float
    cols::Int
    up::DeviceVector{T}
    dn::DeviceVector{T}
end
Smooth2DOperator(w::DeviceVector{T}, mu::Real, cols::Integer) where {T} =
    Smooth2DOperator{T}(w, Float(mu), Int(cols), DeviceVector{T}(w.ctx, cols), DeviceVector{T}(w.ctx, cols))
function apply_dot!(A::Smooth2DOperator{T}, q::DeviceVector{T}, p::DeviceVector{T}) where {T}
    out = Vector{Cdouble}(undef, 2)
    check(ccall((:opkb_smooth2d_apply, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ptr{Cdouble}),
                p.ctx.ptr, A.w.ptr, A.mu, A.cols, p.ptr, A.up.ptr, A.dn.ptr, q.ptr, out))
    return out[1], out[2]
end
(A::Smooth2DOperator{T})(dst::DeviceVector{T}, src::DeviceVector{T}) where {T} = (apply_dot!(A, dst, src); dst)

# direction (p = r | beta*p + r), q = A.p -> (p.q, p.p)
function cg_direction!(A::DiagonalOperator{T}, p, r, q, beta::Float, first::Bool) where {T}
    out = Vector{Cdouble}(undef, 2)
    check(ccall((:opkb_cg_diag_step, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cdouble}),
                p.ctx.ptr, A.a.ptr, p.ptr, r.ptr, q.ptr, beta, first, out))
    return out[1], out[2]
end
# direction + operator + p.q + p.p in ONE pass (`opkb_cg_smooth2d_step`): the new direction is
# written to a scratch vector and the library swaps the two buffers; with several ranks only
# the boundary rows of r travel, `pup` / `pdn` keep the neighbours' rows of the direction
const CG_SCRATCH = IdDict{Any,Any}()
function cg_direction!(A::Smooth2DOperator{T}, p, r, q, beta::Float, first::Bool) where {T}
    sc = get!(CG_SCRATCH, A) do
        (vcreate(p), ntuple(_ -> DeviceVector{T}(p.ctx, A.cols), 4))
    end
    scratch, (pup, pdn, rup, rdn) = sc
    out = Vector{Cdouble}(undef, 2)
    check(ccall((:opkb_cg_smooth2d_step, libopkb200), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid},
                 Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Cdouble}),
                p.ctx.ptr, A.w.ptr, A.mu, A.cols, p.ptr, scratch.ptr, r.ptr, q.ptr, pup.ptr, pdn.ptr,
                rup.ptr, rdn.ptr, beta, first, out))
    return out[1], out[2]
end
function cg_direction!(A, p, r, q, beta::Float, first::Bool)      # any `A!(dst, src)`
    first ? vcopy!(p, r) : vcombine!(p, beta, p, 1, r)
    A(q, p)
    out = Vector{Cdouble}(undef, 2)
    check(ccall((:opkb_cg_curvature, libopkb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}),
                p.ctx.ptr, p.ptr, q.ptr, out))
    return out[1], out[2]
end

function conjgrad!(x::DeviceVector{T}, A, b::DeviceVector{T}, x0::DeviceVector{T} = vfill!(x, 0),
                   p::DeviceVector{T} = vcreate(x), q::DeviceVector{T} = vcreate(x),
                   r::DeviceVector{T} = vcreate(x);
                   ftol::Real = 1e-8, gtol::NTuple{2,Real} = (0.0, 0.0), xtol::Real = 0.0,
                   maxiter::Integer = typemax(Int), restart::Integer = min(50, length(b)),
                   verb::Bool = false, io::IO = stdout, quiet::Bool = false,
                   strict::Bool = true) where {T}
    (ftol ≥ 0 && gtol[1] ≥ 0 && gtol[2] ≥ 0 && xtol ≥ 0) || throw(ArgumentError("bad tolerance"))
    restart ≥ 1 || throw(ArgumentError("restart must be at least 1"))
    x0 === x || vcopy!(x, x0)
    if vdot(x, x) > 0
        A(r, x); vcombine!(r, 1, b, -1, r)
    else
        vcopy!(r, b)
    end
    rho::Float = vdot(r, r)
    oldrho::Float = 0; phi::Float = 0; psimax::Float = 0
    gtest = max(Float(gtol[1]), Float(gtol[2])*sqrt(rho))
    out = Vector{Cdouble}(undef, 2)
    k = 0
    while true
        verb && @printf(io, " %6d %12.4e %12.4e\n", k, phi, sqrt(rho))
        sqrt(rho) ≤ gtest && break
        if k ≥ maxiter
            quiet || @warn "Too many iteration(s)"
            break
        end
        first = (rem(k, restart) == 0)
        if first && k > 0
            A(r, x); vcombine!(r, 1, b, -1, r)
        end
        gamma, pp = cg_direction!(A, p, r, q, first ? 0.0 : rho/oldrho, first)
        if !(gamma > 0)
            strict && error("Operator is not positive definite")
            quiet || @warn "Operator is not positive definite"
            break
        end
        alpha = rho/gamma
        psi = alpha*rho/2
        phi -= psi
        psimax = max(psi, psimax)
        if psi ≤ ftol*psimax
            vupdate!(x, alpha, p)
            break
        end
        check(ccall((:opkb_cg_update, libopkb200), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cdouble}),
                    x.ctx.ptr, x.ptr, r.ptr, p.ptr, q.ptr, alpha, out))
        k += 1
        (xtol > 0 && alpha*sqrt(pp) ≤ xtol*sqrt(out[2])) && break
        oldrho = rho
        rho = out[1]
    end
    return x
end

conjgrad(A, b::DeviceVector{T}, x0::DeviceVector{T} = vfill!(vcreate(b), 0); kwds...) where {T} =
    conjgrad!(vcreate(b), A, b, x0; kwds...)

end # module

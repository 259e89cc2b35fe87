#!/usr/bin/env julia
#
# make_reference_golden.jl -- golden traces of the UNMODIFIED reference.
#
# This repository's parity is "unpinned" (DESIGN.md section 2): the reference is pure
# Julia, there is no Julia runtime where the repository is built and tested, and the
# reference's own tests hold no golden vector for vmlmb / spg.  Anyone who HAS Julia can
# close that gap with this script: it runs emmt/OptimPackNextGen.jl itself (public API,
# stock code path: `vmlmb`, `spg`) on the inputs of the repository's golden cases and
# records what the reference's own `printer` hooks see (src/quasinewton.jl:501-503,
# src/spg.jl:265-275), plus the final x.
#
#     julia --project=/path/to/OptimPackNextGen.jl tools/make_reference_golden.jl
#
# reads   tests/golden/reference_inputs/manifest.txt + *.bin   (tools/make_reference_inputs.py)
# writes  tests/golden/reference_vmlmb.json, tests/golden/reference_spg.json
#
# tests/test_reference_golden.py then holds the CPU oracle AND the CUDA path to these files
# at the north_star tolerances (first 20 iterations: same counts; f, |g|, step, x within 1e-10
# in Float64, 1e-4 in Float32).  Only the standard library is used besides the reference
# (no JSON.jl / NPZ.jl: raw binary in, hand-written JSON out).

using OptimPackNextGen
using LinearAlgebra
using Printf

const ROOT = normpath(joinpath(@__DIR__, ".."))
const INP = joinpath(ROOT, "tests", "golden", "reference_inputs")
const OUT = joinpath(ROOT, "tests", "golden")

eltype_of(s) = s == "float64" ? Float64 : s == "float32" ? Float32 : error("unknown dtype $s")

function readvec(::Type{T}, file, n) where {T}
    v = Vector{T}(undef, n)
    open(io -> read!(io, v), joinpath(INP, file))
    return v
end

# 'none' | 'scalar:<value>' | 'array:<file>'  ->  what `lower=` / `upper=` take
function bound(::Type{T}, field, n, default) where {T}
    field == "none" && return default
    kind, val = split(field, ':'; limit = 2)
    kind == "scalar" && return parse(Float64, val)
    kind == "array" && return readvec(T, val, n)
    error("bad bound field $field")
end

# --- objectives: the arithmetic of test/rosenbrock.jl:16-29 (element type T, no FMA) and of the
# --- repository's quadratic (g = a*(x - c); f = sum(g*(x - c))/2, products rounded in T).
struct Rosen end
function (::Rosen)(x::Vector{T}, g::Vector{T}) where {T}
    f1 = 0.0
    f2 = 0.0
    @inbounds for i in 1:2:length(x)
        x1 = x[i]
        x2 = x[i+1]
        t1 = one(T) - x1
        u = x2 - x1 * x1
        t2 = T(10) * u
        g2 = T(200) * u
        g[i] = -T(2) * (x1 * g2 + t1)
        g[i+1] = g2
        f1 += Float64(t1 * t1)
        f2 += Float64(t2 * t2)
    end
    return f1 + f2
end

struct Quad{T}
    a::Vector{T}
    c::Vector{T}
end
function (q::Quad{T})(x::Vector{T}, g::Vector{T}) where {T}
    f = 0.0
    @inbounds for i in eachindex(x)
        r = x[i] - q.c[i]
        gi = q.a[i] * r
        g[i] = gi
        f += Float64(gi * r)
    end
    return 0.5 * f
end

function objective(::Type{T}, field, n) where {T}
    field == "rosenbrock" && return Rosen()
    parts = split(field, ':')
    parts[1] == "quadratic" || error("bad objective field $field")
    return Quad{T}(readvec(T, parts[2], n), readvec(T, parts[3], n))
end

jnum(x::Real) = isfinite(x) ? @sprintf("%.17g", Float64(x)) : (isnan(x) ? "NaN" : (x > 0 ? "Infinity" : "-Infinity"))
jvec(v) = "[" * join((jnum(t) for t in v), ", ") * "]"
head(v) = v[1:min(8, length(v))]

function run_vmlmb(T, case, kind, n, mem, maxiter, x0, lo, hi, obj)
    lastx = copy(x0)
    fg!(x, g) = (copyto!(lastx, x); obj(x, g))
    recs = String[]
    printer(io, iter, eval, rejects, f, gnorm, stp) = push!(recs,
        "{\"iter\": $iter, \"eval\": $eval, \"rejects\": $rejects, \"f\": $(jnum(f)), \"gnorm\": $(jnum(gnorm)), " *
        "\"stp\": $(jnum(stp)), \"xnorm\": $(jnum(norm(Float64.(lastx)))), \"x_head\": $(jvec(head(lastx)))}")
    kw = Dict{Symbol,Any}(:mem => mem, :maxiter => maxiter, :verb => 1, :printer => printer, :output => devnull)
    lo === nothing || (kw[:lower] = lo)
    hi === nothing || (kw[:upper] = hi)
    x = vmlmb(fg!, x0; kw...)
    return "{\"case\": \"$case\", \"kind\": \"$kind\", \"n\": $n, \"dtype\": \"$(T == Float64 ? "float64" : "float32")\", " *
           "\"mem\": $mem, \"maxiter\": $maxiter,\n  \"final\": {\"xnorm\": $(jnum(norm(Float64.(x)))), \"x_head\": $(jvec(head(x)))},\n" *
           "  \"trace\": [\n   " * join(recs, ",\n   ") * "]}"
end

function run_spg(T, case, kind, n, m, maxit, x0, lo, hi, obj)
    lov = lo === nothing ? T(-Inf) : lo
    hiv = hi === nothing ? T(Inf) : hi
    # the box projection as a `prj!(dst, src)` (src/spg.jl:64-78; cf. `nonnegative!`, test/rosenbrock.jl:50-58)
    function prj!(dst, src)
        @inbounds for i in eachindex(dst, src)
            l = lov isa AbstractVector ? lov[i] : T(lov)
            h = hiv isa AbstractVector ? hiv[i] : T(hiv)
            v = src[i]
            v = v > h ? h : v
            dst[i] = v < l ? l : v
        end
        return dst
    end
    lastx = copy(x0)
    fg!(x, g) = (copyto!(lastx, x); obj(x, g))
    recs = String[]
    printer(io, ws) = push!(recs,
        "{\"iter\": $(ws.iter), \"fcnt\": $(ws.fcnt), \"pcnt\": $(ws.pcnt), \"f\": $(jnum(ws.f)), \"fbest\": $(jnum(ws.fbest)), " *
        "\"pgtwon\": $(jnum(ws.pgtwon)), \"pginfn\": $(jnum(ws.pginfn)), \"x_head\": $(jvec(head(lastx)))}")
    ws = OptimPackNextGen.SPG.Info()
    x = spg(fg!, prj!, x0, m; maxit = maxit, verb = 1, printer = printer, io = devnull, ws = ws)
    return "{\"case\": \"$case\", \"kind\": \"$kind\", \"n\": $n, \"dtype\": \"$(T == Float64 ? "float64" : "float32")\", " *
           "\"m\": $m, \"maxit\": $maxit,\n  \"final\": {\"iter\": $(ws.iter), \"fcnt\": $(ws.fcnt), \"pcnt\": $(ws.pcnt), " *
           "\"status\": $(Int(ws.status)), \"fbest\": $(jnum(ws.fbest)), \"xnorm\": $(jnum(norm(Float64.(x)))), " *
           "\"x_head\": $(jvec(head(x)))},\n  \"trace\": [\n   " * join(recs, ",\n   ") * "]}"
end

function main()
    out = Dict("vmlmb" => String[], "spg" => String[])
    for line in eachline(joinpath(INP, "manifest.txt"))
        (isempty(strip(line)) || startswith(line, "#")) && continue
        solver, case, kind, ns, dt, mems, its, x0f, lof, hif, objf = split(line)
        T = eltype_of(dt)
        n, mem, maxiter = parse(Int, ns), parse(Int, mems), parse(Int, its)
        x0 = readvec(T, x0f, n)
        lo = bound(T, lof, n, nothing)
        hi = bound(T, hif, n, nothing)
        obj = objective(T, objf, n)
        rec = solver == "vmlmb" ? run_vmlmb(T, case, kind, n, mem, maxiter, x0, lo, hi, obj) :
                                  run_spg(T, case, kind, n, mem, maxiter, x0, lo, hi, obj)
        push!(out[solver], rec)
        println(stderr, "done: $solver $case ($kind n=$n $dt)")
    end
    ver = try
        string(pkgversion(OptimPackNextGen))
    catch
        "unknown"
    end
    for solver in ("vmlmb", "spg")
        open(joinpath(OUT, "reference_$(solver).json"), "w") do io
            print(io, "{\"generator\": \"tools/make_reference_golden.jl\", \"reference\": \"OptimPackNextGen.jl $ver\", " *
                      "\"julia\": \"$(VERSION)\",\n \"cases\": [\n ", join(out[solver], ",\n "), "\n]}\n")
        end
    end
    println(stderr, "wrote ", joinpath(OUT, "reference_vmlmb.json"), " and reference_spg.json")
end

main()

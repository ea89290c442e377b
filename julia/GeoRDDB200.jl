# GeoRDDB200.jl -- the reference-side binding: thin `ccall` wrappers that keep GeoRDD.jl's Julia
# signatures and route the hot path into libgeordd_b200.so (include/geordd_b200.h).
#
# A maintainer `include`s this file at the end of src/GeoRDD.jl (after the stock definitions); the
# methods below then replace the bodies of cliff_face / boot_chi2test / nsim_chi / boot_mlltest /
# nsim_logP / boot_invvar / nsim_invvar_pval_uncalib / pval_invvar_calib / nsim_power /
# inverse_variance for `B200GPE` arguments.  Julia is not installed in the build container, so this
# file is syntax-reviewed only; the executable twin of every call below is geordd.jl_b200/capi.py +
# geordd.jl_b200/geordd.py (same symbols, same argument order), which the GPU tests exercise.
#
# No fallback: unsupported kernels throw ArgumentError, a missing library / GPU is an error.

module GeoRDDB200

using LinearAlgebra
using GaussianProcesses: SEIso, Mat12Iso, Mat32Iso, Mat52Iso, Const, SumKernel, FixedKernel,
                         MeanZero, MeanConst, Kernel, Mean
using Distributions: Normal

const LIB = get(ENV, "GEORDD_B200_LIB", "libgeordd_b200")

struct CKernel                      # geordd_kernel
    kind::Cint
    ell::Cdouble
    sigma2::Cdouble
    const_sigma2::Cdouble
    mean_const::Cdouble
end

# ---- kernel / mean pattern matching (SURVEY §8b) ------------------------------------------------
unfix(k::FixedKernel) = k.kernel
unfix(k::Kernel) = k
spatial(k::SEIso)    = (Cint(0), sqrt(k.ℓ2), k.σ2)
spatial(k::Mat12Iso) = (Cint(1), k.ℓ, k.σ2)
spatial(k::Mat32Iso) = (Cint(2), k.ℓ, k.σ2)
spatial(k::Mat52Iso) = (Cint(3), k.ℓ, k.σ2)
spatial(k) = throw(ArgumentError("geordd_b200: unsupported kernel $(typeof(k))"))
meanconst(::MeanZero) = 0.0
meanconst(m::MeanConst) = m.β
meanconst(m) = throw(ArgumentError("geordd_b200: unsupported mean $(typeof(m))"))

function ckernel(k::Kernel, m::Mean)
    k = unfix(k)
    csig2 = 0.0
    if k isa SumKernel
        a, b = unfix(k.kleft), unfix(k.kright)
        if a isa Const
            a, b = b, a
        end
        b isa Const || throw(ArgumentError("geordd_b200: only spatial + Const sums are supported"))
        csig2 = b.σ2
        k = a
    end
    kind, ell, s2 = spatial(k)
    return CKernel(kind, ell, s2, csig2, meanconst(m))
end

# ---- context -------------------------------------------------------------------------------------
mutable struct Context
    ptr::Ptr{Cvoid}
end
function Context(device::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:geordd_ctx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), out, device)
    rc == 0 || error("geordd_b200: no usable sm_100 CUDA device (rc=$rc); there is no CPU fallback")
    ctx = Context(out[])
    finalizer(c -> ccall((:geordd_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end
const DEFAULT = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT[] === nothing && (DEFAULT[] = Context(parse(Int, get(ENV, "GEORDD_DEVICE", "0")))); DEFAULT[])

function check(ctx::Context, rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:geordd_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx.ptr))
    rc > 0 && throw(LinearAlgebra.PosDefException(rc))      # keeps src/GPrealisationsCovars.jl:239-241 working
    rc == -1 && throw(ArgumentError(msg))
    error("geordd_b200 ($rc): $msg")
end

# ---- GPE on the device (replaces GPE(x, y, mean, kernel, logNoise), src/region_gp_fit.jl:12,28) ----
mutable struct B200GPE
    ctx::Context
    ptr::Ptr{Cvoid}
    x::Matrix{Float64}
    y::Vector{Float64}
    mean::Mean
    kernel::Kernel
    logNoise::Float64
    alpha::Vector{Float64}
    mll::Float64
    nobs::Int
    dim::Int
end

function B200GPE(x::AbstractMatrix, y::AbstractVector, mean::Mean, kernel::Kernel, logNoise::Real;
                 ctx::Context=default_context())
    X = Matrix{Float64}(x); Y = Vector{Float64}(y)
    dim, n = size(X)
    ck = Ref(ckernel(kernel, mean))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    alpha = Vector{Float64}(undef, n); mll = Ref{Cdouble}(0.0); rounds = Ref{Cint}(0)
    GC.@preserve X Y alpha begin
        rc = ccall((:geordd_gp_fit, LIB), Cint,
                   (Ptr{Cvoid}, Ref{CKernel}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ref{Ptr{Cvoid}},
                    Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ref{Cint}),
                   ctx.ptr, ck, X, dim, n, Y, Float64(logNoise), out, alpha, mll, C_NULL, rounds)
    end
    check(ctx, rc)
    gp = B200GPE(ctx, out[], X, Y, mean, kernel, Float64(logNoise), alpha, mll[], n, dim)
    finalizer(g -> ccall((:geordd_gp_destroy, LIB), Cint, (Ptr{Cvoid},), g.ptr), gp)
    return gp
end

"update_alpha!(gp) + mll(gp) (src/gp_utils.jl:15-22)"
function update_alpha!(gp::B200GPE)
    mll = Ref{Cdouble}(0.0)
    GC.@preserve gp begin
        check(gp.ctx, ccall((:geordd_gp_set_y, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                            gp.ptr, gp.y, gp.alpha, mll))
    end
    gp.mll = mll[]
    return gp.alpha
end

# ---- cliff_face(gpT, gpC, sentinels) (src/cliff_face.jl:3-9) ----------------------------------------
function cliff_face(gpT::B200GPE, gpC::B200GPE, sentinels::AbstractMatrix)
    Xb = Matrix{Float64}(sentinels); nb = size(Xb, 2)
    μ = Vector{Float64}(undef, nb); Σ = Matrix{Float64}(undef, nb, nb)
    GC.@preserve Xb μ Σ begin
        check(gpT.ctx, ccall((:geordd_cliff_face, LIB), Cint,
                             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                             gpT.ptr, gpC.ptr, Xb, nb, μ, Σ))
    end
    return μ, Σ
end

"inverse_variance(μ, Σ; nugget) (src/average_treatment_effect.jl:3-11)"
function inverse_variance(μ::AbstractVector, Σ::AbstractMatrix; nugget=1e-10, ctx::Context=default_context())
    m = Vector{Float64}(μ); S = Matrix{Float64}(Σ); τ = Ref{Cdouble}(0.0); V = Ref{Cdouble}(0.0)
    GC.@preserve m S begin
        check(ctx, ccall((:geordd_inverse_variance, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}, Ref{Cdouble}),
                         ctx.ptr, m, S, length(m), Float64(nugget), τ, V))
    end
    return Normal(τ[], √(V[]))
end

# ---- null handle: make_null + draw-independent setup (src/hypotest.jl:1-12, src/hypotest_chi2.jl:33-48) ----
mutable struct B200Null
    ctx::Context
    ptr::Ptr{Cvoid}
    gpT::B200GPE
    gpC::B200GPE
    mll::Float64
    nobs::Int
end

function make_null(gpT::B200GPE, gpC::B200GPE; Xb::Union{Nothing,AbstractMatrix}=nothing, nugget::Real=0.0)
    out = Ref{Ptr{Cvoid}}(C_NULL); rounds = Ref{Cint}(0); mll = Ref{Cdouble}(0.0)
    X = Xb === nothing ? Matrix{Float64}(undef, 2, 0) : Matrix{Float64}(Xb)
    GC.@preserve X begin
        rc = ccall((:geordd_null_prepare, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Ptr{Cvoid}}, Ref{Cint}, Ref{Cdouble}),
                   gpT.ptr, gpC.ptr, size(X, 2) == 0 ? C_NULL : pointer(X), size(X, 2), Float64(nugget), out, rounds, mll)
    end
    check(gpT.ctx, rc)
    h = B200Null(gpT.ctx, out[], gpT, gpC, mll[], gpT.nobs + gpC.nobs)
    finalizer(n -> ccall((:geordd_null_destroy, LIB), Cint, (Ptr{Cvoid},), n.ptr), h)
    return h
end

function set_sentinels!(h::B200Null, Xb::AbstractMatrix, nugget::Real)
    X = Matrix{Float64}(Xb); rounds = Ref{Cint}(0)
    GC.@preserve X begin
        check(h.ctx, ccall((:geordd_null_set_sentinels, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cint}),
                           h.ptr, X, size(X, 2), Float64(nugget), rounds))
    end
    return rounds[]
end

"chistat(gpT, gpC, Xb) observed statistic (src/hypotest_chi2.jl:4-10)"
function chistat(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix)
    X = Matrix{Float64}(Xb); out = Ref{Cdouble}(0.0)
    GC.@preserve X begin
        check(gpT.ctx, ccall((:geordd_chistat_obs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ref{Cdouble}),
                             gpT.ptr, gpC.ptr, X, size(X, 2), out))
    end
    return out[]
end

"nsim_chi(gpT, gpC, gpNull, Xb, nsim) (src/hypotest_chi2.jl:33-52); seed/draw0 select the Philox draws"
function nsim_chi(gpT::B200GPE, gpC::B200GPE, gpNull::B200Null, Xb::AbstractMatrix, nsim::Int;
                  seed::Integer=rand(UInt64), draw0::Integer=0, chi_obs::Real=Inf)
    set_sentinels!(gpNull, Xb, 0.0)
    stats = Vector{Float64}(undef, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve stats begin
        check(gpT.ctx, ccall((:geordd_null_chi2, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                             gpNull.ptr, nsim, seed, draw0, C_NULL, Float64(chi_obs), stats, cnt))
    end
    return stats
end

"boot_chi2test(gpT, gpC, Xb, nsim) (src/hypotest_chi2.jl:54-59)"
function boot_chi2test(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Int; kwargs...)
    gpNull = make_null(gpT, gpC)
    chi_obs = chistat(gpT, gpC, Xb)
    chi_sims = nsim_chi(gpT, gpC, gpNull, Xb, nsim; kwargs...)
    return mean(chi_sims .> chi_obs)
end

"nsim_logP(gpT, gpC, gpNull, nsim) (src/hypotest_mll.jl:21-30): vector of (mll_null, mll_altv)"
function nsim_logP(gpT::B200GPE, gpC::B200GPE, gpNull::B200Null, nsim::Int; seed::Integer=rand(UInt64), draw0::Integer=0)
    mnull = Vector{Float64}(undef, nsim); malt = Vector{Float64}(undef, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve mnull malt begin
        check(gpT.ctx, ccall((:geordd_null_mll, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble},
                              Ref{Clonglong}, Ptr{Cdouble}, Ptr{Cdouble}),
                             gpNull.ptr, nsim, seed, draw0, C_NULL, Inf, mnull, malt, cnt, C_NULL, C_NULL))
    end
    gpNull.mll = mnull[end]                      # the reference leaves gpNull at the last draw (:6,10,14)
    return collect(zip(mnull, malt))
end

"""
    projection_points(gp, border, maxdist)   (src/border_projection.jl:4-22)

Nearest border point of every unit within `maxdist` of the border -- one device call instead of a LibGEOS call per unit.
`border` is the LibGEOS (Multi)LineString; only its vertex coordinates cross the ABI.
"""
function projection_points(gp::B200GPE, border, maxdist::Float64)
    lines = border isa LibGEOS.MultiLineString ? GeoInterface.coordinates(border) : [GeoInterface.coordinates(border)]
    V = Matrix{Float64}(undef, 2, sum(length, lines)); starts = Cint[]; v = 0
    for (k, line) in enumerate(lines)
        k > 1 && push!(starts, v)
        for p in line
            v += 1; V[1, v] = p[1]; V[2, v] = p[2]
        end
    end
    n = size(gp.x, 2); Xp = Matrix{Float64}(undef, 2, n); d = Vector{Float64}(undef, n)
    GC.@preserve V starts Xp d begin
        check(gp.ctx, ccall((:geordd_projection_points, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                            gp.ctx.ptr, gp.x, n, V, size(V, 2), starts, length(starts), Xp, d))
    end
    return Xp[:, d .<= maxdist]
end

"proj_estimator(gpT, gpC, border, maxdist) (src/border_projection.jl:24-31)"
function proj_estimator(gpT::B200GPE, gpC::B200GPE, border, maxdist::Float64)
    X∂ = [projection_points(gpT, border, maxdist) projection_points(gpC, border, maxdist)]
    μpost, Σpost = cliff_face(gpT, gpC, X∂)
    return unweighted_mean(μpost, Σpost)
end

"""
    within_region(ctx, P, region)

`LibGEOS.within(point, region)` for every column of `P` (2 x n) on the device -- the grid filter of
`infinite_proj_sentinels` (src/border_projection.jl:44-102).  The (Multi)Polygon crosses the ABI as the vertex arrays
of its closed rings.
"""
function within_region(ctx, P::Matrix{Float64}, region)
    polys = region isa LibGEOS.MultiPolygon ? GeoInterface.coordinates(region) : [GeoInterface.coordinates(region)]
    rings = [ring for poly in polys for ring in poly]
    R = Matrix{Float64}(undef, 2, sum(length, rings)); starts = Cint[]; v = 0
    for (k, ring) in enumerate(rings)
        k > 1 && push!(starts, v)
        for p in ring
            v += 1; R[1, v] = p[1]; R[2, v] = p[2]
        end
    end
    n = size(P, 2); inside = Vector{UInt8}(undef, n)
    GC.@preserve P R starts inside begin
        check(ctx, ccall((:geordd_points_in_region, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{UInt8}),
                         ctx.ptr, P, n, R, size(R, 2), starts, length(starts), inside))
    end
    return inside .!= 0
end

"""
    set_mll_forward_only!(ctx, on)

Opt-in (default off): evaluate the quadratic forms of `nsim_logP` / `boot_mlltest` as |L⁻¹(y-m)|² (and z'z for the
pooled null GP) instead of the reference's stepwise `alpha = cK \\ (y-m); dot(y-m, alpha)` -- the same statistic to
rounding (~1e-12 relative) at 6 m² instead of 16 m² flop per draw.
"""
set_mll_forward_only!(ctx, on::Bool) =
    check(ctx, ccall((:geordd_ctx_set_mll_forward_only, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.ptr, on ? 1 : 0))

"boot_mlltest(gpT, gpC, nsim) (src/hypotest_mll.jl:32-45)"
function boot_mlltest(gpT::B200GPE, gpC::B200GPE, nsim::Int; kwargs...)
    mll_alt = gpT.mll + gpC.mll
    gpNull = make_null(gpT, gpC)
    mll_null = gpNull.mll
    sims = nsim_logP(gpT, gpC, gpNull, nsim; kwargs...)
    Δobs = mll_alt - mll_null
    Δsim = [s[2] - s[1] for s in sims]
    return mean(Δsim .> Δobs)
end

"pval_invvar_uncalib(gpT, gpC, Xb; nugget) (src/hypotest_invvar.jl:1-6)"
function pval_invvar_uncalib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix; nugget=1e-10)
    X = Matrix{Float64}(Xb); τ = Ref{Cdouble}(0.0); V = Ref{Cdouble}(0.0); p = Ref{Cdouble}(0.0)
    GC.@preserve X begin
        check(gpT.ctx, ccall((:geordd_invvar_obs, LIB), Cint,
                             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
                             gpT.ptr, gpC.ptr, X, size(X, 2), Float64(nugget), τ, V, p))
    end
    return p[]
end

"nsim_invvar_pval_uncalib(gpT, gpC, Xb, nsim; nugget) (src/hypotest_invvar.jl:48-63)"
function nsim_invvar_pval_uncalib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Int; nugget=1e-10,
                                  seed::Integer=rand(UInt64), draw0::Integer=0)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=nugget)
    pv = Vector{Float64}(undef, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve pv begin
        check(gpT.ctx, ccall((:geordd_null_invvar, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                             gpNull.ptr, nsim, seed, draw0, C_NULL, -1.0, pv, cnt))
    end
    return pv
end

"boot_invvar(gpT, gpC, Xb, nsim; nugget) (src/hypotest_invvar.jl:65-69)"
function boot_invvar(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Int; nugget=1e-10, kwargs...)
    pval_obs = pval_invvar_uncalib(gpT, gpC, Xb; nugget=nugget)
    pval_sims = nsim_invvar_pval_uncalib(gpT, gpC, Xb, nsim; nugget=nugget, kwargs...)
    return mean(pval_sims .< pval_obs)
end

"pval_invvar_calib(gpT, gpC, Xb; nugget) (src/hypotest_invvar.jl:74-105)"
function pval_invvar_calib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix; nugget=1e-10)
    X = Matrix{Float64}(Xb); p = Ref{Cdouble}(0.0)
    GC.@preserve X begin
        check(gpT.ctx, ccall((:geordd_invvar_calib, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}),
                             gpT.ptr, gpC.ptr, X, size(X, 2), Float64(nugget), p))
    end
    return p[]
end

"nsim_power(gpT, gpC, τ, Xb, chi_null, mll_null, pval_invvar_null, nsim) (src/power.jl:46-73)"
function nsim_power(gpT::B200GPE, gpC::B200GPE, τ::Float64, Xb::AbstractMatrix, chi_null::Vector{Float64},
                    mll_null::Vector{Float64}, pval_invvar_null::Vector{Float64}, nsim::Int;
                    seed::Integer=rand(UInt64), draw0::Integer=0, compat_calib_bug::Bool=false)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=0.0)
    out = Matrix{Float64}(undef, 5, nsim)
    GC.@preserve chi_null mll_null pval_invvar_null out begin
        check(gpT.ctx, ccall((:geordd_power, LIB), Cint,
                             (Ptr{Cvoid}, Cdouble, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                              Ptr{Cdouble}, Clong, Cint, Ptr{Cdouble}),
                             gpNull.ptr, τ, nsim, seed, draw0, C_NULL, chi_null, mll_null, pval_invvar_null,
                             length(chi_null), compat_calib_bug ? 1 : 0, out))
    end
    return [Tuple(out[:, s]) for s in 1:nsim]
end

end # module

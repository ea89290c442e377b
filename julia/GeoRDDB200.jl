# GeoRDDB200.jl -- the reference-side binding: thin `ccall` wrappers that keep GeoRDD.jl's Julia signatures and route
# the hot path into libgeordd_b200.so (include/geordd_b200.h).
#
# A maintainer adds ONE line at the end of src/GeoRDD.jl (after the stock definitions):
#     include(joinpath(@__DIR__, "..", "julia", "GeoRDDB200.jl"))
# which defines the submodule GeoRDD.GeoRDDB200 and -- unless ENV["GEORDD_B200_OVERRIDE"] == "0" -- re-defines the bodies
# of the stock hot-path entry points for their STOCK argument types (`cliff_face(gpT::GPE, gpC::GPE, Xb)`,
# `boot_chi2test`, `nsim_chi`, `boot_mlltest`, `nsim_logP`, `boot_invvar`, `nsim_invvar_pval_uncalib`, `nsim_invvar_calib`,
# `pval_invvar_uncalib`, `pval_invvar_calib`, `nsim_power`, `chistat`, `weight_at_units`, `projection_points`,
# `update_mll!` / `update_mll_and_dmll!` / `initialise_mll!` / `postmean_β` of `MultiGPCovars`, `GPRealisations`,
# `placebo_chi` / `placebo_mll` / `placebo_invvar`), so existing scripts and notebooks run unchanged (section 9 below).
# The device-resident types (`B200GPE`, `B200Null`, `B200MultiGP`, `DeviceGroup`) are also usable directly.
#
# Julia is not installed in the build container, so this file is syntax-reviewed only; the executable twin of every call
# below is geordd.jl_b200/capi.py + geordd.jl_b200/geordd.py (same symbols, same argument order), which the GPU tests
# exercise.  No fallback: unsupported kernels throw ArgumentError, a missing library / GPU is an error.
#
# Draws: every Monte-Carlo entry point takes `Z` (n x nsim standard normals, column s = the randn(n) that
# `rand(MvNormal)` would consume for draw s -- pass `randn(n, nsim)` to reproduce a seeded Julia run exactly) or, with
# `Z === nothing`, device Philox normals keyed by (`seed`, `draw0` + s, element).

module GeoRDDB200

using LinearAlgebra
using Statistics: mean
import GaussianProcesses
using GaussianProcesses: GPE, SEIso, Mat12Iso, Mat32Iso, Mat52Iso, Const, LinIso, SumKernel, FixedKernel,
                         MeanZero, MeanConst, Kernel, Mean, get_params, num_params
using Distributions: Normal, ccdf
import LibGEOS
import GeoInterface
import Optim
import ..GeoRDD                      # the enclosing package: stock types, geometry helpers, LATE reducers
using ..GeoRDD: MultiGPCovars, RegionData, unweighted_mean, weighted_mean

const LIB = get(ENV, "GEORDD_B200_LIB", "libgeordd_b200")

struct CKernel                      # geordd_kernel
    kind::Cint
    ell::Cdouble
    sigma2::Cdouble
    const_sigma2::Cdouble
    mean_const::Cdouble
end

# ---- 1. kernel / mean pattern matching (SURVEY §8b) ---------------------------------------------------------------
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

lognoise(x::Real) = Float64(x)
lognoise(x) = Float64(x.value)       # GaussianProcesses wraps logNoise in a `Scalar`

# ---- 2. context ---------------------------------------------------------------------------------------------------
mutable struct Context
    ptr::Ptr{Cvoid}
    owned::Bool
end
function Context(device::Integer=0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:geordd_ctx_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), out, device)
    rc == 0 || error("geordd_b200: no usable sm_100 CUDA device (rc=$rc); there is no CPU fallback")
    ctx = Context(out[], true)
    finalizer(c -> c.owned && ccall((:geordd_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
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

zptr(::Nothing) = Ptr{Cdouble}(C_NULL)
zptr(Z::Matrix{Float64}) = pointer(Z)
zmat(::Nothing, n) = nothing
function zmat(Z::AbstractMatrix, n)
    size(Z, 1) == n || throw(ArgumentError("Z must be n x nsim (column s = randn(n) of draw s)"))
    return Matrix{Float64}(Z)
end
ndraws(::Nothing, nsim) = Int(nsim)
ndraws(Z::Matrix{Float64}, nsim) = size(Z, 2)

# ---- 3. GPE on the device (replaces GPE(x, y, mean, kernel, logNoise), src/region_gp_fit.jl:12,28; src/hypotest.jl:4) ----
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
    owned::Bool
end
destroy_gp(g::B200GPE) = g.owned && ccall((:geordd_gp_destroy, LIB), Cint, (Ptr{Cvoid},), g.ptr)

function B200GPE(x::AbstractMatrix, y::AbstractVector, mean::Mean, kernel::Kernel, logNoise; ctx::Context=default_context())
    X = Matrix{Float64}(x); Y = Vector{Float64}(y)
    dim, n = size(X)
    ck = Ref(ckernel(kernel, mean))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    alpha = Vector{Float64}(undef, n); mll = Ref{Cdouble}(0.0); rounds = Ref{Cint}(0)
    rc = GC.@preserve X Y alpha ccall((:geordd_gp_fit, LIB), Cint,
                   (Ptr{Cvoid}, Ref{CKernel}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ref{Ptr{Cvoid}},
                    Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ref{Cint}),
                   ctx.ptr, ck, X, dim, n, Y, lognoise(logNoise), out, alpha, mll, C_NULL, rounds)
    check(ctx, rc)
    gp = B200GPE(ctx, out[], X, Y, mean, kernel, lognoise(logNoise), alpha, mll[], n, dim, true)
    finalizer(destroy_gp, gp)
    return gp
end

# A stock GPE moved to the device (same data, kernel, mean, noise; re-assembled and re-factored there).  Cached per
# object and parameter state, so repeated calls of the stock-signature entry points fit once.
const GPE_CACHE = IdDict{Any,Tuple{UInt64,B200GPE}}()
gpe_state(gp::GPE) = hash((gp.x, gp.y, get_params(gp.kernel), get_params(gp.mean), lognoise(gp.logNoise)))
function B200GPE(gp::GPE; ctx::Context=default_context())
    st = gpe_state(gp)
    hit = get(GPE_CACHE, gp, nothing)
    hit !== nothing && hit[1] == st && return hit[2]
    dgp = B200GPE(gp.x, gp.y, gp.mean, gp.kernel, gp.logNoise; ctx=ctx)
    GPE_CACHE[gp] = (st, dgp)
    return dgp
end
B200GPE(gp::B200GPE; kwargs...) = gp
const AnyGPE = Union{GPE,B200GPE}

"""
    fit_many(xs, ys, mean, kernel, logNoise)

The independent fits that `GPRealisations` (src/GPrealisations.jl:10-19) and the placebo drivers make one after another, as
ONE call (`geordd_gp_fit_batch`): all factorisations run in a single dataflow launch.  Same results as
`[B200GPE(x, y, mean, kernel, logNoise) for ...]`, bit for bit.
"""
function fit_many(xs::Vector{<:AbstractMatrix}, ys::Vector{<:AbstractVector}, mean::Mean, kernel::Kernel, logNoise;
                  ctx::Context=default_context())
    nfit = length(xs)
    Xs = [Matrix{Float64}(x) for x in xs]; Ys = [Vector{Float64}(y) for y in ys]
    dim = size(Xs[1], 1)
    all(size(X, 1) == dim for X in Xs) || throw(ArgumentError("all regions must have the same spatial dimension"))
    ns = Cint[size(X, 2) for X in Xs]
    alphas = [Vector{Float64}(undef, n) for n in ns]
    ck = fill(ckernel(kernel, mean), nfit)
    lns = fill(lognoise(logNoise), nfit)
    Xp = [pointer(X) for X in Xs]; Yp = [pointer(Y) for Y in Ys]; Ap = [pointer(a) for a in alphas]
    out = fill(Ptr{Cvoid}(C_NULL), nfit); mll = zeros(Cdouble, nfit); jr = zeros(Cint, nfit); info = zeros(Cint, nfit)
    rc = GC.@preserve Xs Ys alphas Xp Yp Ap ccall((:geordd_gp_fit_batch, LIB), Cint,
                   (Ptr{Cvoid}, Cint, Ptr{CKernel}, Ptr{Ptr{Cdouble}}, Cint, Ptr{Cint}, Ptr{Ptr{Cdouble}}, Ptr{Cdouble},
                    Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cdouble}}, Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}),
                   ctx.ptr, nfit, ck, Xp, dim, ns, Yp, lns, out, Ap, mll, jr, info)
    check(ctx, rc)
    gps = B200GPE[]
    for f in 1:nfit
        gp = B200GPE(ctx, out[f], Xs[f], Ys[f], mean, kernel, lognoise(logNoise), alphas[f], mll[f], Int(ns[f]), dim, true)
        finalizer(destroy_gp, gp)
        push!(gps, gp)
    end
    return gps
end

"update_alpha!(gp) + mll(gp) (src/gp_utils.jl:15-22)"
function update_alpha!(gp::B200GPE)
    mll = Ref{Cdouble}(0.0)
    GC.@preserve gp check(gp.ctx, ccall((:geordd_gp_set_y, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                                        gp.ptr, gp.y, gp.alpha, mll))
    gp.mll = mll[]
    return gp.alpha
end
mll(gp::B200GPE) = gp.mll

"predict_mu(gp, X, cK) (src/gp_utils.jl:1-5); cK is recomputed on the device, the argument is accepted for signature compatibility"
function predict_mu(gp::B200GPE, X::AbstractMatrix, cK=nothing)
    Xb = Matrix{Float64}(X); mu = Vector{Float64}(undef, size(Xb, 2))
    GC.@preserve Xb mu check(gp.ctx, ccall((:geordd_gp_predict_mu, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}),
                                           gp.ptr, Xb, size(Xb, 2), mu))
    return mu
end

"prior_rand(gp) (src/hypotest.jl:13-18) for nsim draws: the columns of m + L z"
function prior_rand(gp::B200GPE, nsim::Integer=1; seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing)
    Zm = zmat(Z, gp.nobs); ns = ndraws(Zm, nsim)
    Y = Matrix{Float64}(undef, gp.nobs, ns)
    GC.@preserve Zm Y check(gp.ctx, ccall((:geordd_gp_prior_rand, LIB), Cint,
                                          (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}),
                                          gp.ptr, ns, seed, draw0, zptr(Zm), Y))
    return ns == 1 ? vec(Y) : Y
end

# ---- 4. cliff face, LATE reducers, unit weights ------------------------------------------------------------------------
"cliff_face(gpT, gpC, sentinels) (src/cliff_face.jl:3-9)"
function cliff_face(gpT::B200GPE, gpC::B200GPE, sentinels::AbstractMatrix)
    Xb = Matrix{Float64}(sentinels); nb = size(Xb, 2)
    μ = Vector{Float64}(undef, nb); Σ = Matrix{Float64}(undef, nb, nb)
    GC.@preserve Xb μ Σ check(gpT.ctx, ccall((:geordd_cliff_face, LIB), Cint,
                             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                             gpT.ptr, gpC.ptr, Xb, nb, μ, Σ))
    return μ, Σ
end

"inverse_variance(μ, Σ; nugget) (src/average_treatment_effect.jl:3-11)"
function inverse_variance(μ::AbstractVector, Σ::AbstractMatrix; nugget=1e-10, ctx::Context=default_context())
    m = Vector{Float64}(μ); S = Matrix{Float64}(Σ); τ = Ref{Cdouble}(0.0); V = Ref{Cdouble}(0.0)
    GC.@preserve m S check(ctx, ccall((:geordd_inverse_variance, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}, Ref{Cdouble}),
                         ctx.ptr, m, S, length(m), Float64(nugget), τ, V))
    return Normal(τ[], √(V[]))
end

"weight_at_units(gp, sentinels, weights) (src/weight_at_units.jl:12-18)"
function weight_at_units(gp::B200GPE, sentinels::AbstractMatrix, weights::AbstractVector)
    Xb = Matrix{Float64}(sentinels); w = Vector{Float64}(weights); out = Vector{Float64}(undef, gp.nobs)
    size(Xb, 2) == length(w) || throw(ArgumentError("one weight per sentinel"))
    GC.@preserve Xb w out check(gp.ctx, ccall((:geordd_weight_at_units, LIB), Cint,
                                              (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                                              gp.ptr, Xb, size(Xb, 2), w, out))
    return out
end

# ---- 5. null handle and the Monte-Carlo tests ---------------------------------------------------------------------
mutable struct B200Null               # make_null + draw-independent setup (src/hypotest.jl:1-12, src/hypotest_chi2.jl:33-48)
    ctx::Context
    ptr::Ptr{Cvoid}
    gpT::B200GPE
    gpC::B200GPE
    mll::Float64
    nobs::Int
    y::Vector{Float64}
    alpha::Vector{Float64}
    Xb::Union{Nothing,Matrix{Float64}}   # sentinels the handle is currently prepared for (and their nugget)
    nugget::Float64
end

function make_null(gpT::B200GPE, gpC::B200GPE; Xb::Union{Nothing,AbstractMatrix}=nothing, nugget::Real=0.0)
    out = Ref{Ptr{Cvoid}}(C_NULL); rounds = Ref{Cint}(0); mll = Ref{Cdouble}(0.0)
    X = Xb === nothing ? Matrix{Float64}(undef, 2, 0) : Matrix{Float64}(Xb)
    rc = GC.@preserve X ccall((:geordd_null_prepare, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Ptr{Cvoid}}, Ref{Cint}, Ref{Cdouble}),
                   gpT.ptr, gpC.ptr, size(X, 2) == 0 ? C_NULL : pointer(X), size(X, 2), Float64(nugget), out, rounds, mll)
    check(gpT.ctx, rc)
    n = gpT.nobs + gpC.nobs
    h = B200Null(gpT.ctx, out[], gpT, gpC, mll[], n, vcat(gpT.y, gpC.y), Vector{Float64}(undef, n),
                 size(X, 2) == 0 ? nothing : X, Float64(nugget))
    finalizer(n -> ccall((:geordd_null_destroy, LIB), Cint, (Ptr{Cvoid},), n.ptr), h)
    return h
end

function set_sentinels!(h::B200Null, Xb::AbstractMatrix, nugget::Real)
    X = Matrix{Float64}(Xb); rounds = Ref{Cint}(0)
    h.Xb !== nothing && h.Xb == X && h.nugget == nugget && return 0     # already prepared for these sentinels
    GC.@preserve X check(h.ctx, ccall((:geordd_null_set_sentinels, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cint}),
                                      h.ptr, X, size(X, 2), Float64(nugget), rounds))
    h.Xb = X; h.nugget = Float64(nugget)
    return rounds[]
end

"chistat(gpT, gpC, Xb) of the observed outcomes from the cliff face the handle evaluated for its sentinels (bit-identical to chistat)"
function chistat(h::B200Null)
    out = Ref{Cdouble}(0.0)
    check(h.ctx, ccall((:geordd_null_chistat_obs, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}), h.ptr, out))
    return out[]
end

"pval_invvar_uncalib(gpT, gpC, Xb; nugget) from the handle's cliff face (bit-identical to the three-argument method)"
function pval_invvar_uncalib(h::B200Null)
    τ = Ref{Cdouble}(0.0); V = Ref{Cdouble}(0.0); p = Ref{Cdouble}(0.0)
    check(h.ctx, ccall((:geordd_null_invvar_obs, LIB), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}), h.ptr, τ, V, p))
    return p[]
end

"chistat(gpT, gpC, Xb) observed statistic (src/hypotest_chi2.jl:4-10)"
function chistat(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix)
    X = Matrix{Float64}(Xb); out = Ref{Cdouble}(0.0)
    GC.@preserve X check(gpT.ctx, ccall((:geordd_chistat_obs, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ref{Cdouble}),
                                        gpT.ptr, gpC.ptr, X, size(X, 2), out))
    return out[]
end

"nsim_chi(gpT, gpC, gpNull, Xb, nsim) (src/hypotest_chi2.jl:33-52)"
function nsim_chi(gpT::B200GPE, gpC::B200GPE, gpNull::B200Null, Xb::AbstractMatrix, nsim::Integer;
                  seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing, chi_obs::Real=Inf)
    set_sentinels!(gpNull, Xb, 0.0)
    Zm = zmat(Z, gpNull.nobs); ns = ndraws(Zm, nsim)
    stats = Vector{Float64}(undef, ns); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm stats check(gpT.ctx, ccall((:geordd_null_chi2, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                             gpNull.ptr, ns, seed, draw0, zptr(Zm), Float64(chi_obs), stats, cnt))
    return stats
end

"boot_chi2test(gpT, gpC, Xb, nsim) (src/hypotest_chi2.jl:54-59)"
function boot_chi2test(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Integer; kwargs...)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=0.0)      # one cliff face serves the observed statistic and the draws
    chi_obs = chistat(gpNull)
    chi_sims = nsim_chi(gpT, gpC, gpNull, Xb, nsim; kwargs...)
    return mean(chi_sims .> chi_obs)
end

"nsim_logP(gpT, gpC, gpNull, nsim) (src/hypotest_mll.jl:21-30): vector of (mll_null, mll_altv); leaves gpNull at the last draw (:6,10,14)"
function nsim_logP(gpT::B200GPE, gpC::B200GPE, gpNull::B200Null, nsim::Integer; seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing)
    Zm = zmat(Z, gpNull.nobs); ns = ndraws(Zm, nsim)
    mnull = Vector{Float64}(undef, ns); malt = Vector{Float64}(undef, ns); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm mnull malt gpNull check(gpT.ctx, ccall((:geordd_null_mll, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble},
                              Ref{Clonglong}, Ptr{Cdouble}, Ptr{Cdouble}),
                             gpNull.ptr, ns, seed, draw0, zptr(Zm), Inf, mnull, malt, cnt, gpNull.y, gpNull.alpha))
    gpNull.mll = mnull[end]
    return collect(zip(mnull, malt))
end

"boot_mlltest(gpT, gpC, nsim) (src/hypotest_mll.jl:32-45)"
function boot_mlltest(gpT::B200GPE, gpC::B200GPE, nsim::Integer; kwargs...)
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
    GC.@preserve X check(gpT.ctx, ccall((:geordd_invvar_obs, LIB), Cint,
                             (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
                             gpT.ptr, gpC.ptr, X, size(X, 2), Float64(nugget), τ, V, p))
    return p[]
end

"nsim_invvar_pval_uncalib(gpT, gpC, Xb, nsim; nugget) (src/hypotest_invvar.jl:48-63)"
function nsim_invvar_pval_uncalib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Integer; nugget=1e-10,
                                  seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing, gpNull::Union{Nothing,B200Null}=nothing)
    gpNull === nothing ? (gpNull = make_null(gpT, gpC; Xb=Xb, nugget=nugget)) : set_sentinels!(gpNull, Xb, nugget)
    Zm = zmat(Z, gpNull.nobs); ns = ndraws(Zm, nsim)
    pv = Vector{Float64}(undef, ns); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm pv check(gpT.ctx, ccall((:geordd_null_invvar, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                             gpNull.ptr, ns, seed, draw0, zptr(Zm), -1.0, pv, cnt))
    return pv
end

"boot_invvar(gpT, gpC, Xb, nsim; nugget) (src/hypotest_invvar.jl:65-69)"
function boot_invvar(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Integer; nugget=1e-10, kwargs...)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=nugget)
    pval_obs = pval_invvar_uncalib(gpNull)
    pval_sims = nsim_invvar_pval_uncalib(gpT, gpC, Xb, nsim; nugget=nugget, gpNull=gpNull, kwargs...)
    return mean(pval_sims .< pval_obs)
end

"pval_invvar_calib(gpT, gpC, Xb; nugget) (src/hypotest_invvar.jl:74-105)"
function pval_invvar_calib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix; nugget=1e-10)
    X = Matrix{Float64}(Xb); p = Ref{Cdouble}(0.0)
    GC.@preserve X check(gpT.ctx, ccall((:geordd_invvar_calib, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Cdouble}),
                                        gpT.ptr, gpC.ptr, X, size(X, 2), Float64(nugget), p))
    return p[]
end

"nsim_invvar_calib(gpT, gpC, Xb, nsim) (src/hypotest_invvar.jl:138-160): the 3-argument pval_invvar_calib of every null draw"
function nsim_invvar_calib(gpT::B200GPE, gpC::B200GPE, Xb::AbstractMatrix, nsim::Integer;
                           seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=1e-10)
    Zm = zmat(Z, gpNull.nobs); ns = ndraws(Zm, nsim)
    pv = Vector{Float64}(undef, ns)
    GC.@preserve Zm pv check(gpT.ctx, ccall((:geordd_null_invvar_calib, LIB), Cint,
                             (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}),
                             gpNull.ptr, ns, seed, draw0, zptr(Zm), pv))
    return pv
end

"""
    nsim_power(gpT, gpC, τ, Xb, chi_null, mll_null, pval_invvar_null, nsim)   (src/power.jl:46-73)

`compat_calib_bug=false` (default) evaluates the analytic calibration with all four covariance terms -- the formula that
reproduces the published size / power table of notebooks/Mississippi_sharp_null_sims.ipynb.  `true` reproduces the 7-argument
`pval_invvar_calib` AS SHIPPED (src/hypotest_invvar.jl:125-128: the last three terms are a separate, discarded expression).
"""
function nsim_power(gpT::B200GPE, gpC::B200GPE, τ::Real, Xb::AbstractMatrix, chi_null::Vector{Float64},
                    mll_null::Vector{Float64}, pval_invvar_null::Vector{Float64}, nsim::Integer;
                    seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing, compat_calib_bug::Bool=false)
    gpNull = make_null(gpT, gpC; Xb=Xb, nugget=0.0)
    Zm = zmat(Z, gpNull.nobs); ns = ndraws(Zm, nsim)
    out = Matrix{Float64}(undef, 5, ns)
    GC.@preserve Zm chi_null mll_null pval_invvar_null out check(gpT.ctx, ccall((:geordd_power, LIB), Cint,
                             (Ptr{Cvoid}, Cdouble, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                              Ptr{Cdouble}, Clong, Cint, Ptr{Cdouble}),
                             gpNull.ptr, Float64(τ), ns, seed, draw0, zptr(Zm), chi_null, mll_null, pval_invvar_null,
                             length(chi_null), compat_calib_bug ? 1 : 0, out))
    return [Tuple(out[:, s]) for s in 1:ns]
end

# ---- 6. border geometry on the device (src/border_projection.jl) ------------------------------------------------------
function line_vertices(border)
    lines = border isa LibGEOS.MultiLineString ? GeoInterface.coordinates(border) : [GeoInterface.coordinates(border)]
    V = Matrix{Float64}(undef, 2, sum(length, lines)); starts = Cint[]; v = 0
    for (k, line) in enumerate(lines)
        k > 1 && push!(starts, v)
        for p in line
            v += 1; V[1, v] = p[1]; V[2, v] = p[2]
        end
    end
    return V, starts
end

"projection_points(gp, border, maxdist) (src/border_projection.jl:4-22): one device call instead of a LibGEOS call per unit"
function projection_points(gp::B200GPE, border, maxdist::Real)
    V, starts = line_vertices(border)
    n = size(gp.x, 2); Xp = Matrix{Float64}(undef, 2, n); d = Vector{Float64}(undef, n)
    GC.@preserve gp V starts Xp d check(gp.ctx, ccall((:geordd_projection_points, LIB), Cint,
                            (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
                            gp.ctx.ptr, gp.x, n, V, size(V, 2), starts, length(starts), Xp, d))
    return Xp[:, d .<= maxdist]
end

"proj_estimator(gpT, gpC, border, maxdist) (src/border_projection.jl:24-31)"
function proj_estimator(gpT::B200GPE, gpC::B200GPE, border, maxdist::Real)
    X∂ = [projection_points(gpT, border, maxdist) projection_points(gpC, border, maxdist)]
    μpost, Σpost = cliff_face(gpT, gpC, X∂)
    return unweighted_mean(μpost, Σpost)
end

"`LibGEOS.within(point, region)` for every column of P (the grid filter of infinite_proj_sentinels, src/border_projection.jl:44-102)"
function within_region(ctx::Context, P::Matrix{Float64}, region)
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
    GC.@preserve P R starts inside check(ctx, ccall((:geordd_points_in_region, LIB), Cint,
                         (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{UInt8}),
                         ctx.ptr, P, n, R, size(R, 2), starts, length(starts), inside))
    return inside .!= 0
end

"Opt-in (default off): mll quadratic forms as |L⁻¹(y-m)|² / z'z, 6 m² instead of 16 m² flop per draw, equal to rounding"
set_mll_forward_only!(ctx::Context, on::Bool) =
    check(ctx, ccall((:geordd_ctx_set_mll_forward_only, LIB), Cint, (Ptr{Cvoid}, Cint), ctx.ptr, on ? 1 : 0))

# ---- 7. joint fit with covariates: the numerics of MultiGPCovars on the device -----------------------------------------
mutable struct B200MultiGP            # geordd_multigp: D'D cached on the device, one ABI call per objective / gradient
    ctx::Context
    ptr::Ptr{Cvoid}
    n::Int
    p::Int
end

function B200MultiGP(D::Matrix{Float64}, X::Matrix{Float64}, offsets::Vector{Cint}, y::Vector{Float64}; ctx::Context=default_context())
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = GC.@preserve D X offsets y ccall((:geordd_multigp_create, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Cint, Ptr{Cint}, Cint, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                   ctx.ptr, D, size(D, 1), X, size(X, 1), offsets, length(offsets) - 1, y, out)
    check(ctx, rc)
    h = B200MultiGP(ctx, out[], size(D, 2), size(D, 1))
    finalizer(m -> ccall((:geordd_multigp_destroy, LIB), Cint, (Ptr{Cvoid},), m.ptr), h)
    return h
end

const MGP_CACHE = IdDict{Any,B200MultiGP}()
function device_handle(mgpcv::MultiGPCovars)
    h = get(MGP_CACHE, mgpcv, nothing)
    h !== nothing && return h
    X = hcat([mgpcv.mgp[key].x for key in mgpcv.groupKeys]...)          # groups contiguous, in groupKeys order
    offsets = Cint[0]
    for key in mgpcv.groupKeys
        push!(offsets, offsets[end] + length(mgpcv.groupIndices[key]))
    end
    h = B200MultiGP(Matrix{Float64}(mgpcv.D), Matrix{Float64}(X), offsets, Vector{Float64}(mgpcv.y))
    MGP_CACHE[mgpcv] = h
    return h
end

"update_mll!(mgpcv) (src/GPrealisationsCovars.jl:110-125)"
function update_mll!(mgpcv::MultiGPCovars)
    h = device_handle(mgpcv)
    ck = Ref(ckernel(mgpcv.kernel, mgpcv.mean)); m = Ref{Cdouble}(0.0); jr = Ref{Cint}(0)
    alpha = Vector{Float64}(undef, mgpcv.nobs)
    GC.@preserve alpha check(h.ctx, ccall((:geordd_multigp_mll, LIB), Cint,
                     (Ptr{Cvoid}, Ref{CKernel}, Cdouble, Cdouble, Ptr{Cdouble}, Ref{Cdouble}, Ref{Cint}),
                     h.ptr, ck, mgpcv.βkern.ℓ2, Float64(mgpcv.logNoise), alpha, m, jr))
    mgpcv.alpha = alpha
    mgpcv.mll = m[]
    return mgpcv.mll
end

# The device reports kernel gradients as [log ell, log sigma, (log sigma_const)]; get_params(kernel) may order the Const
# first and hides the parameters of fix()-ed sub-kernels.
kernel_mask(k::FixedKernel) = fill(false, length(kernel_mask(k.kernel)))
kernel_mask(k::Const) = [true]
kernel_mask(k::Kernel) = [true, true]
function kernel_mask(k::SumKernel)
    ma, mb = kernel_mask(k.kleft), kernel_mask(k.kright)
    return unfix(k.kleft) isa Const ? vcat(mb, ma) : vcat(ma, mb)      # device order: spatial first
end
function kernel_grads(k::Kernel, g::Vector{Float64})
    mask = kernel_mask(k)
    kept = g[1:length(mask)][mask]
    if k isa SumKernel && unfix(k.kleft) isa Const && length(kept) == 3
        kept = kept[[3, 1, 2]]                                          # get_params order with the Const on the left
    end
    return kept
end

"update_mll_and_dmll!(mgpcv, buf; noise, domean, kern, beta) (src/GPrealisationsCovars.jl:145-193); buf is unused"
function update_mll_and_dmll!(mgpcv::MultiGPCovars, buf=nothing; noise::Bool=true, domean::Bool=true, kern::Bool=true, beta::Bool=true)
    h = device_handle(mgpcv)
    ck = Ref(ckernel(mgpcv.kernel, mgpcv.mean)); m = Ref{Cdouble}(0.0); nd = Ref{Cint}(0)
    g = zeros(Float64, 8)
    GC.@preserve g check(h.ctx, ccall((:geordd_multigp_mll_grad, LIB), Cint,
                     (Ptr{Cvoid}, Ref{CKernel}, Cdouble, Cdouble, Cint, Cint, Cint, Cint, Ref{Cdouble}, Ptr{Cdouble}, Ref{Cint}),
                     h.ptr, ck, mgpcv.βkern.ℓ2, Float64(mgpcv.logNoise), noise ? 1 : 0, domean ? 1 : 0, kern ? 1 : 0, beta ? 1 : 0,
                     m, g, nd))
    mgpcv.mll = m[]
    g = g[1:nd[]]
    out = Float64[]; i = 1
    if noise; push!(out, g[i]); i += 1; end
    if domean && num_params(mgpcv.mean) > 0 && meanconst(mgpcv.mean) != 0.0; push!(out, g[i]); i += 1
    elseif domean; append!(out, zeros(num_params(mgpcv.mean))); end
    if kern
        nk = length(kernel_mask(mgpcv.kernel))
        append!(out, kernel_grads(mgpcv.kernel, g[i:i + nk - 1])); i += nk
    end
    if beta; push!(out, g[i]); end
    mgpcv.dmll = out
    return mgpcv.dmll
end
initialise_mll!(mgpcv::MultiGPCovars) = update_mll!(mgpcv)             # src/GPrealisationsCovars.jl:127-142

"postmean_β(mgpcv) (src/GPrealisationsCovars.jl:335-344)"
function postmean_β(mgpcv::MultiGPCovars)
    h = device_handle(mgpcv)
    ck = Ref(ckernel(mgpcv.kernel, mgpcv.mean)); β = Vector{Float64}(undef, mgpcv.p)
    GC.@preserve β check(h.ctx, ccall((:geordd_multigp_postmean_beta, LIB), Cint,
                     (Ptr{Cvoid}, Ref{CKernel}, Cdouble, Cdouble, Ptr{Cdouble}),
                     h.ptr, ck, mgpcv.βkern.ℓ2, Float64(mgpcv.logNoise), β))
    return β
end
# optimize!(mgpcv; ...) (src/GPrealisationsCovars.jl:224-279) needs no change: its closures call update_mll! and
# update_mll_and_dmll!, i.e. every Optim.jl objective / gradient evaluation becomes one ABI call.

"""
    B200Realisations(rdict, groupKeys, kernel, logNoise, mean)   (src/region_gp_fit.jl:23-37 -> src/GPrealisations.jl:10-19)

One fitted GP per region sharing (kernel, logNoise): all regions are fitted by ONE device call (`fit_many`).
"""
struct B200Realisations{KEY}
    groupKeys::Vector{KEY}
    mgp::Dict{KEY,B200GPE}
    kernel::Kernel
    logNoise::Float64
    mean::Mean
    nobs::Int
    mll::Float64
end
function B200Realisations(rdict::Dict{KEY,RegionData}, kernel::Kernel, logNoise, mean::Mean=MeanZero();
                          groupKeys::Vector{KEY}=collect(keys(rdict))) where {KEY}
    all(length(rdict[key].y) > 0 for key in groupKeys) || throw("empty group")
    gps = fit_many([rdict[key].x for key in groupKeys], [rdict[key].y for key in groupKeys], mean, kernel, logNoise)
    d = Dict{KEY,B200GPE}(zip(groupKeys, gps))
    return B200Realisations{KEY}(groupKeys, d, kernel, lognoise(logNoise), mean, sum(g.nobs for g in gps), sum(g.mll for g in gps))
end
Base.getindex(r::B200Realisations, key) = r.mgp[key]
Base.keys(r::B200Realisations) = r.groupKeys
Base.values(r::B200Realisations) = values(r.mgp)

# ---- 8. placebo drivers (src/hypotest_chi2.jl:62-76, src/hypotest_mll.jl:47-60, src/hypotest_invvar.jl:162-175) --------
# The O(n) geometry (shift_for_even_split, left_points, placebo_sentinels; src/placebo_geometry.jl) stays in Julia.
function placebo_split(angle::Float64, X::AbstractMatrix, Y::AbstractVector, kern::Kernel, m::Mean, logNoise)
    Xm = Matrix{Float64}(X)
    shift = GeoRDD.shift_for_even_split(angle, Xm)
    left = GeoRDD.left_points(angle, shift, Xm)
    gl, gr = fit_many([Xm[:, left], Xm[:, .!left]], [Y[left], Y[.!left]], m, kern, logNoise)
    return shift, gl, gr
end
function placebo_chi(angle::Float64, X::AbstractMatrix, Y::AbstractVector, kern::Kernel, m::Mean, logNoise, nsim::Integer; kwargs...)
    shift, gl, gr = placebo_split(angle, X, Y, kern, m, logNoise)
    return boot_chi2test(gl, gr, GeoRDD.placebo_sentinels(angle, shift, X, 100), nsim; kwargs...)
end
function placebo_mll(angle::Float64, X::AbstractMatrix, Y::AbstractVector, kern::Kernel, m::Mean, logNoise, nsim::Integer; kwargs...)
    _, gl, gr = placebo_split(angle, X, Y, kern, m, logNoise)
    return boot_mlltest(gl, gr, nsim; kwargs...)
end
function placebo_invvar(angle::Float64, X::AbstractMatrix, Y::AbstractVector, kern::Kernel, m::Mean, logNoise)
    shift, gl, gr = placebo_split(angle, X, Y, kern, m, logNoise)
    return pval_invvar_calib(gl, gr, GeoRDD.placebo_sentinels(angle, shift, X, 100))
end

# ---- 9. multi-GPU from one Julia task: the group layer of the C ABI (SURVEY §8e) ---------------------------------------
# Fits replicated on every device, null draws sharded by index, tallies summed with one NCCL all-reduce INSIDE the library:
# no MPI, no Distributed.  Results equal the single-GPU ones bit for bit.
mutable struct DeviceGroup
    ptr::Ptr{Cvoid}
    ndev::Int
end
function DeviceGroup(devices::Vector{<:Integer})
    out = Ref{Ptr{Cvoid}}(C_NULL); devs = Cint.(devices)
    rc = ccall((:geordd_group_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}), out, length(devs), devs)
    rc == 0 || error("geordd_b200: geordd_group_create failed (rc=$rc): no usable devices, or NCCL missing for more than one")
    g = DeviceGroup(out[], length(devs))
    finalizer(x -> ccall((:geordd_group_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), g)
    return g
end
function gcheck(g::DeviceGroup, rc::Integer)
    rc == 0 && return
    msg = unsafe_string(ccall((:geordd_group_last_error, LIB), Cstring, (Ptr{Cvoid},), g.ptr))
    rc > 0 && throw(LinearAlgebra.PosDefException(rc))
    rc == -1 && throw(ArgumentError(msg))
    error("geordd_b200 group ($rc): $msg")
end
context(g::DeviceGroup, i::Integer=0) = Context(ccall((:geordd_group_ctx, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), g.ptr, i), false)

mutable struct GroupGPE
    group::DeviceGroup
    ptr::Ptr{Cvoid}
    rep0::B200GPE                      # device 0's replica (borrowed) for the single-GPU entry points
end
function GroupGPE(g::DeviceGroup, x::AbstractMatrix, y::AbstractVector, mean::Mean, kernel::Kernel, logNoise)
    X = Matrix{Float64}(x); Y = Vector{Float64}(y); dim, n = size(X)
    ck = Ref(ckernel(kernel, mean)); out = Ref{Ptr{Cvoid}}(C_NULL)
    alpha = Vector{Float64}(undef, n); m = Ref{Cdouble}(0.0); jr = Ref{Cint}(0)
    rc = GC.@preserve X Y alpha ccall((:geordd_group_gp_fit, LIB), Cint,
                   (Ptr{Cvoid}, Ref{CKernel}, Ptr{Cdouble}, Cint, Cint, Ptr{Cdouble}, Cdouble, Ref{Ptr{Cvoid}}, Ptr{Cdouble},
                    Ref{Cdouble}, Ref{Cint}), g.ptr, ck, X, dim, n, Y, lognoise(logNoise), out, alpha, m, jr)
    gcheck(g, rc)
    rep = ccall((:geordd_group_gp_replica, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), out[], 0)
    h = GroupGPE(g, out[], B200GPE(context(g, 0), rep, X, Y, mean, kernel, lognoise(logNoise), alpha, m[], n, dim, false))
    finalizer(x -> ccall((:geordd_group_gp_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

mutable struct GroupNull
    group::DeviceGroup
    ptr::Ptr{Cvoid}
    mll::Float64
    nobs::Int
end
function GroupNull(gpT::GroupGPE, gpC::GroupGPE; Xb::Union{Nothing,AbstractMatrix}=nothing, nugget::Real=0.0)
    out = Ref{Ptr{Cvoid}}(C_NULL); jr = Ref{Cint}(0); m = Ref{Cdouble}(0.0)
    X = Xb === nothing ? Matrix{Float64}(undef, 2, 0) : Matrix{Float64}(Xb)
    rc = GC.@preserve X ccall((:geordd_group_null_prepare, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Cint, Cdouble, Ref{Ptr{Cvoid}}, Ref{Cint}, Ref{Cdouble}),
                   gpT.ptr, gpC.ptr, size(X, 2) == 0 ? C_NULL : pointer(X), size(X, 2), Float64(nugget), out, jr, m)
    gcheck(gpT.group, rc)
    h = GroupNull(gpT.group, out[], m[], gpT.rep0.nobs + gpC.rep0.nobs)
    finalizer(x -> ccall((:geordd_group_null_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

function boot_mlltest(gpT::GroupGPE, gpC::GroupGPE, nsim::Integer; seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing)
    h = GroupNull(gpT, gpC)
    Zm = zmat(Z, h.nobs); ns = ndraws(Zm, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm gcheck(h.group, ccall((:geordd_group_null_mll, LIB), Cint,
                   (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Clonglong}),
                   h.ptr, ns, seed, draw0, zptr(Zm), gpT.rep0.mll + gpC.rep0.mll - h.mll, C_NULL, C_NULL, cnt))
    return cnt[] / ns
end
function boot_chi2test(gpT::GroupGPE, gpC::GroupGPE, Xb::AbstractMatrix, nsim::Integer; seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing)
    h = GroupNull(gpT, gpC; Xb=Xb, nugget=0.0)
    Zm = zmat(Z, h.nobs); ns = ndraws(Zm, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm gcheck(h.group, ccall((:geordd_group_null_chi2, LIB), Cint,
                   (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                   h.ptr, ns, seed, draw0, zptr(Zm), chistat(gpT.rep0, gpC.rep0, Xb), C_NULL, cnt))
    return cnt[] / ns
end
function boot_invvar(gpT::GroupGPE, gpC::GroupGPE, Xb::AbstractMatrix, nsim::Integer; nugget=1e-10, seed::Integer=rand(UInt64),
                     draw0::Integer=0, Z=nothing)
    h = GroupNull(gpT, gpC; Xb=Xb, nugget=nugget)
    Zm = zmat(Z, h.nobs); ns = ndraws(Zm, nsim); cnt = Ref{Clonglong}(0)
    GC.@preserve Zm gcheck(h.group, ccall((:geordd_group_null_invvar, LIB), Cint,
                   (Ptr{Cvoid}, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Clonglong}),
                   h.ptr, ns, seed, draw0, zptr(Zm), pval_invvar_uncalib(gpT.rep0, gpC.rep0, Xb; nugget=nugget), C_NULL, cnt))
    return cnt[] / ns
end
function nsim_power(gpT::GroupGPE, gpC::GroupGPE, τ::Real, Xb::AbstractMatrix, chi_null::Vector{Float64}, mll_null::Vector{Float64},
                    pval_invvar_null::Vector{Float64}, nsim::Integer; seed::Integer=rand(UInt64), draw0::Integer=0, Z=nothing,
                    compat_calib_bug::Bool=false)
    h = GroupNull(gpT, gpC; Xb=Xb, nugget=0.0)
    Zm = zmat(Z, h.nobs); ns = ndraws(Zm, nsim); out = Matrix{Float64}(undef, 5, ns)
    GC.@preserve Zm chi_null mll_null pval_invvar_null out gcheck(h.group, ccall((:geordd_group_power, LIB), Cint,
                   (Ptr{Cvoid}, Cdouble, Clong, Culonglong, Culonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Clong,
                    Cint, Ptr{Cdouble}), h.ptr, Float64(τ), ns, seed, draw0, zptr(Zm), chi_null, mll_null, pval_invvar_null,
                   length(chi_null), compat_calib_bug ? 1 : 0, out))
    return [Tuple(out[:, s]) for s in 1:ns]
end

# ---- 10. stock signatures: the bodies of the reference's entry points become device calls -------------------------------
# (`@eval GeoRDD ...` re-defines the methods in the enclosing module for the STOCK argument types; set
#  ENV["GEORDD_B200_OVERRIDE"] = "0" before loading GeoRDD to keep the stock CPU bodies.)
if get(ENV, "GEORDD_B200_OVERRIDE", "1") != "0"
    B = GeoRDDB200
    @eval GeoRDD begin
        cliff_face(gpT::GPE, gpC::GPE, sentinels::AbstractMatrix) = $B.cliff_face($B.B200GPE(gpT), $B.B200GPE(gpC), sentinels)
        chistat(gpT::GPE, gpC::GPE, Xb::AbstractMatrix) = $B.chistat($B.B200GPE(gpT), $B.B200GPE(gpC), Xb)
        boot_chi2test(gpT::GPE, gpC::GPE, Xb::AbstractMatrix, nsim::Int; kw...) =
            $B.boot_chi2test($B.B200GPE(gpT), $B.B200GPE(gpC), Xb, nsim; kw...)
        nsim_chi(gpT::GPE, gpC::GPE, gpNull::GPE, Xb::AbstractMatrix, nsim::Int; kw...) =
            $B.nsim_chi($B.B200GPE(gpT), $B.B200GPE(gpC), $B.make_null($B.B200GPE(gpT), $B.B200GPE(gpC)), Xb, nsim; kw...)
        boot_mlltest(gpT::GPE, gpC::GPE, nsim::Int; kw...) = $B.boot_mlltest($B.B200GPE(gpT), $B.B200GPE(gpC), nsim; kw...)
        function nsim_logP(gpT::GPE, gpC::GPE, gpNull::GPE, nsim::Int; kw...)
            dT, dC = $B.B200GPE(gpT), $B.B200GPE(gpC)
            dN = $B.make_null(dT, dC)
            sims = $B.nsim_logP(dT, dC, dN, nsim; kw...)
            gpNull.y = dN.y; gpNull.alpha = dN.alpha; gpNull.mll = dN.mll      # the reference mutates gpNull (:6,10,14)
            return sims
        end
        pval_invvar_uncalib(gpT::GPE, gpC::GPE, Xb::AbstractMatrix; nugget=1e-10) =
            $B.pval_invvar_uncalib($B.B200GPE(gpT), $B.B200GPE(gpC), Xb; nugget=nugget)
        nsim_invvar_pval_uncalib(gpT::GPE, gpC::GPE, Xb::AbstractMatrix, nsim::Int; nugget=1e-10, kw...) =
            $B.nsim_invvar_pval_uncalib($B.B200GPE(gpT), $B.B200GPE(gpC), Xb, nsim; nugget=nugget, kw...)
        boot_invvar(gpT::GPE, gpC::GPE, Xb::AbstractMatrix, nsim::Int; nugget=1e-10, kw...) =
            $B.boot_invvar($B.B200GPE(gpT), $B.B200GPE(gpC), Xb, nsim; nugget=nugget, kw...)
        pval_invvar_calib(gpT::GPE, gpC::GPE, Xb::Matrix; nugget=1e-10) =
            $B.pval_invvar_calib($B.B200GPE(gpT), $B.B200GPE(gpC), Xb; nugget=nugget)
        nsim_invvar_calib(gpT::GPE, gpC::GPE, Xb::AbstractMatrix, nsim::Int; kw...) =
            $B.nsim_invvar_calib($B.B200GPE(gpT), $B.B200GPE(gpC), Xb, nsim; kw...)
        nsim_power(gpT::GPE, gpC::GPE, τ::Float64, Xb::AbstractMatrix, chi_null::Vector{Float64}, mll_null::Vector{Float64},
                   pval_invvar_null::Vector{Float64}, nsim::Int; kw...) =
            $B.nsim_power($B.B200GPE(gpT), $B.B200GPE(gpC), τ, Xb, chi_null, mll_null, pval_invvar_null, nsim; kw...)
        inverse_variance(μ::AbstractVector, Σ::AbstractMatrix; nugget=1e-10) = $B.inverse_variance(μ, Σ; nugget=nugget)
        weight_at_units(gp::GPE, sentinels::AbstractMatrix, weights::AbstractVector) = $B.weight_at_units($B.B200GPE(gp), sentinels, weights)
        projection_points(gp::GPE, border, maxdist::Float64) = $B.projection_points($B.B200GPE(gp), border, maxdist)
        update_mll!(mgpcv::MultiGPCovars) = $B.update_mll!(mgpcv)
        initialise_mll!(mgpcv::MultiGPCovars) = $B.initialise_mll!(mgpcv)
        update_mll_and_dmll!(mgpcv::MultiGPCovars, buf::Matrix{Float64}; kw...) = $B.update_mll_and_dmll!(mgpcv, buf; kw...)
        postmean_β(mgpcv::MultiGPCovars) = $B.postmean_β(mgpcv)
        placebo_chi(angle::Float64, X::AbstractMatrix, Y::Vector, kern::Kernel, mean::Mean, logNoise::Float64, nsim::Int; kw...) =
            $B.placebo_chi(angle, X, Y, kern, mean, logNoise, nsim; kw...)
        placebo_mll(angle::Float64, X::Matrix, Y::Vector, kern::Kernel, m::Mean, logNoise::Float64, nsim::Int; kw...) =
            $B.placebo_mll(angle, X, Y, kern, m, logNoise, nsim; kw...)
        placebo_invvar(angle::Float64, X::AbstractMatrix, Y::Vector, kern::Kernel, mean::Mean, logNoise::Float64) =
            $B.placebo_invvar(angle, X, Y, kern, mean, logNoise)
    end
end

end # module

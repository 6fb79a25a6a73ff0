# Julia side of the drop-in boundary (see INTEGRATION.md): every reference function on the spread / gradient
# path gets a method that calls libwannier_b200.so (include/wannier_b200.h) through `ccall`.
#
# NOT executed in this repository's CI: no Julia runtime exists in the build or GPU images.  Each `ccall`
# below has a one-to-one twin in wannier.jl_b200/lib.py (ctypes, same symbol, same argument order), and
# those twins are what tests/ exercise on the GPU.
#
# Two ways to use it:
#   1. explicit:   c = B200Cache(model);  omega(c, U);  max_localize(model, c)
#   2. drop-in:    WannierB200Ext.enable!()  — afterwards the reference's own entry points
#                  `omega(model)`, `omega_grad(stencil, M, U)`, `center(model)`, `max_localize(model)`,
#                  `disentangle(model)`, `opt_rotate(model)`, `transform_gauge(M, kpb_k, U)`,
#                  `position_op(model)`, `compute_berry_connection_kspace(model)`, `U_to_X_Y`, the MagModel
#                  functions and the internal hooks `compute_MU_UtMU!` / `omega!` / `omega_grad!`
#                  run on the GPU (methods are re-defined inside `Wannier` at run time; undoing it is a
#                  restart).  One device context per `model.overlaps` object is kept in `CACHES`.
#                  `center!(r, UtMU, stencil)` keeps the reference's host loop: its input is the host-side
#                  `Cache.UtMU`, which `compute_MU_UtMU!` above fills from the device.
module WannierB200Ext
using Wannier
import LinearAlgebra
using Wannier: KspaceStencil, Model, Spread, Vec3, AbstractPenalty, SpreadPenalty, CenterSpreadPenalty,
               n_bands, n_wannier, n_kpoints, n_bvectors

const LIB = get(ENV, "WANNIER_B200_LIB", "libwannier_b200.so")
const DEVICES = Ref{Vector{Cint}}(Cint[0])      # more than one entry: k-points sharded over these GPUs

struct B200Error <: Exception; code::Cint; msg::String; end
lasterror(ptr) = unsafe_string(ccall((:wb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ptr))
function check(ctx::Ptr{Cvoid}, st::Cint)
    st == 0 && return nothing
    throw(B200Error(st, lasterror(ctx)))
end

# ------------------------------------------------------------------------------------------------------
# packing: Vector{Vector{Matrix}} / Vector{Matrix} <-> the contiguous column-major arrays of the C-ABI
# ------------------------------------------------------------------------------------------------------
"gauges as one (nb, nw, nk) array — the layout max_localize.jl:68-71 hands to Optim"
pack_gauges(U::AbstractVector{<:AbstractMatrix}) = ComplexF64[U[ik][m, n] for m in axes(U[1], 1), n in axes(U[1], 2), ik in eachindex(U)]
pack_gauges(U::Array{ComplexF64,3}) = U
unpack_gauges(U::Array{ComplexF64,3}) = [U[:, :, ik] for ik in axes(U, 3)]

function pack_model(stencil::KspaceStencil, M)
    nk, nn = length(M), length(M[1]); nb = size(M[1][1], 1)
    Mp = Array{ComplexF64,4}(undef, nb, nb, nn, nk)
    kpb = Matrix{Int32}(undef, nn, nk); bvec = Array{Float64,3}(undef, 3, nn, nk)
    for ik in 1:nk, ib in 1:nn
        Mp[:, :, ib, ik] .= M[ik][ib]
        ikpb = stencil.kpb_k[ik][ib]
        kpb[ib, ik] = ikpb - 1                                     # 0-based
        bvec[:, ib, ik] .= stencil.recip_lattice *                 # src/spread.jl:293
            (stencil.kpoints[ikpb] + stencil.kpb_G[ik][ib] - stencil.kpoints[ik])
    end
    return Mp, kpb, bvec, Vector{Float64}(stencil.bweights), nb, nk, nn
end

"page-lock an array that is reused across calls (Julia arrays are pageable; see wb200_host_register)"
pin!(A::Array) = (ccall((:wb200_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), A, sizeof(A)); A)
unpin!(A::Array) = (ccall((:wb200_host_unregister, LIB), Cint, (Ptr{Cvoid},), A); A)

"""
The one-shot calls (`omega(stencil, M, U)`, `omega_grad(...)`, `max_localize(model)`: a `Cache` per call in the
reference, src/spread.jl:378, 497; a device context per call here) are cheap because the library keeps the device /
page-locked blocks and streams of destroyed contexts.  `trim_pools()` hands them back to the driver (e.g. before
CUDA.jl needs the memory); `pool_bytes(device)` reports what is kept.
"""
trim_pools() = ccall((:wb200_trim_pools, LIB), Cvoid, ())
pool_bytes(device::Integer=0) = ccall((:wb200_pool_stats, LIB), Clonglong, (Cint, Ptr{Clonglong}, Ptr{Clonglong}), device, C_NULL, C_NULL)

# ------------------------------------------------------------------------------------------------------
# device contexts: replace `Cache(model)` (src/spread.jl:127-162)
# ------------------------------------------------------------------------------------------------------
abstract type AbstractB200Cache end

"one GPU (wb200_create)"
mutable struct B200Cache <: AbstractB200Cache
    ptr::Ptr{Cvoid}
    nb::Int; nw::Int; nk::Int; nn::Int
end
"several GPUs behind one call (wb200_multi_create): k-slabs, one host thread per slab inside the library"
mutable struct B200MultiCache <: AbstractB200Cache
    ptr::Ptr{Cvoid}
    nb::Int; nw::Int; nk::Int; nn::Int
end

function B200Cache(stencil::KspaceStencil, M, nw::Integer; device::Integer=0, flags::Integer=0)
    Mp, kpb, bvec, wb, nb, nk, nn = pack_model(stencil, M)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:wb200_create, LIB), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Cint, Cint,
         Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}, Cuint),
        ref, device, nb, nw, nk, 0, nk, nn, kpb, bvec, wb, Mp, flags)
    st == 0 || throw(B200Error(st, lasterror(C_NULL)))
    c = B200Cache(ref[], nb, nw, nk, nn)
    finalizer(x -> ccall((:wb200_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), c)
    return c
end
function B200MultiCache(stencil::KspaceStencil, M, nw::Integer; devices::Vector{<:Integer}, flags::Integer=0)
    Mp, kpb, bvec, wb, nb, nk, nn = pack_model(stencil, M)
    ref = Ref{Ptr{Cvoid}}(C_NULL); dv = Vector{Cint}(devices)
    st = ccall((:wb200_multi_create, LIB), Cint,
        (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint, Cint, Cint, Cint, Cint,
         Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}, Cuint),
        ref, dv, length(dv), nb, nw, nk, nn, kpb, bvec, wb, Mp, flags)
    st == 0 || throw(B200Error(st, unsafe_string(ccall((:wb200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    c = B200MultiCache(ref[], nb, nw, nk, nn)
    finalizer(x -> ccall((:wb200_multi_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), c)
    return c
end
B200Cache(model::Model; kw...) = B200Cache(model.kstencil, model.overlaps, n_wannier(model); kw...)
B200MultiCache(model::Model; kw...) = B200MultiCache(model.kstencil, model.overlaps, n_wannier(model); kw...)
function check(c::B200MultiCache, st::Cint)
    st == 0 && return nothing
    throw(B200Error(st, unsafe_string(ccall((:wb200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), c.ptr))))
end
check(c::B200Cache, st::Cint) = check(c.ptr, st)

"one context per overlaps object (the registry behind `enable!`)"
const CACHES = IdDict{Any,AbstractB200Cache}()
function cache_for(stencil::KspaceStencil, M, nw::Integer)
    get!(CACHES, M) do
        length(DEVICES[]) > 1 ? B200MultiCache(stencil, M, nw; devices=DEVICES[]) :
                                B200Cache(stencil, M, nw; device=DEVICES[][1])
    end
end
cache_for(model::Model) = cache_for(model.kstencil, model.overlaps, n_wannier(model))
release!(M) = (delete!(CACHES, M); nothing)          # the finalizer frees the device memory

# ------------------------------------------------------------------------------------------------------
# penalties and frozen bands (src/spread.jl:203-227, src/localization/disentangle.jl:481-501)
# ------------------------------------------------------------------------------------------------------
function set_penalty!(c::B200Cache, p::CenterSpreadPenalty)
    check(c, ccall((:wb200_set_penalty, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, reduce(hcat, p.r₀), p.λ))
end
set_penalty!(c::B200Cache, ::AbstractPenalty) =
    check(c, ccall((:wb200_set_penalty, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, C_NULL, 0.0))
function set_penalty!(c::B200MultiCache, p::CenterSpreadPenalty)
    check(c, ccall((:wb200_multi_set_penalty, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, reduce(hcat, p.r₀), p.λ))
end
set_penalty!(c::B200MultiCache, ::AbstractPenalty) =
    check(c, ccall((:wb200_multi_set_penalty, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, C_NULL, 0.0))

"model.frozen_bands :: Vector{BitVector} as the (nb, nk) byte mask of wb200_set_frozen"
function frozen_mask(c::AbstractB200Cache, frozen_bands::AbstractVector)
    mask = Matrix{UInt8}(undef, c.nb, c.nk)
    for ik in 1:c.nk
        mask[:, ik] .= frozen_bands[ik]
    end
    return mask
end
set_frozen!(c::B200Cache, fb::AbstractVector) =
    check(c, ccall((:wb200_set_frozen, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), c.ptr, frozen_mask(c, fb)))
set_frozen!(c::B200MultiCache, fb::AbstractVector) =
    check(c, ccall((:wb200_multi_set_frozen, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), c.ptr, frozen_mask(c, fb)))

# ------------------------------------------------------------------------------------------------------
# evaluations
# ------------------------------------------------------------------------------------------------------
"fg!(F, G, U) with Optim's only_fg! convention (src/localization/max_localize.jl:13-23); U, G (nb, nw, nk)"
function fg!(c::B200Cache, F, G, U::Array{ComplexF64,3})
    f = Ref{Float64}(0.0)
    check(c, ccall((:wb200_fg, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
        c.ptr, U, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G))
    return F === nothing ? nothing : f[]
end
function fg!(c::B200MultiCache, F, G, U::Array{ComplexF64,3})
    f = Ref{Float64}(0.0)
    check(c, ccall((:wb200_multi_fg, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
        c.ptr, U, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G))
    return F === nothing ? nothing : f[]
end

"get_fg!_maxloc(p, model) (max_localize.jl:10-28)"
function get_fg!_maxloc(p::AbstractPenalty, c::AbstractB200Cache)
    return function (F, G, U)
        set_penalty!(c, p)          # the penalty travels with every call in the reference
        fg!(c, F, G, U)
    end
end

"omega(stencil, M, U) (src/spread.jl:377-381) -> Spread"
function spread(c::AbstractB200Cache, U::Array{ComplexF64,3})
    out = zeros(5); ω = zeros(c.nw); r = zeros(3, c.nw); mn = Ref{Float64}(0.0)
    st = c isa B200Cache ?
        ccall((:wb200_spread, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              c.ptr, U, out, ω, r, mn) :
        ccall((:wb200_multi_spread, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
              c.ptr, U, out, ω, r, mn)
    check(c, st)
    return Spread(out[1], out[2], out[3], out[4], out[5], ω, [Vec3(r[:, n]) for n in 1:c.nw])
end

"omega_grad(stencil, M, U) (src/spread.jl:496-501) -> G (nb, nw, nk)"
function spread_grad(p::AbstractPenalty, c::AbstractB200Cache, U::Array{ComplexF64,3})
    G = similar(U)
    set_penalty!(c, p)
    fg!(c, nothing, G, U)
    return G
end

"center(stencil, M, U) (src/spread.jl:560-564)"
function centers(c::B200Cache, U::Array{ComplexF64,3})
    r = zeros(3, c.nw)
    check(c, ccall((:wb200_center, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}), c.ptr, U, r))
    return [Vec3(r[:, n]) for n in 1:c.nw]
end
centers(c::B200MultiCache, U::Array{ComplexF64,3}) = spread(c, U).r

"transform_gauge(M, kpb_k, U) (src/localization/gauge.jl:83-98) and compute_MU_UtMU! (src/spread.jl:164-198)"
function rotate_overlaps(c::B200Cache, U::Array{ComplexF64,3}; want_N::Bool=true, want_MU::Bool=false)
    N = want_N ? Array{ComplexF64,4}(undef, c.nw, c.nw, c.nn, c.nk) : nothing
    MU = want_MU ? Array{ComplexF64,4}(undef, c.nb, c.nw, c.nn, c.nk) : nothing
    check(c, ccall((:wb200_rotate_overlaps, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
        c.ptr, U, want_N ? N : C_NULL, want_MU ? MU : C_NULL))
    return N, MU
end

"fg!(Ω, G, XY) of get_fg!_disentangle (src/localization/disentangle.jl:586-614); XY, G (nw^2 + nb*nw, nk)"
function fg_xy!(c::AbstractB200Cache, F, G, XY::Matrix{ComplexF64})
    f = Ref{Float64}(0.0)
    st = c isa B200Cache ?
        ccall((:wb200_fg_xy, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
              c.ptr, XY, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G) :
        ccall((:wb200_multi_fg_xy, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
              c.ptr, XY, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G)
    check(c, st)
    return F === nothing ? nothing : f[]
end

"X_Y_to_XY(U_to_X_Y(U, frozen)...) on the device (disentangle.jl:369-402, 454-466); set_frozen! first"
function u_to_xy(c::B200Cache, U::Array{ComplexF64,3})
    XY = Matrix{ComplexF64}(undef, c.nw^2 + c.nb * c.nw, c.nk)
    check(c, ccall((:wb200_u_to_xy, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), c.ptr, U, XY))
    return XY
end

"Optim.Stiefel_SVD retract! / project_tangent! for a stack of matrices (nrow, ncol, count)"
function retract!(c::B200Cache, X::Array{ComplexF64,3})
    check(c, ccall((:wb200_retract, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Cint, Cint, Cint), c.ptr, X, size(X, 1), size(X, 2), size(X, 3)))
    return X
end
function project_tangent!(c::B200Cache, G::Array{ComplexF64,3}, X::Array{ComplexF64,3})
    check(c, ccall((:wb200_project_tangent, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint, Cint),
        c.ptr, X, G, size(X, 1), size(X, 2), size(X, 3)))
    return G
end

# ------------------------------------------------------------------------------------------------------
# drivers with the optimiser loop on the device
# ------------------------------------------------------------------------------------------------------
"max_localize(p, model; ...) (src/localization/max_localize.jl:44-101)"
function max_localize(p::AbstractPenalty, model::Model, c::AbstractB200Cache; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3)
    n_bands(model) != n_wannier(model) && error("n_bands != n_wann, run instead disentanglement?")
    set_penalty!(c, p)
    U = pack_gauges(model.gauges)
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    st = c isa B200Cache ?
        ccall((:wb200_max_localize, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
              c.ptr, U, f_tol, g_tol, max_iter, history_size, f, it, nfg) :
        ccall((:wb200_multi_max_localize, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
              c.ptr, U, f_tol, g_tol, max_iter, history_size, f, it, nfg)
    check(c, st)
    @info "max_localize on the device" Ω = f[] iterations = it[] evaluations = nfg[]
    return unpack_gauges(U)                                        # Vector{Matrix}, as max_localize.jl:94-98
end

"disentangle(p, model; ...) (src/localization/disentangle.jl:631-728)"
function disentangle(p::AbstractPenalty, model::Model, c::AbstractB200Cache; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3)
    set_penalty!(c, p)
    set_frozen!(c, model.frozen_bands)
    XY = c isa B200Cache ? u_to_xy(c, pack_gauges(model.gauges)) :                       # disentangle.jl:666-670
         Wannier.X_Y_to_XY(Wannier.U_to_X_Y(model.gauges, model.frozen_bands)...)
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    st = c isa B200Cache ?
        ccall((:wb200_disentangle, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
              c.ptr, XY, f_tol, g_tol, max_iter, history_size, f, it, nfg) :
        ccall((:wb200_multi_disentangle, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
              c.ptr, XY, f_tol, g_tol, max_iter, history_size, f, it, nfg)
    check(c, st)
    Xmin, Ymin = Wannier.XY_to_X_Y(XY, c.nb, c.nw)                 # disentangle.jl:721-722
    return Wannier.X_Y_to_U(Xmin, Ymin)
end

"opt_rotate(model; ...) (src/localization/opt_rotate.jl:97-154): one global rotation W"
function opt_rotate(model::Model; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3)
    n_bands(model) != n_wannier(model) && error("n_bands != n_wann, run instead disentanglement?")
    c0 = cache_for(model)
    N, _ = rotate_overlaps(c0, pack_gauges(model.gauges))          # opt_rotate.jl:113-116: M rotated by the gauges
    nw = c0.nw
    Mrot = [[N[:, :, ib, ik] for ib in 1:c0.nn] for ik in 1:c0.nk]
    c = B200Cache(model.kstencil, Mrot, nw; device=DEVICES[][1])
    W = Matrix{ComplexF64}(LinearAlgebra.I, nw, nw)
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    check(c, ccall((:wb200_opt_rotate, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
        c.ptr, W, f_tol, g_tol, max_iter, history_size, f, it, nfg))
    return W
end

"omega_local(stencil, M, U) (src/spread.jl:522-548)"
function omega_local(c::B200Cache, U::Array{ComplexF64,3})
    loc = zeros(c.nk)
    check(c, ccall((:wb200_omega_local, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}), c.ptr, U, loc))
    return loc
end

"position_op(stencil, M, U) (src/spread.jl:674-716) -> Matrix{Vec3{ComplexF64}} (nw, nw)"
function position_op(c::B200Cache, U::Array{ComplexF64,3})
    R = Array{ComplexF64,3}(undef, 3, c.nw, c.nw)                  # component fastest: the memory image of Matrix{Vec3}
    check(c, ccall((:wb200_position_op, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}), c.ptr, U, R))
    return [Vec3(R[:, m, n]) for m in 1:c.nw, n in 1:c.nw]
end

"compute_berry_connection_kspace (src/spread.jl:747-801) -> Vector{Matrix{Vec3{ComplexF64}}}"
function berry_connection(c::B200Cache, U::Array{ComplexF64,3}; imlog_diag::Bool=true, force_hermiticity::Bool=true)
    A = Array{ComplexF64,4}(undef, 3, c.nw, c.nw, c.nk)
    check(c, ccall((:wb200_berry_connection, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Cint, Cint),
        c.ptr, U, A, imlog_diag, force_hermiticity))
    return [[Vec3(A[:, m, n, ik]) for m in 1:c.nw, n in 1:c.nw] for ik in 1:c.nk]
end

# ------------------------------------------------------------------------------------------------------
# MagModel (src/localization/coopt.jl)
# ------------------------------------------------------------------------------------------------------
mutable struct B200MagCache
    ptr::Ptr{Cvoid}
    up::B200Cache; dn::B200Cache
end
function B200MagCache(model::Wannier.MagModel; device::Integer=0)
    up = B200Cache(model.up; device); dn = B200Cache(model.dn; device)
    nb = up.nb; nk = up.nk
    Mud = ComplexF64[model.M[ik][m, l] for m in 1:nb, l in 1:nb, ik in 1:nk]
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:wb200_mag_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{ComplexF64}), ref, up.ptr, dn.ptr, Mud)
    st == 0 || throw(B200Error(st, unsafe_string(ccall((:wb200_mag_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    c = B200MagCache(ref[], up, dn)
    finalizer(x -> ccall((:wb200_mag_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), c)   # before up / dn: they are fields of c
    return c
end
function check(c::B200MagCache, st::Cint)
    st == 0 && return nothing
    throw(B200Error(st, unsafe_string(ccall((:wb200_mag_last_error, LIB), Cstring, (Ptr{Cvoid},), c.ptr))))
end
"overlap_updn(model, Uup, Udn) (coopt.jl:114-135) and omega_updn (:146-160)"
function overlap_updn(c::B200MagCache, Uup, Udn)
    Mw = Matrix{Float64}(undef, c.up.nw, c.up.nw); o = Ref{Float64}(0.0)
    check(c, ccall((:wb200_mag_overlap, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}),
        c.ptr, pack_gauges(Uup), pack_gauges(Udn), Mw, o))
    return Mw, o[]
end
"omega_updn_grad(model, Uup, Udn) (coopt.jl:222-229)"
function omega_updn_grad(c::B200MagCache, Uup, Udn)
    Gu = Array{ComplexF64,3}(undef, c.up.nb, c.up.nw, c.up.nk); Gd = similar(Gu)
    check(c, ccall((:wb200_mag_omega_updn_grad, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}),
        c.ptr, pack_gauges(Uup), pack_gauges(Udn), Gu, Gd))
    return unpack_gauges(Gu), unpack_gauges(Gd)
end
"(f, g!) of get_fg!_disentangle(model::MagModel, λ) (coopt.jl:252-321); XY (2 n_inner, nk)"
function get_fg!_disentangle(c::B200MagCache, λ::Real=1.0)
    f(XY) = (r = Ref{Float64}(0.0);
             check(c, ccall((:wb200_mag_fg_xy, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}),
                            c.ptr, XY, λ, r, C_NULL, C_NULL)); r[])
    g!(G, XY) = (check(c, ccall((:wb200_mag_fg_xy, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Ptr{Float64}, Ptr{ComplexF64}, Ptr{Float64}),
                                c.ptr, XY, λ, C_NULL, G, C_NULL)); nothing)
    return f, g!
end
"disentangle(model::MagModel, λ; ...) (coopt.jl:339-420) -> (Uup, Udn)"
function disentangle(model::Wannier.MagModel, λ::Real=1.0; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3, device::Integer=0)
    c = B200MagCache(model; device)
    set_frozen!(c.up, model.up.frozen_bands); set_frozen!(c.dn, model.dn.frozen_bands)
    XY = vcat(u_to_xy(c.up, pack_gauges(model.up.gauges)), u_to_xy(c.dn, pack_gauges(model.dn.gauges)))   # coopt.jl:369-374
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    check(c, ccall((:wb200_mag_disentangle, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
        c.ptr, XY, λ, f_tol, g_tol, max_iter, history_size, f, it, nfg))
    ni = c.up.nw^2 + c.up.nb * c.up.nw
    Xu, Yu = Wannier.XY_to_X_Y(XY[1:ni, :], c.up.nb, c.up.nw); Xd, Yd = Wannier.XY_to_X_Y(XY[(ni + 1):end, :], c.up.nb, c.up.nw)
    return Wannier.X_Y_to_U(Xu, Yu), Wannier.X_Y_to_U(Xd, Yd)
end

# ------------------------------------------------------------------------------------------------------
# drop-in: re-define the reference's own methods so existing user code runs on the GPU unchanged
# ------------------------------------------------------------------------------------------------------
function enable!(; devices::Vector{<:Integer}=[0])
    DEVICES[] = Vector{Cint}(devices)
    ext = @__MODULE__
    @eval Wannier begin
        const _B200 = $ext
        # ---- public API (src/spread.jl:370-381, 496-510, 560-564, 649-662) ----
        omega(bvectors::KspaceStencil, M, U) = _B200.spread(_B200.cache_for(bvectors, M, size(U isa Vector ? U[1] : U, 2)), _B200.pack_gauges(U))
        omega(bvectors::KspaceStencil, M, X, Y) = omega(bvectors, M, X_Y_to_U(X, Y))
        omega_grad(bvectors::KspaceStencil, M, U) = _B200.spread_grad(SpreadPenalty(), _B200.cache_for(bvectors, M, size(U isa Vector ? U[1] : U, 2)), _B200.pack_gauges(U))
        omega_grad(p::CenterSpreadPenalty, bvectors::KspaceStencil, M, U) = _B200.spread_grad(p, _B200.cache_for(bvectors, M, size(U isa Vector ? U[1] : U, 2)), _B200.pack_gauges(U))
        center(bvectors::KspaceStencil, M, U) = _B200.centers(_B200.cache_for(bvectors, M, size(U isa Vector ? U[1] : U, 2)), _B200.pack_gauges(U))
        transform_gauge(M::Vector, kpb_k::Vector, U::Vector) = begin
            st = _B200.identity_stencil(kpb_k)
            N, _ = _B200.rotate_overlaps(_B200.B200Cache(st, M, size(U[1], 2); device=_B200.DEVICES[][1]), _B200.pack_gauges(U))
            [[N[:, :, ib, ik] for ib in axes(N, 3)] for ik in axes(N, 4)]
        end
        omega_local(bvectors::KspaceStencil, M, U) = _B200.omega_local(_B200.cache_for(bvectors, M, size(U[1], 2)), _B200.pack_gauges(U))
        position_op(bvectors::KspaceStencil, M, U) = _B200.position_op(_B200.cache_for(bvectors, M, size(U[1], 2)), _B200.pack_gauges(U))
        compute_berry_connection_kspace(bvectors::KspaceStencil, M, U; kw...) =
            _B200.berry_connection(_B200.cache_for(bvectors, M, size(U[1], 2)), _B200.pack_gauges(U); kw...)
        # ---- the closure Optim calls and the drivers (max_localize.jl:10-101, disentangle.jl:586-728, opt_rotate.jl:97-154) ----
        get_fg!_maxloc(p::AbstractPenalty, model::Model) = _B200.get_fg!_maxloc(p, _B200.cache_for(model))
        max_localize(p::AbstractPenalty, model::Model; kw...) = _B200.max_localize(p, model, _B200.cache_for(model); kw...)
        function get_fg!_disentangle(p::AbstractPenalty, model::Model)
            c = _B200.cache_for(model); _B200.set_frozen!(c, model.frozen_bands)
            return (F, G, XY) -> (_B200.set_penalty!(c, p); _B200.fg_xy!(c, F, G, XY))
        end
        disentangle(p::AbstractPenalty, model::Model; kw...) = _B200.disentangle(p, model, _B200.cache_for(model); kw...)
        disentangle(model::MagModel, λ::Real=1.0; kw...) = _B200.disentangle(model, λ; kw...)
        opt_rotate(model::Model; kw...) = _B200.opt_rotate(model; kw...)
        # ---- internal hooks (src/spread.jl:164-198, 254-360, 398-481, 565-594): same side effects on `Cache` ----
        function compute_MU_UtMU!(cache::Cache, bvectors::KspaceStencil, M, U)
            c = _B200.cache_for(bvectors, M, n_wann(cache))
            N, MU = _B200.rotate_overlaps(c, _B200.pack_gauges(U); want_N=true, want_MU=true)
            for ik in axes(N, 4), ib in axes(N, 3)
                cache.UtMU[ik][ib] .= view(N, :, :, ib, ik); cache.MU[ik][ib] .= view(MU, :, :, ib, ik)
            end
            cache.U .= U isa Vector ? U : _B200.unpack_gauges(U)      # the gauges the cached products belong to
            return cache.MU, cache.UtMU
        end
        omega!(cache::Cache, bvectors::KspaceStencil, M) = begin
            Ω = _B200.spread(_B200.cache_for(bvectors, M, n_wann(cache)), _B200.pack_gauges(cache.U)); cache.r .= Ω.r; Ω
        end
        function omega_grad!(penalty::Function, cache::Cache, bvectors, M)
            p = _B200.penalty_of(penalty)
            cache.G .= _B200.spread_grad(p, _B200.cache_for(bvectors, M, n_wann(cache)), _B200.pack_gauges(cache.U))
            return cache.G
        end
        omega_grad!(p::CenterSpreadPenalty, cache::Cache, bvectors, M) =
            (cache.G .= _B200.spread_grad(p, _B200.cache_for(bvectors, M, n_wann(cache)), _B200.pack_gauges(cache.U)); cache.G)
        # U_to_X_Y (disentangle.jl:369-402) without a model: a Stiefel-only context
        function U_to_X_Y(U::AbstractVector{<:AbstractMatrix{T}}, frozen::Vector{BitVector}) where {T<:Complex}
            XY = _B200.u_to_xy_standalone(U, frozen)
            return XY_to_X_Y(XY, size(U[1], 1), size(U[1], 2))
        end
    end
    return nothing
end

"the identity closure `(r, _) -> r` of src/spread.jl:398 carries no penalty; a CenterSpreadPenalty reaches omega_grad! by type"
penalty_of(::Function) = SpreadPenalty()

"a stencil that carries only kpb_k (transform_gauge needs nothing else)"
function identity_stencil(kpb_k::Vector)
    nk = length(kpb_k); nn = length(kpb_k[1])
    KspaceStencil(Wannier.Mat3{Float64}(LinearAlgebra.I), (1, 1, nk), [Vec3(0.0, 0.0, 0.0) for _ in 1:nk],
                  [Vec3(0.0, 0.0, 0.0) for _ in 1:nn], ones(nn), kpb_k, [[Vec3(0, 0, 0) for _ in 1:nn] for _ in 1:nk])
end

function u_to_xy_standalone(U::AbstractVector, frozen::AbstractVector)
    nb, nw = size(U[1]); nk = length(U)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:wb200_create_stiefel, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint), ref, DEVICES[][1], nb, nw, nk)
    st == 0 || throw(B200Error(st, "wb200_create_stiefel failed"))
    c = B200Cache(ref[], nb, nw, nk, 0)
    try
        set_frozen!(c, frozen)
        return u_to_xy(c, pack_gauges(U))
    finally
        ccall((:wb200_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL
    end
end

end # module

# Julia side of the drop-in boundary (see INTEGRATION.md).  Not executed in this repository's CI:
# no Julia runtime exists in the build or GPU images; the identical calls are exercised through
# wannier.jl_b200/lib.py (ctypes) by tests/test_gpu_parity.py.
module WannierB200Ext
using Wannier
using Wannier: KspaceStencil, Model, Spread, n_bands, n_wannier, n_kpoints, n_bvectors

const LIB = get(ENV, "WANNIER_B200_LIB", "libwannier_b200.so")

struct B200Error <: Exception; code::Cint; msg::String; end
function check(ctx, st::Cint)
    st == 0 && return
    msg = unsafe_string(ccall((:wb200_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    throw(B200Error(st, msg))
end

"Device context: replaces `Cache(model)` (src/spread.jl:139-162)."
mutable struct B200Cache
    ptr::Ptr{Cvoid}
    nb::Int; nw::Int; nk::Int; nn::Int
end

function B200Cache(stencil::KspaceStencil, M::Vector{<:Vector{<:Matrix{ComplexF64}}}, nw::Int; device=0)
    nk, nn = length(M), length(M[1]); nb = size(M[1][1], 1)
    # pack once: M[ik][ib] (nb x nb col-major) -> one contiguous [nb, nb, nn, nk] array
    Mp = Array{ComplexF64,4}(undef, nb, nb, nn, nk)
    kpb = Matrix{Int32}(undef, nn, nk); bvec = Array{Float64,3}(undef, 3, nn, nk)
    for ik in 1:nk, ib in 1:nn
        Mp[:, :, ib, ik] .= M[ik][ib]
        ikpb = stencil.kpb_k[ik][ib]
        kpb[ib, ik] = ikpb - 1                                     # 0-based
        bvec[:, ib, ik] .= stencil.recip_lattice *                 # src/spread.jl:293
            (stencil.kpoints[ikpb] + stencil.kpb_G[ik][ib] - stencil.kpoints[ik])
    end
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:wb200_create, LIB), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Cint, Cint, Cint,
         Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{ComplexF64}, Cuint),
        ref, device, nb, nw, nk, 0, nk, nn, kpb, bvec, Vector{Float64}(stencil.bweights), Mp, 0)
    st == 0 || throw(B200Error(st, unsafe_string(ccall((:wb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    c = B200Cache(ref[], nb, nw, nk, nn)
    finalizer(x -> ccall((:wb200_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ptr), c)
    return c
end
B200Cache(model::Model; kw...) = B200Cache(model.kstencil, model.overlaps, n_wannier(model); kw...)

"fg!(F, G, U) of get_fg!_maxloc (src/localization/max_localize.jl:13-23); U, G :: Array{ComplexF64,3} (nb, nw, nk)."
function get_fg!_maxloc(c::B200Cache)
    return function fg!(F, G, U::Array{ComplexF64,3})
        f = Ref{Float64}(0.0)
        check(c.ptr, ccall((:wb200_fg, LIB), Cint,
            (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
            c.ptr, U, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G))
        return F === nothing ? nothing : f[]
    end
end

"omega(stencil, M, U) (src/spread.jl:377-381)"
function Wannier.omega(c::B200Cache, U::Array{ComplexF64,3})
    out = zeros(5); ω = zeros(c.nw); r = zeros(3, c.nw); mn = Ref{Float64}(0.0)
    check(c.ptr, ccall((:wb200_spread, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        c.ptr, U, out, ω, r, mn))
    return Spread(out[1], out[2], out[3], out[4], out[5], ω, [Wannier.Vec3(r[:, n]) for n in 1:c.nw])
end

"center(stencil, M, U) (src/spread.jl:560-564)"
function Wannier.center(c::B200Cache, U::Array{ComplexF64,3})
    r = zeros(3, c.nw)
    check(c.ptr, ccall((:wb200_center, LIB), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}), c.ptr, U, r))
    return [Wannier.Vec3(r[:, n]) for n in 1:c.nw]
end

"CenterSpreadPenalty (src/spread.jl:214-227)"
set_penalty!(c::B200Cache, p::Wannier.CenterSpreadPenalty) = check(c.ptr, ccall((:wb200_set_penalty, LIB),
    Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, reduce(hcat, p.r₀), p.λ))
set_penalty!(c::B200Cache, ::Wannier.SpreadPenalty) = check(c.ptr, ccall((:wb200_set_penalty, LIB),
    Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), c.ptr, C_NULL, 0.0))

"max_localize(model; ...) (src/localization/max_localize.jl:44-101) with the optimiser loop on the device"
function Wannier.max_localize(p, model::Model, c::B200Cache; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3)
    n_bands(model) == n_wannier(model) || error("n_bands != n_wann, run instead disentanglement?")
    set_penalty!(c, p)
    U = cat(model.gauges...; dims=3)                    # (nb, nw, nk), as max_localize.jl:68-71
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    check(c.ptr, ccall((:wb200_max_localize, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
        c.ptr, U, f_tol, g_tol, max_iter, history_size, f, it, nfg))
    return [U[:, :, ik] for ik in 1:size(U, 3)]          # Vector{Matrix}, as max_localize.jl:94-98
end

"frozen bands of the disentanglement (model.frozen_bands :: Vector{BitVector}) as the [nb, nk] byte mask the library wants"
function set_frozen!(c::B200Cache, frozen_bands::AbstractVector)
    mask = Matrix{UInt8}(undef, c.nb, c.nk)
    for ik in 1:c.nk
        mask[:, ik] .= frozen_bands[ik]
    end
    check(c.ptr, ccall((:wb200_set_frozen, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}), c.ptr, mask))
end

"fg!(Ω, G, XY) of get_fg!_disentangle (src/localization/disentangle.jl:586-614); XY, G :: Matrix{ComplexF64} (nw^2 + nb*nw, nk)."
function get_fg!_disentangle(c::B200Cache)
    return function fg!(F, G, XY::Matrix{ComplexF64})
        f = Ref{Float64}(0.0)
        check(c.ptr, ccall((:wb200_fg_xy, LIB), Cint,
            (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{Float64}, Ptr{ComplexF64}),
            c.ptr, XY, F === nothing ? C_NULL : f, G === nothing ? C_NULL : G))
        return F === nothing ? nothing : f[]
    end
end

"disentangle(model; ...) (src/localization/disentangle.jl:631-728) with the optimiser loop on the device"
function Wannier.disentangle(p, model::Model, c::B200Cache; f_tol=1e-7, g_tol=1e-5, max_iter=200, history_size=3)
    set_penalty!(c, p)
    set_frozen!(c, model.frozen_bands)
    X0, Y0 = Wannier.U_to_X_Y(model.gauges, model.frozen_bands)   # disentangle.jl:666 (host, one-off)
    XY = Wannier.X_Y_to_XY(X0, Y0)                                 # :670, (nw^2 + nb*nw, nk)
    f = Ref{Float64}(0.0); it = Ref{Cint}(0); nfg = Ref{Cint}(0)
    check(c.ptr, ccall((:wb200_disentangle, LIB), Cint,
        (Ptr{Cvoid}, Ptr{ComplexF64}, Float64, Float64, Cint, Cint, Ptr{Float64}, Ptr{Cint}, Ptr{Cint}),
        c.ptr, XY, f_tol, g_tol, max_iter, history_size, f, it, nfg))
    Xmin, Ymin = Wannier.XY_to_X_Y(XY, c.nb, c.nw)                 # disentangle.jl:721-722
    return Wannier.X_Y_to_U(Xmin, Ymin)
end
end # module

# P2DEB200.jl — the binding a P2DE.jl maintainer adds to use libp2de_b200.so for the hot path.
#
# UNTESTED HERE: Julia is not installed on the build image.  The same C ABI is exercised from
# Python/ctypes (p2de_b200/lib.py, tests/).  Everything except `rhs!` / `SSP33!` (setup,
# post-processing, examples) stays the reference's own code.
module P2DEB200

using P2DE
using StaticArrays

const lib = get(ENV, "P2DE_B200_LIB", "libp2de_b200.so")

# ---- include/p2de_b200.h, mirrored field by field ------------------------------------------
struct Config
    abi_version::Int32; dim::Int32; N::Int32; basis::Int32
    K::Int64; Kx::Int32; Ky::Int32
    Nq::Int32; Nfp::Int32; Nh::Int32; Np::Int32
    rhs_type::Int32; vol_flux::Int32; surf_flux_low::Int32; surf_flux_high::Int32
    proj_limiter::Int32; limiter::Int32; bound::Int32; shockcapture::Int32
    keep_diagnostics::Int32; device::Int32; lgl_projection_roundtrip::Int32; _reserved::Int32
    hennemann_a::Float64; hennemann_c::Float64; bound_beta::Float64
    gamma::Float64; POSTOL::Float64; ZEROTOL::Float64; zeta::Float64; eta::Float64
    CFL::Float64; dt0::Float64; t0::Float64; T::Float64
end
struct Operators
    Srsh_db::NTuple{2,Ptr{Float64}}; Srs0::NTuple{2,Ptr{Float64}}; Brs::NTuple{2,Ptr{Float64}}
    Vf::Ptr{Float64}; Vf_low::Ptr{Float64}; MinvVhT::Ptr{Float64}; MinvVfT::Ptr{Float64}
    VDM_inv::Ptr{Float64}; wq::Ptr{Float64}; fq2q::Ptr{Int64}
end
struct Geometry
    J::Ptr{Float64}; Jq::Ptr{Float64}; GJh::NTuple{4,Ptr{Float64}}
    uniform::Int32; _pad::Int32; J_const::Float64; GJ_const::NTuple{4,Float64}
end
struct BC
    mapP::Ptr{Int64}; periodic_x::Int32; periodic_y::Int32
    nI::Int64; mapI::Ptr{Int64}; Ival::Ptr{Float64}; nO::Int64; mapO::Ptr{Int64}
end

code(::LobattoCollocation) = 0; code(::GaussCollocation) = 1
code(::LowOrderPositivity) = 0; code(::FluxDiffRHS) = 1; code(::LimitedDG) = 2
code(::ChandrashekarFlux) = 0; code(::CentralFlux) = 1
code(::ChandrashekarOnProjectedVal) = 0; code(::LaxFriedrichsOnNodalVal) = 1; code(::LaxFriedrichsOnProjectedVal) = 2
code(::NoEntropyProjectionLimiter) = 0; code(::NodewiseScaledExtrapolation) = 1
code(::NoRHSLimiter) = 0; code(::ZhangShuLimiter) = 1; code(::SubcellLimiter) = 2
code(::PositivityBound) = 0; code(::PositivityAndMinEntropyBound) = 1; code(::PositivityAndRelaxedMinEntropyBound) = 2
code(::PositivityAndCellEntropyBound) = 3; code(::PositivityAndRelaxedCellEntropyBound) = 4
code(::TVDBound) = 5; code(::TVDAndMinEntropyBound) = 6; code(::TVDAndRelaxedMinEntropyBound) = 7
code(::TVDAndCellEntropyBound) = 8; code(::TVDAndRelaxedCellEntropyBound) = 9
bound_beta(b) = hasproperty(b, :beta) ? Float64(b.beta) : 0.0     # *RelaxedCellEntropyBound(beta), Solver.jl:52-63
code(::NoShockCapture) = 0; code(::HennemannShockCapture) = 1

"`State` of the reference plus the device handle; `rhs!`/`SSP33!` dispatch on it."
mutable struct B200State{S}
    state::S
    handle::Ptr{Cvoid}
    keep::Vector{Any}          # arrays the structs point into (GC roots)
end

check(h, rc) = rc == 0 || error("libp2de_b200 error $rc: " *
    unsafe_string(ccall((:p2de_last_error, lib), Cstring, (Ptr{Cvoid},), h)))

function B200State(state, solver, state_param; device=-1, keep_diagnostics=false)
    (; param, discrete_data) = solver
    (; sizes, geom, ops) = discrete_data
    r = param.rhs
    volf, lowf, highf = r isa LimitedDG ? (r.high_order_volume_flux, r.low_order_surface_flux, r.high_order_surface_flux) :
                        r isa FluxDiffRHS ? (r.volume_flux, LaxFriedrichsOnNodalVal(), r.surface_flux) :
                        (ChandrashekarFlux(), r.surface_flux, LaxFriedrichsOnProjectedVal())
    lim = param.rhs_limiter
    sc = P2DE.shockcapture(lim)
    Kx, Ky = param.K isa Tuple ? param.K : (param.K, 1)
    tp, gc, lp = param.timestepping_param, param.global_constants, param.limiting_param
    cfg = Config(1, sizes.Nd, param.N, code(param.approximation_basis), sizes.K, Kx, Ky,
        sizes.Nq, sizes.Nfp, sizes.Nh, sizes.Np, code(r), code(volf), code(lowf), code(highf),
        code(param.entropyproj_limiter), code(lim), lim isa NoRHSLimiter ? 0 : code(P2DE.bound(lim)), code(sc),
        keep_diagnostics, device, 0, 0,
        sc isa HennemannShockCapture ? sc.a : 0.5, sc isa HennemannShockCapture ? sc.c : 1.8,
        lim isa NoRHSLimiter ? 0.0 : bound_beta(P2DE.bound(lim)),
        P2DE.get_gamma(param.equation), gc.POSTOL, gc.ZEROTOL, lp.zeta, lp.eta, tp.CFL, tp.dt0, tp.t0, tp.T)
    bc = state_param.bcdata
    dense = [Matrix(s) for s in ops.Srs0]
    Bd = [collect(Float64, [B[i, i] for i in 1:sizes.Nfp]) for B in ops.Brs]
    Ival = isempty(bc.Ival) ? Float64[] : collect(reinterpret(Float64, bc.Ival))
    keep = Any[ops.Srsh_db..., dense..., Bd..., ops.Vf, ops.Vf_low, ops.MinvVhT, ops.MinvVfT, ops.VDM_inv,
               ops.wq, ops.fq2q, geom.J, geom.Jq, geom.GJh..., bc.mapP, bc.mapI, bc.mapO, Ival]
    p2(xs) = (pointer(xs[1]), length(xs) > 1 ? pointer(xs[2]) : Ptr{Float64}(0))
    o = Operators(p2(ops.Srsh_db), p2(dense), p2(Bd), pointer(ops.Vf), pointer(ops.Vf_low),
        pointer(ops.MinvVhT), pointer(ops.MinvVfT), pointer(ops.VDM_inv), pointer(ops.wq), pointer(ops.fq2q))
    gj = ntuple(i -> i <= length(geom.GJh) ? pointer(geom.GJh[i]) : Ptr{Float64}(0), 4)
    g = Geometry(pointer(geom.J), pointer(geom.Jq), gj, 0, 0, 0.0, (0.0, 0.0, 0.0, 0.0))
    b = BC(pointer(bc.mapP), 0, 0, length(bc.mapI), pointer(bc.mapI), pointer(Ival), length(bc.mapO), pointer(bc.mapO))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep begin
        rc = ccall((:p2de_create, lib), Int32, (Ref{Config}, Ref{Operators}, Ref{Geometry}, Ref{BC}, Ref{Ptr{Cvoid}}),
            cfg, o, g, b, h)
        check(C_NULL, rc)
        Uq = state.preallocation.Uq
        check(h[], ccall((:p2de_set_state, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}), h[], pointer(reinterpret(Float64, Uq))))
    end
    s = B200State(state, h[], keep)
    finalizer(x -> ccall((:p2de_destroy, lib), Int32, (Ptr{Cvoid},), x.handle), s)
    return s
end

const FIELDS = (rhsU=1, rhsH=2, rhsL=3, L=4, L_local=5, theta=6, theta_local=7)

"Copy an observable of src/common/types/State.jl:1-26 back into the reference's own array."
function fetch!(s::B200State, name::Symbol)
    A = getproperty(s.state.preallocation, name)
    flat = reinterpret(Float64, vec(A))
    GC.@preserve A check(s.handle, ccall((:p2de_get_field, lib), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64),
        s.handle, name === :Uq ? 0 : FIELDS[name], pointer(flat), length(flat)))
    return A
end

# rhs!(state, solver, state_param, time_param) -> dt            (src/dg/rhs/rhs.jl:5-13)
function P2DE.rhs!(s::B200State, solver, state_param, time_param)
    dt = Ref{Float64}(0.0)
    check(s.handle, ccall((:p2de_rhs, lib), Int32, (Ptr{Cvoid}, Float64, Float64, Int32, Ref{Float64}),
        s.handle, time_param.t, time_param.dt, time_param.nstage, dt))
    return dt[]
end

# SSP33!(state, solver, state_param) -> DataHistory               (src/timestepping/SSPRK33.jl:1-62)
function P2DE.SSP33!(s::B200State, solver, state_param)
    (; t0, T) = solver.param.timestepping_param
    (; output_interval) = solver.param.postprocessing_param
    Nc = P2DE.num_components(solver)
    Uhist, Lhist, thetahist, thist, dthist = [], [], [], [], []
    t, i = t0, 1
    while t < T
        dt = Ref{Float64}(0.0)
        check(s.handle, ccall((:p2de_ssp33_step, lib), Int32, (Ptr{Cvoid}, Float64, Ref{Float64}), s.handle, t, dt))
        t += dt[]; i += 1
        push!(dthist, dt[])
        if mod(i, output_interval) == 0 || abs(t - T) < 1e-10
            push!(thist, t)
            push!(Uhist, copy(fetch!(s, :Uq))); push!(Lhist, copy(fetch!(s, :L))); push!(thetahist, copy(fetch!(s, :theta)))
        end
    end
    return P2DE.DataHistory{Nc}(Uhist, Lhist, thetahist, thist, dthist)
end

# check_conservation(state, solver)                              (src/dg/utils.jl:1-12), reduced on the device
function P2DE.check_conservation(s::B200State, solver)
    out = Ref{Float64}(0.0)
    check(s.handle, ccall((:p2de_reduce, lib), Int32, (Ptr{Cvoid}, Int32, Ref{Float64}), s.handle, Int32(0), out))
    return out[]
end

# calculate_error(state, solver, exact_sol) -> ErrorData          (src/dg/postprocess.jl:1-46): the caller's exact_sol
# callback is evaluated on the host at every node exactly as the reference does, the weighted sums are formed on the device
function P2DE.calculate_error(s::B200State, solver, exact_sol)
    (; K, Nq, Nc) = solver.discrete_data.sizes
    T = solver.param.timestepping_param.T
    exact = [P2DE.exact_solution(P2DE.equation(solver), i, k, T, solver.md, exact_sol) for i = 1:Nq, k = 1:K]
    sums = zeros(Float64, Nc, 6)      # C: double[6][Nc] = L1err, L2err, Linferr, L1exact, L2exact, Linfexact
    flat = reinterpret(Float64, vec(exact))
    GC.@preserve exact check(s.handle, ccall((:p2de_calculate_error, lib), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}),
        s.handle, pointer(flat), pointer(sums)))
    L1, L2, Linf = 0.0, 0.0, 0.0
    for c = 1:Nc
        if sums[c, 6] > 1e-14
            L1 += sums[c, 1] / sums[c, 4]; L2 += sqrt(sums[c, 2]) / sqrt(sums[c, 5]); Linf += sums[c, 3] / sums[c, 6]
        end
    end
    return P2DE.ErrorData(L1, L2, Linf)
end

end # module

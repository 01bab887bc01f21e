# UNEXECUTED in this repository's environment (no Julia in the image); kept in sync with INTEGRATION.md section 2.
# lto_b200.jl -- ccall binding of liblto_b200.so (include/lto_b200.h)
module LtoB200
const lib = get(ENV, "LTO_B200_LIB", "liblto_b200.so")

mutable struct DirectParams            # == lto_direct_params (7 doubles, 4 int32)
    MU::Cdouble; DU::Cdouble; TU::Cdouble; Isp::Cdouble; g0::Cdouble; default_mass::Cdouble; tol::Cdouble
    mode::Int32; err_norm::Int32; max_attempts::Int32; kernel::Int32
    DirectParams() = (p = new(); ccall((:lto_direct_params_default, lib), Cvoid, (Ref{DirectParams},), p); p)
end
mutable struct IndirectParams          # == lto_indirect_params (12 doubles, 4 int32)
    MU::Cdouble; DU::Cdouble; TU::Cdouble; thrustLimit::Cdouble; mass::Cdouble; time_direction::Cdouble
    p::Cdouble; rho::Cdouble; Isp::Cdouble; g0::Cdouble; reltol::Cdouble; abstol::Cdouble
    controller::Int32; err_norm::Int32; max_attempts::Int32; kernel::Int32
    IndirectParams() = (p = new(); ccall((:lto_indirect_params_default, lib), Cvoid, (Ref{IndirectParams},), p); p)
end

const handle = Ref{Ptr{Cvoid}}(C_NULL)
function init(devices::Vector{Int32} = Int32[0])
    rc = ccall((:lto_init_devices, lib), Cint, (Cint, Ptr{Int32}, Ref{Ptr{Cvoid}}), length(devices), devices, handle)
    rc == 0 || error("lto_init_devices: ", unsafe_string(ccall((:lto_last_error, lib), Cstring, (Ptr{Cvoid},), C_NULL)))
end
check(rc) = rc == 0 || error("liblto_b200: ", unsafe_string(ccall((:lto_last_error, lib), Cstring, (Ptr{Cvoid},), handle[])))

# ---- pinned (page-locked) result arrays.  The device-to-host copy of the Jacobian / STM blocks is most of a call's bytes; into a
# pinned array it runs asynchronously at the PCIe rate and overlaps the kernels, into an ordinary (pageable) Julia array the driver
# stages it and the pipeline serialises (bench.py `e2e.pageable`).  Blocks are recycled through a small pool (cudaHostAlloc costs
# milliseconds); a finalizer gives the block back when the array is collected.
const _pool = Dict{Int, Vector{Ptr{Cvoid}}}()
const _pool_lock = ReentrantLock()
_size_class(n) = (c = 4096; while c < n; c <<= 1; end; c)
function pinned_array(::Type{T}, dims::Int...) where {T}
    nbytes = sizeof(T) * prod(dims)
    nbytes == 0 && return zeros(T, dims...)
    sz = _size_class(nbytes)
    ptr = lock(_pool_lock) do
        v = get(_pool, sz, nothing)
        (v === nothing || isempty(v)) ? C_NULL : pop!(v)
    end
    if ptr == C_NULL
        ptr = ccall((:lto_host_alloc, lib), Ptr{Cvoid}, (Csize_t,), sz)
        ptr == C_NULL && return zeros(T, dims...)          # pinned memory exhausted: a pageable array is still correct
    end
    a = unsafe_wrap(Array, Ptr{T}(ptr), dims; own = false)
    finalizer(a) do _
        lock(_pool_lock) do
            push!(get!(_pool, sz, Ptr{Cvoid}[]), ptr)
        end
    end
    a                                                       # contents unspecified: every caller below has the library overwrite all of it
end

# ---- direct: replaces defectCalc (multiShoot_CRTBP_direct.jl:66-109)
function defectCalc(X_all::Matrix{Float64}, u_all::Matrix{Float64}, t_TU::Vector{Float64}, nstate, n_nodes, nsteps, Isp, MU, DU, TU)
    p = DirectParams(); p.MU, p.DU, p.TU, p.Isp = MU, DU, TU, Isp
    defect = pinned_array(Float64, nstate, n_nodes - 1); errors = pinned_array(Float64, n_nodes - 1); status = pinned_array(Int32, n_nodes - 1)
    GC.@preserve X_all u_all t_TU defect errors status check(ccall((:lto_direct_defect_traj, lib), Cint,
        (Ptr{Cvoid}, Ref{DirectParams}, Int64, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        handle[], p, 1, n_nodes, nstate, nsteps, X_all, u_all, t_TU, defect, errors, status))
    (defect, errors)
end

# ---- direct: replaces jacobianCalc (:111-166); `defect`, `pert` become unused (variational equations)
function jacobianCalc(X_all, u_all, t_TU, nstate, n_nodes, nsteps, Isp, MU, DU, TU)
    p = DirectParams(); p.MU, p.DU, p.TU, p.Isp = MU, DU, TU, Isp
    nvar = 2 * (nstate + 3)
    defect = pinned_array(Float64, nstate, n_nodes - 1); errors = pinned_array(Float64, n_nodes - 1); status = pinned_array(Int32, n_nodes - 1)
    blocks = pinned_array(Float64, nstate, nvar, n_nodes - 1)              # block i == Jac_temp[(i-1)n+1 : i n, :]  (:139-140)
    GC.@preserve X_all u_all t_TU defect errors status blocks check(ccall((:lto_direct_defect_jac_traj, lib), Cint,
        (Ptr{Cvoid}, Ref{DirectParams}, Int64, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
        handle[], p, 1, n_nodes, nstate, nsteps, X_all, u_all, t_TU, defect, errors, status, blocks))
    Jac_full = zeros(nstate * (n_nodes - 1), n_nodes * (nstate + 3))                      # :146
    for i = 1:(n_nodes - 1)                                                               # :147-162, unchanged
        rows = ((i - 1) * nstate + 1):(i * nstate)
        Jac_full[rows, ((i - 1) * nstate + 1):((i + 1) * nstate)] = blocks[:, 1:(2 * nstate), i]
        c0 = nstate * n_nodes + 3 * (i - 1)
        Jac_full[rows, (c0 + 1):(c0 + 6)] = blocks[:, (2 * nstate + 1):end, i]
    end
    Jac_full
end

# ---- indirect: replaces defectCalc (multiShoot_CRTBP_indirect.jl:63-90) and jacobianCalc (:93-146)
function iparams(params)
    (MU, DU, TU, thrustLimit, mass, td, pp, rho) = params                                 # :260
    p = IndirectParams(); p.MU, p.DU, p.TU = MU, DU, TU
    p.thrustLimit, p.mass, p.time_direction, p.p, p.rho = thrustLimit, mass, td, pp, rho
    p
end
function defectCalc_indirect(XC_all::Matrix{Float64}, t_TU::Vector{Float64}, nstate, n_nodes, params)
    m = 2 * nstate; defect = pinned_array(Float64, m, n_nodes - 1); status = pinned_array(Int32, n_nodes - 1)
    GC.@preserve XC_all t_TU defect status check(ccall((:lto_indirect_defect_traj, lib), Cint,
        (Ptr{Cvoid}, Ref{IndirectParams}, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
        handle[], iparams(params), 1, n_nodes, m, XC_all, t_TU, C_NULL, C_NULL, defect, status, C_NULL))
    (defect, zeros(n_nodes - 1))                                                          # errors == 0 (:85)
end
function jacobianCalc_indirect(XC_all, t_TU, nstate, n_nodes, params)
    m = 2 * nstate; defect = pinned_array(Float64, m, n_nodes - 1); status = pinned_array(Int32, n_nodes - 1); phi = pinned_array(Float64, m, m, n_nodes - 1)
    GC.@preserve XC_all t_TU defect status phi check(ccall((:lto_indirect_defect_jac_traj, lib), Cint,
        (Ptr{Cvoid}, Ref{IndirectParams}, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        handle[], iparams(params), 1, n_nodes, m, XC_all, t_TU, C_NULL, C_NULL, defect, status, C_NULL, phi))
    Jac_full = zeros(m * (n_nodes - 1), n_nodes * m)                                      # :127
    for i = 1:(n_nodes - 1)                                                               # :128-138
        r = ((i - 1) * m + 1):(i * m)
        Jac_full[r, ((i - 1) * m + 1):(i * m)] = phi[:, :, i]
        for k = 1:m; Jac_full[r[k], i * m + k] = -1.0; end                                # hcat(Phi_i, -I) without LinearAlgebra's `I`
    end
    Jac_full[:, 1:nstate] .= 0.0; Jac_full[:, (end - 2 * nstate + 1):(end - nstate)] .= 0.0   # :141-142
    Jac_full
end
# ---- densify (HelperFunctions.jl:51-101): every dense time is its own propagation (node i, t_TU[i]) -> t, all of them in ONE call
# (pairs form of lto_indirect_defect, x_target = NULL returns x(t1)); the reference evaluates Vern8's dense-output interpolant instead.
function densify(XC_all::Matrix{Float64}, t_TU::Vector{Float64}, params, n_desired)
    m = size(XC_all, 1); N = size(XC_all, 2)
    t_dense = collect(LinRange(t_TU[1], t_TU[end], n_desired))
    seg = [searchsortedlast(t_TU, t) for t in t_dense]
    keep = seg .<= N - 1                                                                  # t < t_TU[end] (:64)
    seg = vcat(seg[keep], N - 1); t1 = vcat(t_dense[keep], t_TU[end])                     # + the last segment's end state (:94-97)
    x0 = XC_all[:, seg]; t0 = t_TU[seg]; n = length(seg)
    xend = pinned_array(Float64, m, n); status = pinned_array(Int32, n)
    GC.@preserve x0 t0 t1 xend status check(ccall((:lto_indirect_defect, lib), Cint,
        (Ptr{Cvoid}, Ref{IndirectParams}, Int64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}),
        handle[], iparams(params), n, m, x0, t0, t1, C_NULL, C_NULL, C_NULL, xend, status, C_NULL))
    (xend, t_dense)
end

# ---- indirect: the linear step of optimizeTraj_OLS (multiShoot_CRTBP_indirect.jl:181-182, :207) on the device.
# `phi` is what jacobianCalc_blocks returns (m x m x (n_nodes-1)); Jac_full is never formed.
function jacobianCalc_blocks(XC_all, t_TU, nstate, n_nodes, params)
    m = 2 * nstate; defect = pinned_array(Float64, m, n_nodes - 1); status = pinned_array(Int32, n_nodes - 1); phi = pinned_array(Float64, m, m, n_nodes - 1)
    GC.@preserve XC_all t_TU defect status phi check(ccall((:lto_indirect_defect_jac_traj, lib), Cint,
        (Ptr{Cvoid}, Ref{IndirectParams}, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        handle[], iparams(params), 1, n_nodes, m, XC_all, t_TU, C_NULL, C_NULL, defect, status, C_NULL, phi))
    phi
end
function newtonUpdate(phi::Array{Float64,3}, defect::Matrix{Float64}, n_nodes, flag_adjointsOnly::Bool)
    xc_update = zeros(12, n_nodes); status = zeros(Int32, 1)              # == reshape(xc_update2, 2*nstate, n_nodes) (:185)
    GC.@preserve phi defect xc_update status check(ccall((:lto_indirect_newton, lib), Cint,
        (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
        handle[], 1, n_nodes, flag_adjointsOnly ? 1 : 0, phi, defect, xc_update, status))
    xc_update
end

# ---- indirect: multiShoot_CRTBP_indirect (:58-345) for a BATCH of trajectories, iterated on the device.
# XC_batch: 12 x n_nodes x n_traj, t_batch: n_nodes x n_traj; thrustLimit / rho: per-trajectory vectors (continuation ladders).
function multiShoot_CRTBP_indirect_batch(XC_batch::Array{Float64,3}, t_batch::Matrix{Float64}, MU, DU, TU, mass0, thrustLimit::Vector{Float64},
                                         flag_adjointsOnly::Bool, maxIter, p, rho::Vector{Float64})
    n_nodes = size(XC_batch, 2); n_traj = size(XC_batch, 3)
    XC = copy(XC_batch); defect = zeros(12, n_nodes - 1, n_traj)
    status_flag = zeros(Int32, n_traj); iters = zeros(Int32, n_traj); er = zeros(n_traj)
    prm = iparams((MU, DU, TU, thrustLimit[1], mass0, 1.0, p, rho[1]))
    GC.@preserve XC t_batch thrustLimit rho defect status_flag iters er check(ccall((:lto_indirect_solve_batch, lib), Cint,
        (Ptr{Cvoid}, Ref{IndirectParams}, Int64, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        handle[], prm, n_traj, n_nodes, maxIter, flag_adjointsOnly ? 1 : 0, XC, t_batch, thrustLimit, rho, defect, status_flag, iters, er))
    (XC, defect, status_flag)
end
# ---- direct: multiShoot_CRTBP_direct (:465-594) for a BATCH of trajectories, iterated on the device (flagEnd = false, allowImpulsive = false).
# X_batch: nstate x n_nodes x n_traj, u_batch: 3 x n_nodes x n_traj, t_batch: n_nodes x n_traj, state_0 / state_f: 6 x n_traj (interpEndStates, :483)
function multiShoot_CRTBP_direct_batch(X_batch::Array{Float64,3}, u_batch::Array{Float64,3}, t_batch::Matrix{Float64}, state_0::Matrix{Float64},
                                       state_f::Matrix{Float64}, MU, DU, TU, nsteps, mass, Isp, maxIter)
    nstate = size(X_batch, 1); n_nodes = size(X_batch, 2); n_traj = size(X_batch, 3)
    p = DirectParams(); p.MU, p.DU, p.TU, p.Isp = MU, DU, TU, Isp
    X = copy(X_batch); u = copy(u_batch); defect = zeros(nstate, n_nodes - 1, n_traj); iters = zeros(Int32, n_traj); er = zeros(n_traj)
    GC.@preserve X u t_batch state_0 state_f defect iters er check(ccall((:lto_direct_solve_batch, lib), Cint,
        (Ptr{Cvoid}, Ref{DirectParams}, Int64, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble,
         Ptr{Float64}, Ptr{Int32}, Ptr{Float64}),
        handle[], p, n_traj, n_nodes, nstate, nsteps, maxIter, X, u, t_batch, state_0, state_f, mass, defect, iters, er))
    (X, u, defect, iters)
end
end # module

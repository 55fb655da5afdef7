# PigeonB200.jl — Julia host shim over libpigeon_b200.so (include/pigeon_b200.h).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no `julia` (probed; see DESIGN.md §2).  The Python mirror
# pigeon.jl_b200/mpc.py binds exactly the same symbols with ctypes and is what the tests drive; this file is the binding a
# Pigeon.jl maintainer would load instead (INTEGRATION.md).  Written for Julia >= 1.0 (the reference targets 1.0.x).
#
# It keeps the reference's API surface for the hot path:
#   X1()                                                   src/vehicles.jl:1-59
#   CoupledControlParams / DecoupledControlParams          src/coupled_lat_long.jl:23-40, src/decoupled_lat_long.jl:18-30
#   TrajectoryTube, straight_trajectory                    src/trajectories.jl:8-44, 96-105
#   HJICache, placeholder_HJICache                         src/HJI_computation.jl:26-37
#   Batched{Coupled,Decoupled}TrajectoryTrackingMPC        src/coupled_lat_long.jl:42-60, src/decoupled_lat_long.jl:32-50
#   compute_time_steps!, compute_linearization_nodes!, update_QP!, solve!, get_next_control, simulate
#                                                          src/model_predictive_control.jl:70-100
# Every array is vehicle-major: a Julia `Matrix{Float64}(k, B)` is the C `[B][k]` the ABI expects, so no transposes happen.
module PigeonB200

# Two ways to load this file:
#   (a) EMBEDDED — `include("…/julia/PigeonB200.jl")` inside `module Pigeon`, after src/model_predictive_control.jl (INTEGRATION.md).  The
#       shim then EXTENDS the reference's own generic functions (compute_time_steps!, compute_linearization_nodes!, update_QP!,
#       get_next_control, simulate from the parent module, solve! from Parametron) with methods for BatchedTrajectoryTrackingMPC, takes the
#       parent's TrajectoryTube / HJICache / control-parameter structs as they are, and exports only the batched names — nothing collides with
#       Pigeon's exports (X1, TrajectoryTube, HJICache, straight_trajectory, …), so `using .PigeonB200` inside Pigeon is safe.
#   (b) STANDALONE — `include` at top level (no Pigeon around): the shim defines its own generics and light-weight stand-ins of the reference's
#       types, and exports them too.
const EMBEDDED = parentmodule(@__MODULE__) !== @__MODULE__ && parentmodule(@__MODULE__) !== Main &&
                 isdefined(parentmodule(@__MODULE__), :TrajectoryTrackingMPC)

@static if EMBEDDED
    import ..compute_time_steps!, ..compute_linearization_nodes!, ..update_QP!, ..get_next_control, ..simulate
    import Parametron: solve!
else
    function compute_time_steps! end
    function compute_linearization_nodes! end
    function update_QP! end
    function solve! end
    function get_next_control end
    function simulate end
    export X1, CoupledControlParams, DecoupledControlParams, TrajectoryTube, straight_trajectory, HJICache, placeholder_HJICache,
           compute_time_steps!, compute_linearization_nodes!, update_QP!, solve!, get_next_control, simulate
end

export BatchedTrajectoryTrackingMPC, BatchedCoupledTrajectoryTrackingMPC, BatchedDecoupledTrajectoryTrackingMPC, step!,
       set_state!, set_HJI_cache!, reset_solved!, reset_solver!, solver_stats, set_guards!, from_autobox!, set_hji_policy!, hji_values,
       hji_optimal_control, set_path_search_window!, step_rollout_device!, set_pipeline_parts!, pipeline_parts, simulate_device!,
       set_history!, history, comm_init_all!, gather_all

const libpigeon = get(ENV, "PGN_LIB_PATH", joinpath(@__DIR__, "..", "pigeon.jl_b200", "libpigeon_b200.so"))

const PGN_COUPLED   = Int32(0)
const PGN_DECOUPLED = Int32(1)

struct PigeonError <: Exception
    code::Int32
    msg::String
end

@inline function check(rc::Integer)
    rc == 0 && return nothing
    throw(PigeonError(Int32(rc), unsafe_string(ccall((:pgn_last_error, libpigeon), Cstring, ()))))
end

# mirrors `pgn_config` field for field (include/pigeon_b200.h)
mutable struct PgnConfig
    kind::Int32; batch::Int32; N_short::Int32; N_long::Int32
    dt_short::Float64; dt_long::Float64
    use_correction_step::Int32; device::Int32
    rho::Float64; sigma::Float64; alpha::Float64; eps_abs::Float64; eps_rel::Float64; eps_prim_inf::Float64; eps_dual_inf::Float64
    max_iter::Int32; scaling::Int32; check_termination::Int32; adaptive_rho::Int32; adaptive_rho_interval::Int32
    adaptive_rho_tolerance::Float64
    warm_start::Int32; rk4_substeps::Int32
    hji_eps::Float64
    kkt_ordering::Int32; reserved::Int32
    PgnConfig() = new()
end

const VP_NAMES = (:L, :a, :b, :h, :G, :m, :Izz, :μ, :Cαf, :Cαr, :Cd0, :Cd1, :Cd2, :fwd_frac, :rwd_frac, :fwb_frac, :rwb_frac,
                  :Fx_max, :Fx_min, :Px_max, :δ_max, :κ_max, :inv_fiala_corrected)
const CP_NAMES = (:V_min, :V_max, :k_V, :k_s, :δ̇_max, :Q_Δs, :Q_Δψ, :Q_e, :W_β, :W_r, :W_HJI, :N_HJI, :R_δ, :R_Δδ, :R_Fx, :R_ΔFx)

const TRAJ_FIELDS = (:t, :s, :V, :A, :E, :N, :ψ, :κ, :θ, :ϕ, :edge_L, :edge_R)

function _control_params(kind::Int32; kw...)
    c = zeros(Float64, 16)
    check(ccall((:pgn_default_control_params, libpigeon), Cint, (Int32, Ptr{Float64}), kind, c))
    d = Dict{Symbol,Float64}(zip(CP_NAMES, c))
    for (k, v) in kw
        haskey(d, k) || throw(ArgumentError("unknown control parameter $k"))
        d[k] = Float64(v)
    end
    d
end
# control parameters arrive as the reference's CoupledControlParams / DecoupledControlParams struct (embedded) or as a Dict (standalone): same field names
_cp(cp::AbstractDict, k) = Float64(cp[k])
_cp(cp, k) = Float64(getproperty(cp, k))

@static if !EMBEDDED      # light-weight stand-ins of the reference's types; inside Pigeon the reference's own are used
"X1(): the reference's parameter Dict (src/vehicles.jl:1-59), values served by the library so both sides agree bit for bit."
function X1()
    v = zeros(Float64, 23)
    check(ccall((:pgn_x1_vehicle_params, libpigeon), Cint, (Ptr{Float64},), v))
    Dict{Symbol,Float64}(zip(VP_NAMES, v))
end

CoupledControlParams(; kw...)   = _control_params(PGN_COUPLED; kw...)
DecoupledControlParams(; kw...) = _control_params(PGN_DECOUPLED; kw...)

"TrajectoryTube (src/trajectories.jl:8-20): SoA of equal-length Float64 vectors."
struct TrajectoryTube
    t::Vector{Float64}; s::Vector{Float64}; V::Vector{Float64}; A::Vector{Float64}
    E::Vector{Float64}; N::Vector{Float64}; ψ::Vector{Float64}; κ::Vector{Float64}
    θ::Vector{Float64}; ϕ::Vector{Float64}; edge_L::Vector{Float64}; edge_R::Vector{Float64}
    function TrajectoryTube(t, s, V, A, E, N, ψ, κ, θ=zeros(length(t)), ϕ=zeros(length(t)),
                            edge_L=fill(4.0, length(t)), edge_R=fill(-4.0, length(t)))
        @assert length(t) == length(s) == length(V) == length(A) == length(E) == length(N) == length(ψ) == length(κ) ==
                length(θ) == length(ϕ) == length(edge_L) == length(edge_R)
        new(t, s, V, A, E, N, ψ, κ, θ, ϕ, edge_L, edge_R)
    end
end
Base.length(tr::TrajectoryTube) = length(tr.t)

"straight_trajectory(len, vel) (src/trajectories.jl:96-105)"
straight_trajectory(len, vel) = TrajectoryTube([0.0, len / vel], [0.0, len], [vel, vel], [0.0, 0.0], [0.0, 0.0], [0.0, len], [0.0, 0.0], [0.0, 0.0])

"HJICache (src/HJI_computation.jl:26-30): knots, V[n1..n7] and ∇V as a 7×n1×…×n7 Float32 array (component fastest)."
struct HJICache
    grid_knots::NTuple{7,Vector{Float32}}
    V::Array{Float32,7}
    ∇V::Array{Float32,8}
end
placeholder_HJICache() = HJICache(ntuple(_ -> Float32[-1000, 1000], 7), zeros(Float32, ntuple(_ -> 2, 7)), zeros(Float32, 7, ntuple(_ -> 2, 7)...))

end # !EMBEDDED

"Batched TrajectoryTrackingMPC (src/model_predictive_control.jl:32-68): B independent controllers living on one B200."
mutable struct BatchedTrajectoryTrackingMPC
    handle::Ptr{Cvoid}
    kind::Int32
    B::Int
    N::Int; nx::Int; nu::Int; n::Int; m::Int
    vehicle::Dict{Symbol,Float64}
    control_params                 # Dict, or the reference's Coupled/DecoupledControlParams struct
    trajectories::Vector           # TrajectoryTubes (the reference's, or the stand-in above): anything with the 12 vector fields
    HJI_cache                      # nothing | HJICache (the reference's interpolants, or the stand-in's raw arrays)
end

function BatchedTrajectoryTrackingMPC(kind::Int32, vehicle::Dict{Symbol,Float64}, trajectories::Vector, B::Integer;
                                      control_params=_control_params(kind), N_short=10, N_long=20, dt_short=0.01, dt_long=0.2,
                                      use_correction_step=true, device=-1, trajectory_index=nothing)
    cfg = PgnConfig()
    check(ccall((:pgn_default_config, libpigeon), Cint, (Ref{PgnConfig}, Int32), cfg, kind))
    cfg.batch = B; cfg.N_short = N_short; cfg.N_long = N_long; cfg.dt_short = dt_short; cfg.dt_long = dt_long
    cfg.use_correction_step = use_correction_step ? 1 : 0; cfg.device = device
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:pgn_create, libpigeon), Cint, (Ref{PgnConfig}, Ref{Ptr{Cvoid}}), cfg, h))
    d = zeros(Int32, 16)
    check(ccall((:pgn_qp_dims, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}), h[], d))
    mpc = BatchedTrajectoryTrackingMPC(h[], kind, B, d[1], d[2], d[3], d[4], d[5], vehicle, control_params, trajectories, nothing)
    finalizer(m -> (m.handle != C_NULL && ccall((:pgn_destroy, libpigeon), Cint, (Ptr{Cvoid},), m.handle); m.handle = C_NULL), mpc)
    vp = Float64[get(vehicle, k, 0.0) for k in VP_NAMES]
    cp = Float64[_cp(control_params, k) for k in CP_NAMES]
    check(ccall((:pgn_set_vehicle_params, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}), mpc.handle, vp))
    check(ccall((:pgn_set_control_params, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}), mpc.handle, cp))
    set_trajectories!(mpc, trajectories, trajectory_index === nothing ? Int32[(i - 1) % length(trajectories) for i in 1:B] : Int32.(trajectory_index))
    mpc
end
BatchedCoupledTrajectoryTrackingMPC(vehicle, trajectories, B; kw...)   = BatchedTrajectoryTrackingMPC(PGN_COUPLED, vehicle, trajectories, B; kw...)
BatchedDecoupledTrajectoryTrackingMPC(vehicle, trajectories, B; kw...) = BatchedTrajectoryTrackingMPC(PGN_DECOUPLED, vehicle, trajectories, B; kw...)

"mpc.trajectory = ... for the batch: `trajectory_index[i]` (0-based) selects the tube vehicle i tracks."
function set_trajectories!(mpc::BatchedTrajectoryTrackingMPC, trajectories::Vector, trajectory_index::Vector{Int32})
    n = length(trajectories[1])
    all(length(t) == n for t in trajectories) || throw(ArgumentError("all trajectories of a batch must have the same number of nodes"))
    # fields[k] is C [n_traj][n_nodes] == Julia Matrix(n_nodes, n_traj)
    fields = [hcat((getfield(t, f) for t in trajectories)...) for f in TRAJ_FIELDS]
    ptrs = Ptr{Float64}[pointer(f) for f in fields]
    GC.@preserve fields begin
        check(ccall((:pgn_set_trajectories, libpigeon), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Ptr{Float64}}), mpc.handle, length(trajectories), n, ptrs))
    end
    check(ccall((:pgn_assign_trajectories, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}), mpc.handle, trajectory_index))
    mpc.trajectories = trajectories
    mpc
end

"mpc.HJI_cache = cache (Pigeon.jl:40)"
function set_HJI_cache!(mpc::BatchedTrajectoryTrackingMPC, cache)
    dims = Int32[length(k) for k in cache.grid_knots]
    knots = vcat(cache.grid_knots...)
    # the reference's HJICache holds gridded interpolants (coefs: Array{Float32,7} and Array{SVector{7,Float32},7}); the stand-in holds the raw arrays
    V  = cache.V isa Array ? cache.V : cache.V.coefs
    gV = cache.∇V isa Array{Float32} ? cache.∇V : Array(reinterpret(Float32, cache.∇V.coefs))      # 7 components fastest (save(), HJI_computation.jl:59-64)
    GC.@preserve V gV check(ccall((:pgn_set_hji_cache, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                                  mpc.handle, dims, knots, V, gV))
    mpc.HJI_cache = cache
    mpc
end

"mpc.current_state (6×B), mpc.current_control (3×B), mpc.other_car_state (4×B), mpc.time_offset (B); `nothing` keeps the previous value."
function set_state!(mpc::BatchedTrajectoryTrackingMPC; current_state=nothing, current_control=nothing, other_car_state=nothing, time_offset=nothing)
    p(x) = x === nothing ? Ptr{Float64}(C_NULL) : pointer(x)
    GC.@preserve current_state current_control other_car_state time_offset begin
        check(ccall((:pgn_set_state, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    mpc.handle, p(current_state), p(current_control), p(other_car_state), p(time_offset)))
    end
    mpc
end

_mask(m) = m === nothing ? Ptr{UInt8}(C_NULL) : pointer(m)
"mpc.solved = false (per vehicle)"
reset_solved!(mpc, mask::Union{Nothing,Vector{UInt8}}=nothing) = GC.@preserve mask check(ccall((:pgn_reset_solved, libpigeon), Cint, (Ptr{Cvoid}, Ptr{UInt8}), mpc.handle, _mask(mask)))
"Parametron.initialize!(mpc.model) (ros_integration.jl:146), per vehicle"
# guards of the ROS callback (src/ros_integration.jl:84-87, 134-147): pause below a speed, previous control + re-initialisation on NaN
set_guards!(mpc; nan_fallback::Bool=false, pause_below_speed::Float64=0.0) =
    check(ccall((:pgn_set_guards, libpigeon), Cint, (Ptr{Cvoid}, Int32, Float64), mpc.handle, Int32(nan_fallback), pause_below_speed))
"from_autobox_callback (src/ros_integration.jl:48-151) for the whole batch: returns 5×B (δ, Fxf, Fxr, s_m, e_m)"
function from_autobox!(mpc, current_state::Matrix{Float64}, current_control::Matrix{Float64}, stamp::Vector{Float64};
                       other_car_state::Union{Nothing,Matrix{Float64}}=nothing)
    out = Matrix{Float64}(undef, 5, mpc.B)
    po = other_car_state === nothing ? Ptr{Float64}(C_NULL) : pointer(other_car_state)
    GC.@preserve other_car_state check(ccall((:pgn_from_autobox, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                             mpc.handle, current_state, current_control, po, stamp, out))
    out
end
# windowed closest-segment search of path_coordinates (src/trajectories.jl:71-80); 0 = full scan
set_path_search_window!(mpc, half_width::Integer) = check(ccall((:pgn_set_path_search_window, libpigeon), Cint, (Ptr{Cvoid}, Int32), mpc.handle, Int32(half_width)))
# use_HJI_policy[] of the callback (src/ros_integration.jl:47,115-118): V <= HJI_ϵ => BicycleControl(LP, optimal_control(...))
set_hji_policy!(mpc, on::Bool) = check(ccall((:pgn_set_hji_policy, libpigeon), Cint, (Ptr{Cvoid}, Int32), mpc.handle, Int32(on)))
"(V, ∇V) of the last step's HJIRelativeState(current_state, other_car_state) (src/ros_integration.jl:57-58); ∇V is 7×B"
function hji_values(mpc)
    V = Vector{Float64}(undef, mpc.B); g = Matrix{Float64}(undef, 7, mpc.B)
    check(ccall((:pgn_get_hji_values, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), mpc.handle, V, g))
    V, g
end
"optimal_control(dynamics, relative_state, ∇V) (src/HJI_computation.jl:133-158) for 7×M relative states / gradients -> 2×M (δ, Fx)"
function hji_optimal_control(mpc, relative_state::Matrix{Float64}, gradV::Matrix{Float64})
    M = size(relative_state, 2); out = Matrix{Float64}(undef, 2, M)
    check(ccall((:pgn_hji_optimal_control, libpigeon), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), mpc.handle, Int32(M), relative_state, gradV, out))
    out
end
reset_solver!(mpc, mask::Union{Nothing,Vector{UInt8}}=nothing) = GC.@preserve mask check(ccall((:pgn_reset_solver, libpigeon), Cint, (Ptr{Cvoid}, Ptr{UInt8}), mpc.handle, _mask(mask)))

# ---- the 5-call step API (src/model_predictive_control.jl:70-78) --------------------------------------------------------------
compute_time_steps!(mpc::BatchedTrajectoryTrackingMPC, t0::Vector{Float64}) = check(ccall((:pgn_compute_time_steps, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}), mpc.handle, t0))
compute_time_steps!(mpc::BatchedTrajectoryTrackingMPC, t0::Real) = compute_time_steps!(mpc, fill(Float64(t0), mpc.B))
compute_linearization_nodes!(mpc::BatchedTrajectoryTrackingMPC) = check(ccall((:pgn_compute_linearization_nodes, libpigeon), Cint, (Ptr{Cvoid},), mpc.handle))
update_QP!(mpc::BatchedTrajectoryTrackingMPC) = check(ccall((:pgn_update_qp, libpigeon), Cint, (Ptr{Cvoid},), mpc.handle))
solve!(mpc::BatchedTrajectoryTrackingMPC) = check(ccall((:pgn_solve, libpigeon), Cint, (Ptr{Cvoid},), mpc.handle))
"get_next_control(mpc): 3×B matrix of (δ, Fxf, Fxr) — BicycleControl per vehicle"
function get_next_control(mpc::BatchedTrajectoryTrackingMPC)
    out = Matrix{Float64}(undef, 3, mpc.B)
    check(ccall((:pgn_get_next_control, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}), mpc.handle, out))
    out
end
"the five calls fused into one library call"
function step!(mpc::BatchedTrajectoryTrackingMPC, t0::Vector{Float64}, out::Matrix{Float64}=Matrix{Float64}(undef, 3, mpc.B))
    check(ccall((:pgn_step, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), mpc.handle, t0, out))
    out
end

"one iteration of the simulate loop (src/model_predictive_control.jl:87-98) on device-resident data: d_t0 / d_out are device pointers (CuPtr); the plant step runs beside the QP solve"
step_rollout_device!(mpc::BatchedTrajectoryTrackingMPC, d_t0::Ptr{Float64}, d_out::Ptr{Float64}, dt::Float64=0.01) =
    check(ccall((:pgn_step_rollout_device, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64), mpc.handle, d_t0, d_out, dt))

"pipeline parts of the fused calls (step!, step_rollout_device!, simulate): vehicle ranges on their own streams; 0 = automatic, 1 = off; results do not depend on it"
set_pipeline_parts!(mpc::BatchedTrajectoryTrackingMPC, parts::Integer) = check(ccall((:pgn_set_pipeline_parts, libpigeon), Cint, (Ptr{Cvoid}, Int32), mpc.handle, Int32(parts)))
function pipeline_parts(mpc::BatchedTrajectoryTrackingMPC)
    n = Ref{Int32}(0)
    check(ccall((:pgn_get_pipeline_parts, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}), mpc.handle, n))
    Int(n[])
end
"the simulate loop with t0 resident on the device (CuPtr), enqueued on the handle's stream without a host synchronisation"
simulate_device!(mpc::BatchedTrajectoryTrackingMPC, d_t0::Ptr{Float64}, dt::Float64, n_steps::Integer; k0::Integer=0) =
    check(ccall((:pgn_simulate_device, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Int32, Int32), mpc.handle, d_t0, dt, Int32(k0), Int32(n_steps)))

"record every `stride`-th step of simulate on the device (capacity records); 0, 0 switches the recorder off"
set_history!(mpc::BatchedTrajectoryTrackingMPC, capacity::Integer, stride::Integer=1) =
    check(ccall((:pgn_set_history, libpigeon), Cint, (Ptr{Cvoid}, Int32, Int32), mpc.handle, Int32(capacity), Int32(capacity == 0 ? 0 : stride)))
"(qs, xs, us, ps) of the recorded steps: 6×B×n, nx×B×n, 3×B×n, 4×B×n"
function history(mpc::BatchedTrajectoryTrackingMPC)
    n = Ref{Int32}(0)
    check(ccall((:pgn_get_history, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), mpc.handle, n, C_NULL, C_NULL, C_NULL, C_NULL))
    qs = Array{Float64}(undef, 6, mpc.B, n[]); us = Array{Float64}(undef, 3, mpc.B, n[]); xs = Array{Float64}(undef, mpc.nx, mpc.B, n[]); ps = Array{Float64}(undef, 4, mpc.B, n[])
    n[] > 0 && check(ccall((:pgn_get_history, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), mpc.handle, n, qs, us, xs, ps))
    qs, xs, us, ps
end

"""
simulate(mpc, q0, u0, dt=0.01) (src/model_predictive_control.jl:80-100) for the whole batch: `for t in 0:dt:mpc.trajectory.t[end]`, entirely on
the device (one pgn_simulate call); returns the reference's (qs, xs, us, ps), recorded on the device every `stride`-th step.
q0 is 6×B, u0 3×B; t0 (B) shifts every vehicle's time axis.
"""
function simulate(mpc::BatchedTrajectoryTrackingMPC, q0::Matrix{Float64}, u0::Matrix{Float64}, dt=0.01; t0=zeros(mpc.B), stride::Integer=1,
                  n_steps::Integer=length(0:dt:mpc.trajectories[1].t[end]))
    set_state!(mpc; current_state=q0, current_control=u0)
    set_history!(mpc, cld(n_steps, stride), stride)
    check(ccall((:pgn_simulate, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64, Int32), mpc.handle, t0, Float64(dt), Int32(n_steps)))
    out = history(mpc)
    set_history!(mpc, 0, 0)
    out
end

# ---- multi-GPU: one handle per GPU driven by this one Julia thread, final gather over NCCL (SURVEY.md 8e) ------------------------------------
"form the NCCL communicator of the final gather over the handles' devices (one handle per GPU, equal batch sizes)"
function comm_init_all!(mpcs::Vector{BatchedTrajectoryTrackingMPC})
    hs = Ptr{Cvoid}[m.handle for m in mpcs]
    check(ccall((:pgn_comm_init_all, libpigeon), Cint, (Ptr{Ptr{Cvoid}}, Int32), hs, Int32(length(hs))))
end
"controls (3 × n·B), iteration counts and statuses (n·B) of the last step of every handle, rank-major, gathered with ncclAllGather over NVLink"
function gather_all(mpcs::Vector{BatchedTrajectoryTrackingMPC})
    n, B = length(mpcs), mpcs[1].B
    hs = Ptr{Cvoid}[m.handle for m in mpcs]
    c = Matrix{Float64}(undef, 3, n * B); it = Vector{Int32}(undef, n * B); st = Vector{Int32}(undef, n * B)
    check(ccall((:pgn_gather_all, libpigeon), Cint, (Ptr{Ptr{Cvoid}}, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}), hs, Int32(n), c, it, st))
    c, it, st
end

"per-vehicle OSQP-style statistics of the last solve! (the reference never inspects them, ros_integration.jl:127)"
function solver_stats(mpc::BatchedTrajectoryTrackingMPC)
    B = mpc.B
    iters = zeros(Int32, B); status = zeros(Int32, B); nrho = zeros(Int32, B)
    pri = zeros(B); dua = zeros(B); rho = zeros(B)
    check(ccall((:pgn_get_stats, libpigeon), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                mpc.handle, iters, status, pri, dua, rho, nrho))
    (iters=iters, status=status, pri_res=pri, dua_res=dua, rho=rho, rho_updates=nrho)
end

end # module

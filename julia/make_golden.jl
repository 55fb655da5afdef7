# make_golden.jl — the recipe that PINS parity to the reference itself.
#
# NOT RUN IN THIS REPO'S BUILD IMAGE (no julia, no network; DESIGN.md 2).  Run it once on a machine with Julia 1.0.x and the reference's
# pinned environment (env/Manifest.toml: OSQP.jl 0.4.0, Parametron 0.4.0, LinearDynamicsModels / DifferentialDynamicsModels @ master, ...):
#
#     cd /path/to/Pigeon.jl && julia --project=env -e 'using Pkg; Pkg.instantiate()'
#     julia --project=/path/to/Pigeon.jl/env julia/make_golden.jl /path/to/Pigeon.jl tests/golden/ref_golden.bin
#     python tests/golden/import_ref_golden.py tests/golden/ref_golden.bin        # -> tests/golden/ref_*.npz (commit these)
#
# after which tests/test_ref_golden.py (CPU: oracle vs golden) and tests/test_gpu_ref_golden.py (-m gpu: CUDA vs golden) stop skipping
# and the parity status in DESIGN.md can move from "unpinned" to "pinned by the reference's own outputs".
#
# What it runs — the reference's own code, file by file, WITHOUT the ROS layer (src/Pigeon.jl includes src/ros_integration.jl, which needs
# rospy and custom message packages; the hot path does not): src/math.jl, vehicles.jl, vehicle_dynamics.jl, HJI_computation.jl,
# trajectories.jl, model_predictive_control.jl, decoupled_lat_long.jl, coupled_lat_long.jl with the `using` lines of src/Pigeon.jl:1-23.
#
# What it records
#   dry/<ctl>/...      the dry run of src/Pigeon.jl:34-57 (straight_trajectory(30, 5), state (0,0,0,5,0,0), zero control) for
#                      ctl = X1DMPC (decoupled 10/20), X1CMPC (coupled 5/10) and C31 (coupled, default 10/20 horizon)
#   sim/<ctl>/...      200 steps of `simulate` (src/model_predictive_control.jl:80-100) on test/path/skidpadoval.world converted as
#                      TrajectoryTube(p::path) does (src/ros_integration.jl:13-16), started on the path at node 1 with Ux = 6, zero control,
#                      once with OSQP's own (wall-clock dependent) adaptive_rho_interval and once with the interval PINNED to 25 iterations
#                      (`sim25/`) — the oracle's setting; iteration counts can only be compared bit for bit in the pinned run
#   per step: t, current_state, current_control, ts, dt, qs, us, ps, A, B (ZOH) / B0, Bf (FOH), c, H, G, δ_min, δ_max, Fx_max, q_curr, u_curr,
#             M_HJI, b_HJI, the QP solution (q, u[, δ], σ, σ_HJI values), OSQP info.iter / status_val / rho_updates, the adaptive_rho_interval
#             OSQP actually used, get_next_control, the propagated state
#   hji/...            10^4 cache[x] lookups (V, ∇V) on placeholder_HJICache() and on an analytic 7x6x5x5x4x5x4 grid, in-grid / on faces /
#                      outside (src/HJI_computation.jl:66-72), and compute_reachability_constraint (M, b) on 2000 of them (:160-170)
#
# Container: "PGNGOLD1", then records  [u32 name length][name][u8 dtype 1=f64 2=i32 3=f32][u32 ndims][u64 dims...][column-major data].

module PigeonCore
using LinearAlgebra
using StaticArrays
using DifferentialDynamicsModels
using LinearDynamicsModels
using ForwardDiff
using Interpolations
using OSQP.MathOptInterfaceOSQP
import MathOptInterface
const MOI = MathOptInterface
using Parametron
using JLD2
import StaticArrays: SUnitRange
import DifferentialDynamicsModels: mod2piF, adiff
import Interpolations: GriddedInterpolation, Extrapolation
Parametron.Parameter(A::AbstractArray, model) = Parameter(identity, A, model)
const REF = joinpath(ARGS[1], "src")
for f in ("math.jl", "vehicles.jl", "vehicle_dynamics.jl", "HJI_computation.jl", "trajectories.jl", "model_predictive_control.jl",
          "decoupled_lat_long.jl", "coupled_lat_long.jl")
    include(joinpath(REF, f))
end
end # module

using .PigeonCore
using StaticArrays, Random, LinearAlgebra
import Parametron, OSQP
const P = PigeonCore
const MOI = P.MOI

# ---- container ----------------------------------------------------------------------------------------------------------------------
const OUT = open(ARGS[2], "w")
write(OUT, b"PGNGOLD1")
dtype_code(::Type{Float64}) = UInt8(1); dtype_code(::Type{Int32}) = UInt8(2); dtype_code(::Type{Float32}) = UInt8(3)
function put(name::String, a::AbstractArray{T}) where {T<:Union{Float64,Int32,Float32}}
    a = Array(a)
    write(OUT, UInt32(sizeof(name))); write(OUT, name); write(OUT, dtype_code(T)); write(OUT, UInt32(ndims(a)))
    for d in size(a); write(OUT, UInt64(d)); end
    write(OUT, a)
end
put(name::String, x::Real) = put(name, [Float64(x)])
put(name::String, a::AbstractArray{<:Integer}) = put(name, Int32.(a))
put(name::String, a::AbstractArray{<:Real}) = put(name, Float64.(a))
flat(v::AbstractVector{<:StaticVector}) = Float64[x[i] for i in 1:length(v[1]), x in v]      # k x N

pv(p) = try copy(p()) catch; copy(p.val[]) end                                                 # value of a Parametron.Parameter
stackp(ps) = isempty(ps) ? zeros(0) : cat((pv(p) for p in ps)...; dims = ndims(pv(ps[1])) + 1)

# OSQP internals behind Parametron: model.optimizer is the MathOptInterfaceOSQP optimizer; .results holds the last osqp_solve
function osqp_info(mpc)
    opt = mpc.model.optimizer
    res = opt.results
    iters = Int32(res.info.iter); status = Int32(res.info.status_val); rho_upd = Int32(res.info.rho_updates)
    interval = Int32(-1)
    try      # the interval osqp_setup / osqp_solve settled on (workspace->settings->adaptive_rho_interval)
        ws = unsafe_load(opt.inner.workspace)
        interval = Int32(unsafe_load(ws.settings).adaptive_rho_interval)
    catch
    end
    iters, status, rho_upd, interval
end

function put_step(prefix, mpc, coupled::Bool)
    ts = mpc.time_steps
    put("$prefix/ts", collect(ts.ts)); put("$prefix/dt", collect(ts.dt)); put("$prefix/prev_ts", collect(ts.prev_ts))
    put("$prefix/qs", flat(mpc.qs)); put("$prefix/us", flat(mpc.us)); put("$prefix/ps", flat(mpc.ps))
    Q = mpc.parameters
    for f in (:A, :B0, :Bf, :c, :H, :G, :δ_min, :δ_max)
        put("$prefix/$(f)", stackp(getfield(Q, f)))
    end
    put("$prefix/B", stackp(Q.B))
    put("$prefix/q_curr", pv(Q.q_curr))
    if coupled
        put("$prefix/Fx_max", stackp(Q.Fx_max)); put("$prefix/u_curr", pv(Q.u_curr))
        put("$prefix/M_HJI", pv(Q.M_HJI)); put("$prefix/b_HJI", pv(Q.b_HJI))
    else
        put("$prefix/delta_curr", pv(Q.δ_curr))
    end
    V = mpc.variables
    val(x) = Parametron.value.(Ref(mpc.model), x)
    put("$prefix/x_q", val(V.q))
    if coupled
        put("$prefix/x_u", val(V.u)); put("$prefix/x_sigma_HJI", val(V.σ_HJI))
    else
        put("$prefix/x_delta", val(V.δ))
    end
    put("$prefix/x_sigma", val(V.σ))
    it, st, ru, iv = osqp_info(mpc)
    put("$prefix/osqp", Int32[it, st, ru, iv])
    put("$prefix/next_control", collect(P.get_next_control(mpc)))
end

function four_calls!(mpc, t)
    P.compute_time_steps!(mpc, t); P.compute_linearization_nodes!(mpc); P.update_QP!(mpc); P.solve!(mpc)
end

# The default other_car_state is zeros(SimpleCarState): with the placeholder cache (V = 0 <= HJI_ϵ, ∇V = 0) the constraint counts as active and
# optimal_disturbance divides by the other car's speed 0 (src/HJI_computation.jl:103): b_HJI becomes NaN.  The golden runs therefore park the
# other car outside the placeholder grid (|x| > 1000): cache[x] = (Inf, 0), constraint inactive (M = 0, b = 1) — what the tests of this repo use.
const FAR_CAR = P.SimpleCarState(1e4, 1e4, 0., 5.)
pin_interval!(mpc, n) = MOI.set!(mpc.model.optimizer, P.OSQPSettings.AdaptiveRhoInterval(), n)

# ---- (a) the dry run of src/Pigeon.jl:34-57 -----------------------------------------------------------------------------------------------
function dry_run(name, ctor, coupled; kw...)
    for (tag, pin) in (("dry", false), ("dry25", true))
        mpc = ctor(P.X1(), P.straight_trajectory(30., 5.); kw...)
        pin && pin_interval!(mpc, 25)
        mpc.current_state = P.BicycleState(0., 0., 0., 5., 0., 0.)
        mpc.current_control = P.BicycleControl(0., 0., 0.)
        mpc.other_car_state = FAR_CAR
        Parametron.initialize!(mpc.model)
        four_calls!(mpc, 0.)
        put_step("$tag/$name", mpc, coupled)
    end
end
dry_run("X1DMPC", P.DecoupledTrajectoryTrackingMPC, false)
dry_run("X1CMPC", P.CoupledTrajectoryTrackingMPC, true; N_short = 5, N_long = 10)
dry_run("C31", P.CoupledTrajectoryTrackingMPC, true)

# ---- (b) simulate on test/path/skidpadoval.world -----------------------------------------------------------------------------------------
function read_world(fname)
    d = Dict{String,Vector{Float64}}()
    for line in eachline(fname)
        isempty(strip(line)) && continue
        k, v = split(line, ":"; limit = 2)
        d[strip(k)] = [parse(Float64, x) for x in split(strip(strip(v), ['[', ']']), ",") if !isempty(strip(x))]
    end
    d
end
function world_tube(fname)       # TrajectoryTube(p::path), src/ros_integration.jl:13-16
    w = read_world(fname)
    P.TrajectoryTube{Float64}(P.invcumtrapz(w["UxDes_mps"], w["s_m"]), w["s_m"], w["UxDes_mps"], w["AxDes_mps2"], w["posE_m"], w["posN_m"],
                              w["psi_rad"], w["k_1pm"], w["grade_rad"], 0 * w["grade_rad"], w["edgeL_m"], w["edgeR_m"])
end
function sim_run(tag, name, ctor, coupled, pin; nsteps = 200, dt = 0.01, kw...)
    traj = world_tube(joinpath(ARGS[1], "test", "path", "skidpadoval.world"))
    mpc = ctor(P.X1(), traj; kw...)
    pin && pin_interval!(mpc, 25)
    Parametron.initialize!(mpc.model)
    mpc.current_state = P.BicycleState(traj.E[1], traj.N[1], traj.ψ[1], 6., 0., 0.)
    mpc.current_control = P.BicycleControl(0., 0., 0.)
    mpc.other_car_state = FAR_CAR
    for k in 0:nsteps-1          # the body of simulate (model_predictive_control.jl:87-98), with the per-step dump added
        t = k * dt
        pre = "$tag/$name/step$(lpad(k, 3, '0'))"
        put("$pre/t", t); put("$pre/state", collect(mpc.current_state)); put("$pre/control", collect(mpc.current_control))
        four_calls!(mpc, t)
        put_step(pre, mpc, coupled)
        mpc.current_state = P.propagate(mpc.dynamics, mpc.current_state, P.StepControl(dt, P.BicycleControl2(mpc.current_control)))
        mpc.current_control = P.get_next_control(mpc)
        put("$pre/state_next", collect(mpc.current_state))
    end
end
for (tag, pin) in (("sim", false), ("sim25", true))
    sim_run(tag, "C31", P.CoupledTrajectoryTrackingMPC, true, pin)
    sim_run(tag, "X1CMPC", P.CoupledTrajectoryTrackingMPC, true, pin; N_short = 5, N_long = 10)
    sim_run(tag, "X1DMPC", P.DecoupledTrajectoryTrackingMPC, false, pin)
end

# ---- (c) HJI lookups and the reachability constraint -----------------------------------------------------------------------------------------
function analytic_cache()
    dims = (7, 6, 5, 5, 4, 5, 4)
    rng = ((-15., 15.), (-15., 15.), (-π, π), (1., 15.), (-2., 2.), (1., 15.), (-1., 1.))
    knots = tuple((Float32.(collect(range(r[1], stop = r[2], length = n))) for (r, n) in zip(rng, dims))...)
    V = zeros(Float32, dims); G = zeros(SVector{7,Float32}, dims)
    for I in CartesianIndices(dims)
        x = [Float64(knots[d][I[d]]) for d in 1:7]
        dE, dN, dψ, Ux, Uy, Vo, r = x
        R = sqrt((dE / 4)^2 + (dN / 2)^2 + 0.01)
        V[I] = Float32(R - 1 + 0.05 * (Ux - Vo) * cos(dψ) + 0.02 * Uy * r)
        G[I] = SVector{7,Float32}(dE / 16 / R, dN / 4 / R, -0.05 * (Ux - Vo) * sin(dψ), 0.05 * cos(dψ), 0.02 * r, -0.05 * cos(dψ), 0.02 * Uy)
    end
    put("hji/analytic/knots", vcat(knots...)); put("hji/analytic/dims", collect(Int32.(dims)))
    put("hji/analytic/V", V); put("hji/analytic/gradV", Array(reshape(reinterpret(Float32, G), (7, dims...))))
    P.HJICache(knots, P.interpolate(Float32, Float32, knots, V, P.Gridded(P.Linear())),
               P.interpolate(Float32, SVector{7,Float32}, knots, G, P.Gridded(P.Linear())))
end
Random.seed!(0x5049474E)
for (name, cache, lo, hi) in (("placeholder", P.placeholder_HJICache(), fill(-1200., 7), fill(1200., 7)),
                              ("analytic", analytic_cache(), [-17., -17, -3.5, 0.5, -2.2, 0.5, -1.1], [17., 17, 3.5, 15.5, 2.2, 15.5, 1.1]))
    M = 10_000
    X = lo .+ (hi .- lo) .* rand(7, M)
    for j in 1:200                                     # exactly on knots / faces
        d = rand(1:7); X[d, j] = Float64(cache.grid_knots[d][rand(1:length(cache.grid_knots[d]))])
    end
    Vs = zeros(M); Gs = zeros(7, M)
    for j in 1:M
        v, g = cache[P.HJIRelativeState(X[:, j]...)]
        Vs[j] = v; Gs[:, j] = collect(g)
    end
    put("hji/$name/x", X); put("hji/$name/V", Vs); put("hji/$name/grad", Gs)
    # compute_reachability_constraint(dynamics, cache, relative_state, ϵ, uR) (src/HJI_computation.jl:160-170) as update_QP! calls it
    dyn = P.VehicleModel(P.X1())
    K = 2000; Mb = zeros(3, K); U = zeros(2, K)
    for j in 1:K
        uR = P.BicycleControl2(0.3 * (2rand() - 1), 4000 * (2rand() - 1)); U[:, j] = collect(uR)
        M_, b_ = P.compute_reachability_constraint(dyn, cache, P.HJIRelativeState(X[:, j]...), 0.05, uR)
        Mb[1:2, j] = collect(M_); Mb[3, j] = b_
    end
    put("hji/$name/uR", U); put("hji/$name/Mb", Mb)
end

close(OUT)
println("wrote ", ARGS[2], " (", filesize(ARGS[2]), " bytes)")

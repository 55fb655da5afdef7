// pgn_nodes.cu — time steps, linearisation-node generation (one warp per vehicle), control extraction and plant rollout (one vehicle
// per thread) over SoA HBM arrays.
//   compute_time_steps!            reference src/model_predictive_control.jl:17-30
//   compute_linearization_nodes!   src/coupled_lat_long.jl:62-142, src/decoupled_lat_long.jl:52-104
//   trajectory lookups             src/trajectories.jl:47-94, src/math.jl:4-9
//   get_next_control               src/coupled_lat_long.jl:370-374, src/decoupled_lat_long.jl:275-278
//   simulate's plant step          src/model_predictive_control.jl:94-95
#include <algorithm>

#include "pgn_internal.h"

namespace pgn {

// ---- trajectory lookups -----------------------------------------------------------------------------------------------
struct TrajNode { double t, s, V, A, E, N, psi, kappa, theta, phi; };

// number of elements < x (Julia searchsortedfirst - 1) / <= x (searchsortedlast) in a sorted array
__device__ __forceinline__ int count_lt(const double* __restrict__ v, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(v + mid) < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int count_le(const double* __restrict__ v, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(v + mid) <= x) lo = mid + 1; else hi = mid; }
    return lo;
}
// interp_by_s: gridded linear in s with Line() extrapolation (trajectories.jl:32-35)
__device__ __forceinline__ void interp_by_s(const TrajView& tv, int base, double sq, TrajNode& o) {
    const int nn = tv.n_nodes;
    const double* s = tv.f[1] + base;
    int i = count_le(s, nn, sq);
    i = min(max(i, 1), nn - 1);
    const int k = i - 1;
    const double s0 = __ldg(s + k), s1 = __ldg(s + k + 1);
    const double w = (sq - s0) / (s1 - s0);
#define PGN_L(F) ((1 - w) * __ldg(tv.f[F] + base + k) + w * __ldg(tv.f[F] + base + k + 1))
    o.E = PGN_L(4); o.N = PGN_L(5); o.psi = PGN_L(6); o.kappa = PGN_L(7); o.theta = PGN_L(8); o.phi = PGN_L(9);
#undef PGN_L
}
// traj(t) (trajectories.jl:47-54)
__device__ __forceinline__ TrajNode traj_at_time(const TrajView& tv, int base, double tq) {
    const int nn = tv.n_nodes;
    const double *t = tv.f[0] + base, *s = tv.f[1] + base, *V = tv.f[2] + base;
    int i = count_lt(t, nn, tq);
    i = min(max(i, 1), nn - 1);
    const int k = i - 1;
    const double Vk = __ldg(V + k), tk = __ldg(t + k);
    const double A = (__ldg(V + k + 1) - Vk) / (__ldg(t + k + 1) - tk);
    const double dt = tq - tk;
    TrajNode o;
    o.t = tq; o.s = __ldg(s + k) + Vk * dt + A * dt * dt / 2; o.V = Vk + A * dt; o.A = A;
    interp_by_s(tv, base, o.s, o);
    return o;
}
// traj[s] (trajectories.jl:55-68)
__device__ __forceinline__ TrajNode traj_at_s(const TrajView& tv, int base, double sq) {
    const int nn = tv.n_nodes;
    const double *t = tv.f[0] + base, *s = tv.f[1] + base, *V = tv.f[2] + base;
    int i = count_lt(s, nn, sq);
    i = min(max(i, 1), nn - 1);
    const int k = i - 1;
    const double Vk = __ldg(V + k), tk = __ldg(t + k);
    const double A = (__ldg(V + k + 1) - Vk) / (__ldg(t + k + 1) - tk);
    const double ds = sq - __ldg(s + k);
    double dt;
    if (fabs(A) < 1e-3 || sq > __ldg(s + nn - 1)) dt = ds / Vk;
    else dt = (sqrt(2 * A * ds + Vk * Vk) - Vk) / A;
    TrajNode o;
    o.t = tk + dt; o.s = sq; o.V = Vk + A * dt; o.A = A;
    interp_by_s(tv, base, sq, o);
    return o;
}
// path_coordinates (trajectories.jl:71-93): closest segment by an O(n_nodes) scan (first minimum wins), then (s, e).
// Deviation: the sqrt argument is floored at 0 (the reference raises a DomainError on negative round-off).
// The scan is spread over the 32 lanes of a warp (segment i on lane i mod 32), followed by a butterfly arg-min whose tie-break keeps the
// smallest segment index, i.e. exactly the serial first-minimum-wins result.  Every lane returns (s, e).
__device__ __forceinline__ void path_coordinates_warp(const TrajView& tv, int base, double x, double y, int lane, double& s_out, double& e_out, double* t_out = nullptr,
                                                      int seg_lo = 0, int seg_hi = 0x7fffffff, int* seg_out = nullptr) {
    const int nn = tv.n_nodes;
    const double *E = tv.f[4] + base, *Nn = tv.f[5] + base;
    double d2min = INFINITY;
    int imin = 0x7fffffff;
    // [seg_lo, seg_hi): all segments by default; a window around the previous step's segment when the caller asked for one
    const int i_end = min(nn - 1, seg_hi);
    for (int i = max(seg_lo, 0) + lane; i < i_end; i += 32) {
        const double ax = __ldg(E + i), ay = __ldg(Nn + i);
        const double bx = __ldg(E + i + 1), by = __ldg(Nn + i + 1);
        const double vx = bx - ax, vy = by - ay;
        double lam = (vx * (x - ax) + vy * (y - ay)) / (vx * vx + vy * vy);
        lam = lam > 1 ? 1 : (lam < 0 ? 0 : lam);
        const double px = (1 - lam) * ax + lam * bx, py = (1 - lam) * ay + lam * by;
        const double d2 = (px - x) * (px - x) + (py - y) * (py - y);
        if (d2 < d2min) { d2min = d2; imin = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double d2o = __shfl_xor_sync(0xffffffffu, d2min, o);
        const int io = __shfl_xor_sync(0xffffffffu, imin, o);
        if (d2o < d2min || (d2o == d2min && io < imin)) { d2min = d2o; imin = io; }
    }
    const int i = imin == 0x7fffffff ? max(seg_lo, 0) : imin;      // all distances NaN: the serial scan keeps its first segment
    if (seg_out) *seg_out = i;
    const double ex = __ldg(E + i), ey = __ldg(Nn + i);
    const double vx = __ldg(E + i + 1) - ex, vy = __ldg(Nn + i + 1) - ey;
    const double wx = x - ex, wy = y - ey;
    const double arg = wx * wx + wy * wy - d2min;
    const double ds = sqrt(arg > 0 ? arg : 0.0);
    s_out = __ldg(tv.f[1] + base + i) + ds;
    e_out = sqrt(d2min) * sign_(vx * wy - vy * wx);
    if (t_out) {      // third return value of path_coordinates (trajectories.jl:85-92): time at which the trajectory passes the closest point
        const double Vi = __ldg(tv.f[2] + base + i), ti = __ldg(tv.f[0] + base + i);
        const double A = (__ldg(tv.f[2] + base + i + 1) - Vi) / (__ldg(tv.f[0] + base + i + 1) - ti);
        const double dt = fabs(A) < 1e-3 ? ds / Vi : (sqrt(2 * A * ds + Vi * Vi) - Vi) / A;
        *t_out = ti + dt;
    }
}

// ---- compute_time_steps! ---------------------------------------------------------------------------------------------
// The guards of the callback run BEFORE compute_time_steps! (src/ros_integration.jl:77-87 return early): a vehicle that is paused (Ux below
// the threshold) or whose time lies outside the trajectory keeps ts, dt and prev_ts — prev_ts stays the knot vector of the last SOLVED QP,
// which update_interpolations! needs when the vehicle resumes — and is flagged in `skip` for the stages that follow.
__global__ void k_time_steps(int B, int Ns, int Nl, double dt_short, double dt_long, int corr, const double* __restrict__ t0v,
                             double* __restrict__ ts, double* __restrict__ dtv, double* __restrict__ prev_ts,
                             uint8_t* __restrict__ skip, const double* __restrict__ Ux, double pause_below_speed, const uint8_t* __restrict__ tskip,
                             const uint8_t* __restrict__ hold) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    if (hold && hold[v]) return;          // deferred solve in progress / vehicle already at the end of the loop: nothing of the step is touched
    if (skip) {
        const bool sk = (pause_below_speed > 0.0 && Ux[v] < pause_below_speed) || (tskip && tskip[v]);
        skip[v] = sk;
        if (sk) return;
    }
    const int N = 1 + Ns + Nl;
    double* tsv = ts + (size_t)v * N;
    double* pv = prev_ts + (size_t)v * N;
    double* dv = dtv + (size_t)v * (N - 1);
    for (int i = 0; i < N; i++) pv[i] = tsv[i];
    const double t0 = t0v[v];
    double t0_long = t0 + Ns * dt_short;
    if (corr) t0_long = dt_long * ceil((t0_long + dt_short) / dt_long - 1);
    for (int i = 0; i <= Ns; i++) tsv[i] = t0 + dt_short * i;
    for (int i = 1; i <= Nl; i++) tsv[Ns + i] = t0_long + dt_long * i;
    for (int i = 0; i < N - 1; i++) dv[i] = tsv[i + 1] - tsv[i];
}

// ---- compute_linearization_nodes! ------------------------------------------------------------------------------------
struct NodeArgs {
    int B, N, Ns, kind, n_sol;
    int v0, nv;                                   // vehicle range of this launch (a pipeline part, pgn_set_pipeline_parts); B stays the SoA stride
    VehParams P; CtrlParams C; double un0, un1;
    TrajView tv;
    const double *state, *control, *toff; const uint8_t* solved; const int32_t* traj_id;
    const double *ts, *dt, *prev_ts, *sol_x;
    double *qs, *us, *ps;
    uint8_t* skip; double pause_below_speed;      // guard of src/ros_integration.jl:84-87
    double* se0;                                  // decoupled: (s0, e0) of the closest-segment scan, [2][B]
    int window; int32_t* last_seg;                // optional windowed closest-segment search (pgn_set_path_search_window)
    const uint8_t* tskip;                         // callback entry point only: time outside the trajectory interval (src/ros_integration.jl:77-80)
    const uint8_t* hold;                          // simulate loops with deferred solves: vehicles that do not start a new step in this round
};

// decoupled: always the steady-state rollout (decoupled_lat_long.jl:65-103).  A recurrence over the horizon nodes, one vehicle per
// THREAD: run on lane 0 of the vehicle's warp it left 31 lanes idle (1.6 ms per 8192 vehicles); the closest-segment scan stays
// warp-per-vehicle in k_nodes and hands (s0, e0) over through a small buffer.
__device__ void decoupled_cold_rollout(const NodeArgs& a, int v, double s0, double e0) {
    const int B = a.B, N = a.N, Ns = a.Ns;
    const VehParams& P = a.P;
    const CtrlParams& C = a.C;
    const double E0 = a.state[0 * B + v], N0 = a.state[1 * B + v], psi0 = a.state[2 * B + v];
    const double Ux0 = a.state[3 * B + v], Uy0 = a.state[4 * B + v], r0 = a.state[5 * B + v];
    const double d0 = a.control[0 * B + v], Fxf0 = a.control[1 * B + v], Fxr0 = a.control[2 * B + v];
    const double Fx0 = Fxf0 + Fxr0;
    const bool path_mode = isnan(a.toff[v]);
    const int base = a.traj_id[v] * a.tv.n_nodes;
    const double* ts = a.ts + (size_t)v * N;
    const double* dt = a.dt + (size_t)v * (N - 1);
    double* qs = a.qs + (size_t)v * N * 4;
    double* us = a.us + (size_t)v * N * 2;
    double* ps = a.ps + (size_t)v * N * 4;
        // decoupled: always the steady-state rollout (decoupled_lat_long.jl:65-103)
        double s = s0;
        double V = hypot(Ux0, Uy0);
        const double beta0 = atan2(Uy0, Ux0);
        double Fyf0, Fyr0;
        {
            double sd, cd;
            sincos(d0, &sd, &cd);
            lateral_tire_forces<double>(P, atan2(Uy0 + P.a * r0, Ux0) - d0, atan2(Uy0 - P.b * r0, Ux0), Fxf0, Fxr0, sd, cd, Fyf0, Fyr0);
        }
        double sb0, cb0;
        sincos(beta0, &sb0, &cb0);
        for (int i = 0; i < N; i++) {
            const double tau = (i == N - 1) ? dt[i - 1] : dt[i];
            TrajNode tj = traj_at_s(a.tv, base, s);
            const double kappa = tj.kappa;
            double A_des = tj.A + C.k_V * (tj.V - V) / tau + (path_mode ? 0.0 : C.k_s * (traj_at_time(a.tv, base, ts[i]).s - s) / tau / tau);
            A_des = fmin(fmax(A_des, (C.V_min - V) / tau), (C.V_max - V) / tau);
            double A;
            if (i == 0) {
                qs[0] = Uy0; qs[1] = r0; qs[2] = adiff(psi0, tj.psi); qs[3] = e0;
                us[0] = d0; us[1] = Fx0;
                ps[0] = Ux0; ps[1] = kappa; ps[2] = 0; ps[3] = 0;
                double q6[6] = {E0, N0, psi0, Ux0, Uy0, r0}, qd[6];
                vehicle_model<MODEL_BICYCLE, double>(P, q6, d0, Fx0, 0.0, 0.0, qd);
                A = (qd[3] - r0 * Uy0) * cb0 + (qd[4] + r0 * Ux0) * sb0;
            } else if (i <= Ns) {
                SteadyState est = steady_state_estimates(P, V, A_des, kappa, 1, r0, beta0, d0, Fyf0);
                qs[4 * i + 0] = Uy0; qs[4 * i + 1] = r0; qs[4 * i + 2] = adiff(psi0, tj.psi); qs[4 * i + 3] = e0;
                us[2 * i + 0] = est.delta; us[2 * i + 1] = est.Fxf + est.Fxr;
                ps[4 * i + 0] = est.Ux; ps[4 * i + 1] = kappa; ps[4 * i + 2] = 0; ps[4 * i + 3] = 0;
                A = est.A;
            } else {
                SteadyState est = steady_state_estimates(P, V, A_des, kappa, 4, V * kappa, 0.0, 0.0, 0.0);
                qs[4 * i + 0] = est.Uy; qs[4 * i + 1] = est.r; qs[4 * i + 2] = -est.beta; qs[4 * i + 3] = 0;
                us[2 * i + 0] = est.delta; us[2 * i + 1] = est.Fxf + est.Fxr;
                ps[4 * i + 0] = est.Ux; ps[4 * i + 1] = kappa; ps[4 * i + 2] = 0; ps[4 * i + 3] = 0;
                A = est.A;
            }
            if (i == N - 1) break;
            V = V + A * tau;
            s = s + V * tau + A * tau * tau / 2;
        }
}
__global__ void __launch_bounds__(128) k_nodes_decoupled_rollout(const NodeArgs a, const double* __restrict__ se0) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= a.nv) return;
    const int v = a.v0 + iv;
    if (a.hold && a.hold[v]) return;
    if ((a.pause_below_speed > 0.0 || a.tskip) && a.skip[v]) return;
    decoupled_cold_rollout(a, v, se0[v], se0[a.B + v]);
}

// One warp per vehicle: the lanes share the closest-segment scan and, on warm steps, take one horizon node each; the cold rollout
// (a recurrence over the nodes) runs on lane 0.
__global__ void __launch_bounds__(128) k_nodes(const NodeArgs a) {
    const int iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (iv >= a.nv) return;
    const int v = a.v0 + iv;
    if (a.hold && a.hold[v]) return;
    const int B = a.B, N = a.N, Ns = a.Ns;
    const VehParams& P = a.P;
    const CtrlParams& C = a.C;
    const double E0 = a.state[0 * B + v], N0 = a.state[1 * B + v], psi0 = a.state[2 * B + v];
    const double Ux0 = a.state[3 * B + v], Uy0 = a.state[4 * B + v], r0 = a.state[5 * B + v];
    const double d0 = a.control[0 * B + v], Fxf0 = a.control[1 * B + v], Fxr0 = a.control[2 * B + v];
    const double Fx0 = Fxf0 + Fxr0;
    const bool path_mode = isnan(a.toff[v]);
    if ((a.pause_below_speed > 0.0 || a.tskip) && a.skip[v]) return;      // flagged by k_time_steps: the callback returned early, nothing is touched
    const int base = a.traj_id[v] * a.tv.n_nodes;
    const double* ts = a.ts + (size_t)v * N;
    const double* dt = a.dt + (size_t)v * (N - 1);
    const int nx = a.kind == PGN_COUPLED ? 6 : 4;
    double* qs = a.qs + (size_t)v * N * nx;
    double* us = a.us + (size_t)v * N * 2;
    double* ps = a.ps + (size_t)v * N * 4;

    double s0, e0;
    {
        // with a window w > 0 only the segments within w of the previous step's closest segment are scanned (same result
        // as the full scan of trajectories.jl:71-80 as long as the true closest segment lies in the window); a vehicle without a valid
        // previous segment (first step, after pgn_set_state / pgn_assign_trajectories / pgn_reset_solved) scans everything
        const int last = a.window > 0 ? a.last_seg[v] : -1;
        int seg;
        if (last >= 0) path_coordinates_warp(a.tv, base, E0, N0, lane, s0, e0, nullptr, last - a.window, last + a.window + 1, &seg);
        else path_coordinates_warp(a.tv, base, E0, N0, lane, s0, e0, nullptr, 0, 0x7fffffff, &seg);
        if (lane == 0) a.last_seg[v] = seg;
    }

    if (a.kind == PGN_COUPLED) {
        TrajNode tj = traj_at_s(a.tv, base, s0);
        double ds = s0 - traj_at_time(a.tv, base, ts[0]).s;
        const double dpsi = adiff(psi0, tj.psi);
        if (lane == 0) {
            qs[0] = ds; qs[1] = Ux0; qs[2] = Uy0; qs[3] = r0; qs[4] = dpsi; qs[5] = e0;
            us[0] = d0; us[1] = Fx0;
            ps[0] = tj.V; ps[1] = tj.kappa; ps[2] = 0; ps[3] = 0;
        }
        if (a.solved[v]) {
            // warm: previous QP solution interpolated in prev_ts (update_interpolations!, coupled_lat_long.jl:86-102,189-195)
            const double* pts = a.prev_ts + (size_t)v * N;
            const double* X = a.sol_x + (size_t)v * a.n_sol;
            const double tend = pts[N - 1];
            for (int i = 1 + lane; i < N; i += 32) {
                const double t = ts[i];
                const double tq = t < tend ? t : tend;
                int k = 0;
                { int lo = 0, hi = N; while (lo < hi) { int mid = (lo + hi) >> 1; if (pts[mid] <= tq) lo = mid + 1; else hi = mid; } k = lo; }
                k = min(max(k, 1), N - 1) - 1;
                const double w = (tq - pts[k]) / (pts[k + 1] - pts[k]);
                double q0 = 0;
                for (int c = 0; c < 6; c++) {
                    const double qv = (1 - w) * X[6 * k + c] + w * X[6 * (k + 1) + c];
                    qs[6 * i + c] = qv;
                    if (c == 0) q0 = qv;
                }
                us[2 * i + 0] = ((1 - w) * X[6 * N + 2 * k + 0] + w * X[6 * N + 2 * (k + 1) + 0]) * a.un0;
                us[2 * i + 1] = ((1 - w) * X[6 * N + 2 * k + 1] + w * X[6 * N + 2 * (k + 1) + 1]) * a.un1;
                const double s = traj_at_time(a.tv, base, t).s + q0;
                tj = traj_at_s(a.tv, base, s);
                ps[4 * i + 0] = tj.V; ps[4 * i + 1] = tj.kappa; ps[4 * i + 2] = 0; ps[4 * i + 3] = 0;
            }
        } else if (lane == 0) {
            // cold: forward rollout of (V, s) with steady-state cornering estimates (coupled_lat_long.jl:103-141)
            double s = s0, sp, cp;
            sincos(dpsi, &sp, &cp);
            double V = Ux0 * cp - Uy0 * sp;
            const double beta0 = atan2(Uy0, Ux0);
            double Fyf0, Fyr0;
            {
                double sd, cd;
                sincos(d0, &sd, &cd);
                lateral_tire_forces<double>(P, atan2(Uy0 + P.a * r0, Ux0) - d0, atan2(Uy0 - P.b * r0, Ux0), Fxf0, Fxr0, sd, cd, Fyf0, Fyr0);
            }
            for (int i = 0; i < N; i++) {
                const double tau = (i == N - 1) ? dt[i - 1] : dt[i];
                tj = traj_at_s(a.tv, base, s);
                ds = s - traj_at_time(a.tv, base, ts[i]).s;
                double A_des = tj.A + C.k_V * (tj.V - V) / tau + (path_mode ? 0.0 : -C.k_s * ds / tau / tau);
                A_des = fmin(fmax(A_des, (C.V_min - V) / tau), (C.V_max - V) / tau);
                double A;
                if (i == 0) {
                    double q6[6] = {E0, N0, psi0, Ux0, Uy0, r0}, qd[6];
                    vehicle_model<MODEL_BICYCLE, double>(P, q6, d0, Fx0, 0.0, 0.0, qd);
                    A = (qd[3] - r0 * Uy0) * cp - (qd[4] + r0 * Ux0) * sp;
                } else {
                    SteadyState est;
                    if (i <= Ns) {
                        est = steady_state_estimates(P, V, A_des, tj.kappa, 1, r0, beta0, d0, Fyf0);
                        qs[6 * i + 0] = ds; qs[6 * i + 1] = Ux0; qs[6 * i + 2] = Uy0; qs[6 * i + 3] = r0; qs[6 * i + 4] = adiff(psi0, tj.psi); qs[6 * i + 5] = e0;
                    } else {
                        est = steady_state_estimates(P, V, A_des, tj.kappa, 4, V * tj.kappa, 0.0, 0.0, 0.0);
                        qs[6 * i + 0] = ds; qs[6 * i + 1] = est.Ux; qs[6 * i + 2] = est.Uy; qs[6 * i + 3] = est.r; qs[6 * i + 4] = -est.beta; qs[6 * i + 5] = 0;
                    }
                    us[2 * i + 0] = est.delta; us[2 * i + 1] = est.Fxf + est.Fxr;
                    ps[4 * i + 0] = tj.V; ps[4 * i + 1] = tj.kappa; ps[4 * i + 2] = 0; ps[4 * i + 3] = 0;
                    A = est.A;
                }
                if (i == N - 1) break;
                V = V + A * tau;
                s = s + V * tau + A * tau * tau / 2;
            }
        }
    } else if (lane == 0) {
        a.se0[v] = s0; a.se0[B + v] = e0;          // decoupled: the rollout runs one vehicle per thread in k_nodes_decoupled_rollout
    }
}

// ---- get_next_control ------------------------------------------------------------------------------------------------
// With the guards of src/ros_integration.jl: a paused vehicle (Ux below the threshold, :84-87) and a vehicle whose QP returned NaN
// (:134-147) keep their current control; the latter is also re-initialised (cold ADMM iterates, mpc.solved = false).
__global__ void k_controls(int B, int v0, int nv, int kind, int n_sol, int iv_delta, int iv_fx, double un0, double un1, VehParams P,
                           const double* __restrict__ sol_x, const double* __restrict__ us, int N, double* __restrict__ out,
                           const double* __restrict__ control, const uint8_t* __restrict__ skip, int guard_nan, uint8_t* __restrict__ cold,
                           uint8_t* __restrict__ solved, const double* __restrict__ hji_val, double hji_eps, const double* __restrict__ state,
                           const uint8_t* __restrict__ hold) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    if (hold && hold[v]) return;          // QP still pending (or vehicle finished): no new control in this round
    double d, Fx;
    if (hji_val && hji_val[(size_t)7 * B + v] <= hji_eps) {
        // use_HJI_policy && V <= HJI_eps (ros_integration.jl:115-118): optimal_control replaces the QP's control
        double g[7];
#pragma unroll
        for (int k = 0; k < 7; k++) g[k] = hji_val[(size_t)k * B + v];
        hji_optimal_control(P, state[3 * B + v], state[4 * B + v], state[5 * B + v], g, d, Fx);
    } else if (kind == PGN_COUPLED) { d = sol_x[(size_t)v * n_sol + iv_delta] * un0; Fx = sol_x[(size_t)v * n_sol + iv_fx] * un1; }
    else { d = sol_x[(size_t)v * n_sol + iv_delta]; Fx = us[(size_t)v * N * 2 + 2 * 1 + 1]; }
    double Fxf, Fxr;
    if (Fx > 0) { Fxf = Fx * P.fwd_frac; Fxr = Fx * P.rwd_frac; } else { Fxf = Fx * P.fwb_frac; Fxr = Fx * P.rwb_frac; }
    const bool paused = skip && skip[v];
    const bool bad = guard_nan && !paused && (isnan(d) || isnan(Fxf) || isnan(Fxr));
    if (paused || bad) { d = control[0 * B + v]; Fxf = control[1 * B + v]; Fxr = control[2 * B + v]; }
    if (bad) { cold[v] = 1; solved[v] = 0; }
    out[0 * B + v] = d; out[1 * B + v] = Fxf; out[2 * B + v] = Fxr;
}

// ---- plant rollout: state <- propagate(BicycleModel, state, StepControl(dt, (delta, Fxf+Fxr))); control <- new control --
__global__ void __launch_bounds__(128) k_rollout(int B, VehParams P, double dt, int nsub, double* __restrict__ state, double* __restrict__ control,
                                                 const double* __restrict__ new_control) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    double x[6];
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = state[i * B + v];
    double u[4] = {control[0 * B + v], control[1 * B + v] + control[2 * B + v], 0.0, 0.0};
    flow_rk4<MODEL_BICYCLE, 6, double>(P, x, dt, u, u, nsub);
#pragma unroll
    for (int i = 0; i < 6; i++) state[i * B + v] = x[i];
    if (new_control) {
#pragma unroll
        for (int i = 0; i < 3; i++) control[i * B + v] = new_control[i * B + v];
    }
}

// ---- from_autobox_callback (src/ros_integration.jl:48-151) as one batched entry point ------------------------------------------------
// io buffer (doubles): [0] has_other, then q [B][6], u [B][3], other [B][4], stamp [B]  |  out [B][5] = (delta, Fxf, Fxr, s, e)
// One warp per vehicle: scatter the message fields into the SoA state, pick the MPC time (stamp - time_offset, or the path_coordinates
// time in path-tracking mode, :72-75), flag vehicles whose time lies outside the trajectory (:77-80) and keep (s, e) for the reply (:110).
__global__ void __launch_bounds__(128) k_callback_in(int B, TrajView tv, const int32_t* __restrict__ traj_id, const double* __restrict__ io,
                                                     const double* __restrict__ toff, double* __restrict__ state, double* __restrict__ control,
                                                     double* __restrict__ other, double* __restrict__ t0, uint8_t* __restrict__ tskip, double* __restrict__ se,
                                                     int window, int32_t* __restrict__ last_seg) {
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= B) return;
    const double* q = io + 1 + (size_t)v * 6;
    const double* u = io + 1 + (size_t)B * 6 + (size_t)v * 3;
    const double* o = io + 1 + (size_t)B * 9 + (size_t)v * 4;
    const double stamp = io[1 + (size_t)B * 13 + v];
    if (lane < 6) state[(size_t)lane * B + v] = q[lane];
    if (lane < 3) control[(size_t)lane * B + v] = u[lane];
    if (io[0] != 0.0 && lane < 4) other[(size_t)lane * B + v] = o[lane];
    double s, e, tp;
    {   // a stream of callbacks follows one vehicle: the window (if any) is centred on the previous callback's segment
        const int last = window > 0 ? last_seg[v] : -1;
        int seg;
        if (last >= 0) path_coordinates_warp(tv, traj_id[v] * tv.n_nodes, q[0], q[1], lane, s, e, &tp, last - window, last + window + 1, &seg);
        else path_coordinates_warp(tv, traj_id[v] * tv.n_nodes, q[0], q[1], lane, s, e, &tp, 0, 0x7fffffff, &seg);
        if (lane == 0) last_seg[v] = seg;
    }
    if (lane == 0) {
        const double off = toff[v];
        double t = tp;
        bool out_of_interval = false;
        if (!isnan(off)) {
            t = stamp - off;
            out_of_interval = t < 0 || t > __ldg(tv.f[0] + traj_id[v] * tv.n_nodes + tv.n_nodes - 1);
        }
        t0[v] = t; tskip[v] = out_of_interval;
        se[v] = s; se[B + v] = e;
    }
}
__global__ void k_callback_out(int B, const double* __restrict__ controls, const double* __restrict__ se, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    out[(size_t)v * 5 + 0] = controls[v]; out[(size_t)v * 5 + 1] = controls[B + v]; out[(size_t)v * 5 + 2] = controls[2 * B + v];
    out[(size_t)v * 5 + 3] = se[v]; out[(size_t)v * 5 + 4] = se[B + v];
}

// ---- layout helpers --------------------------------------------------------------------------------------------------
__global__ void k_transpose_in(int B, int k, const double* __restrict__ aos, double* __restrict__ soa) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * k) return;
    const int v = idx / k, f = idx - v * k;
    soa[(size_t)f * B + v] = aos[idx];
}
__global__ void k_transpose_out(int B, int k, const double* __restrict__ soa, double* __restrict__ aos) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * k) return;
    const int v = idx / k, f = idx - v * k;
    aos[idx] = soa[(size_t)f * B + v];
}
// t[i] = base[i] + k * dt  (the reference iterates a range `0:dt:T`, i.e. t_k = k*dt, not an accumulated sum)
__global__ void k_time_axpy(int n, const double* __restrict__ base, double k, double dt, double* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = __dadd_rn(base[i], __dmul_rn(k, dt));   // no FMA contraction: bitwise the host's t0 + k*dt
}

// ---- host inputs: one packed buffer [q [B][6] | u [B][3] | other [B][4] | toff [B] | t0 [B]] -> SoA ------------------------------------
__global__ void k_unpack_state(int B, int v0, int nv, int flags, const double* __restrict__ in, double* __restrict__ state, double* __restrict__ control,
                               double* __restrict__ other, double* __restrict__ toff, double* __restrict__ t0, int32_t* __restrict__ last_seg) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    const size_t Bs = (size_t)B;
    if (flags & 1) {
#pragma unroll
        for (int f = 0; f < 6; f++) state[f * Bs + v] = in[(size_t)v * 6 + f];
        last_seg[v] = -1;                 // a new measured state may be anywhere on the path: the next closest-segment search scans everything
    }
    if (flags & 2) {
#pragma unroll
        for (int f = 0; f < 3; f++) control[f * Bs + v] = in[6 * Bs + (size_t)v * 3 + f];
    }
    if (flags & 4) {
#pragma unroll
        for (int f = 0; f < 4; f++) other[f * Bs + v] = in[9 * Bs + (size_t)v * 4 + f];
    }
    if (flags & 8) toff[v] = in[13 * Bs + v];
    if (flags & 16) t0[v] = in[14 * Bs + v];
}
// mpc.solved = false / Parametron.initialize! for the masked vehicles, on the handle's stream (no host round trip)
__global__ void k_masked_reset(int B, const uint8_t* __restrict__ mask, int what, uint8_t* __restrict__ solved, double* __restrict__ ws_xz,
                               double* __restrict__ ws_y, double* __restrict__ rho, int Nk, double rho0) {
    const int v = blockIdx.x;
    if (mask && !mask[v]) return;
    if ((what & 1) && threadIdx.x == 0) solved[v] = 0;
    if (what & 2) {
        for (int p = threadIdx.x; p < Nk; p += blockDim.x) { ws_xz[(size_t)v * Nk + p] = 0.0; ws_y[(size_t)v * Nk + p] = 0.0; }
        if (threadIdx.x == 0) rho[v] = rho0;
    }
}
// history recorder of `simulate` (model_predictive_control.jl:84-99: qs / us before the step, xs = mpc.qs[1], ps = mpc.ps[1] after node generation)
__global__ void k_record(int B, int v0, int nv, int nx, int N, const double* __restrict__ state, const double* __restrict__ control,
                         const double* __restrict__ qs, const double* __restrict__ ps, double* __restrict__ slot,
                         const uint8_t* __restrict__ hold, const int32_t* __restrict__ kstep, int stride, int cap) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    const size_t Bs = (size_t)B;
    if (hold) {      // deferred solves: every vehicle is at its own step; `slot` is the base of the history buffer
        if (hold[v]) return;
        const int k = kstep[v];
        if (k % stride != 0 || k / stride >= cap) return;
        slot += (size_t)(k / stride) * (13 + nx) * Bs;
    }
    for (int f = 0; f < 6; f++) slot[f * Bs + v] = state[f * Bs + v];
    for (int f = 0; f < 3; f++) slot[(6 + f) * Bs + v] = control[f * Bs + v];
    for (int f = 0; f < nx; f++) slot[(9 + f) * Bs + v] = qs[(size_t)v * N * nx + f];
    for (int f = 0; f < 4; f++) slot[(9 + nx + f) * Bs + v] = ps[(size_t)v * N * 4 + f];
}
__global__ void k_pack_out(int B, int v0, int nv, int k, const double* __restrict__ soa, double* __restrict__ aos) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nv * k) return;
    const int iv = idx / k, f = idx - iv * k;
    aos[(size_t)(v0 + iv) * k + f] = soa[(size_t)f * B + v0 + iv];
}

// ---- launchers -------------------------------------------------------------------------------------------------------
void launch_time_steps(pgn_handle* h, const double* d_t0) {
    const int B = h->nv;          // per-vehicle rows only: a part is an offset into every array
    const size_t o = (size_t)h->v0;
    const bool guarded = h->guard_pause > 0.0 || h->in_callback;
    k_time_steps<<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->cfg.N_short, h->cfg.N_long, h->cfg.dt_short, h->cfg.dt_long, h->cfg.use_correction_step,
                                                         d_t0 + o, h->d_ts + o * h->N, h->d_dt + o * h->T, h->d_prev_ts + o * h->N,
                                                         guarded ? h->d_skip + o : nullptr, h->d_state + 3 * (size_t)h->B + o, h->guard_pause,
                                                         h->in_callback ? h->d_tskip + o : nullptr, h->hold_on ? h->d_hold + o : nullptr);
    h->launches++;
}
void launch_nodes(pgn_handle* h) {
    NodeArgs a;
    a.B = h->B; a.N = h->N; a.Ns = h->cfg.N_short; a.kind = h->cfg.kind; a.n_sol = h->tab.n;
    a.v0 = h->v0; a.nv = h->nv;
    a.P = h->veh; a.C = h->ctl; a.un0 = h->un[0]; a.un1 = h->un[1];
    a.tv = h->traj;
    a.state = h->d_state; a.control = h->d_control; a.toff = h->d_toff; a.solved = h->d_solved; a.traj_id = h->d_traj_id;
    a.ts = h->d_ts; a.dt = h->d_dt; a.prev_ts = h->d_prev_ts; a.sol_x = h->d_sol_x;
    a.qs = h->d_qs; a.us = h->d_us; a.ps = h->d_ps;
    a.skip = h->d_skip; a.pause_below_speed = h->guard_pause; a.tskip = h->in_callback ? h->d_tskip : nullptr;
    a.window = h->path_window; a.last_seg = h->d_last_seg; a.se0 = h->d_se0;
    a.hold = h->hold_on ? h->d_hold : nullptr;
    k_nodes<<<(h->nv + 3) / 4, 128, 0, h->stream>>>(a);
    h->launches++;
    if (h->cfg.kind == PGN_DECOUPLED) {
        k_nodes_decoupled_rollout<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(a, h->d_se0);
        h->launches++;
    }
}
void launch_callback_in(pgn_handle* h) {
    k_callback_in<<<(h->B + 3) / 4, 128, 0, h->stream>>>(h->B, h->traj, h->d_traj_id, h->d_io, h->d_toff, h->d_state, h->d_control, h->d_other, h->d_t0, h->d_tskip, h->d_se, h->path_window, h->d_last_seg);
    h->launches++;
}
void launch_callback_out(pgn_handle* h) {
    k_callback_out<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->d_controls, h->d_se, h->d_io + 1 + (size_t)h->B * 14);
    h->launches++;
}
void launch_controls(pgn_handle* h, double* d_out) {
    const int B = h->B;
    k_controls<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(B, h->v0, h->nv, h->cfg.kind, h->tab.n, h->tab.var_u1_delta, h->tab.var_u1_fx, h->un[0], h->un[1], h->veh,
                                                       h->d_sol_x, h->d_us, h->N, d_out, h->d_control, (h->guard_pause > 0.0 || h->in_callback) ? h->d_skip : nullptr, h->guard_nan,
                                                       h->d_cold, h->d_solved, (h->hji_policy && h->cfg.kind == PGN_COUPLED) ? h->d_hji_val : nullptr, h->cfg.hji_eps, h->d_state,
                                                       h->hold_on ? h->d_hold : nullptr);
    h->launches++;
}
void launch_rollout(pgn_handle* h, double dt) {
    const int B = h->B;
    k_rollout<<<(B + 127) / 128, 128, 0, h->stream>>>(B, h->veh, dt, h->cfg.rk4_substeps, h->d_state, h->d_control, h->d_controls);
    h->launches++;
}
// The plant step only needs the state and the control that was applied DURING the interval, both known before the QP is solved, so it
// can run beside the ADMM launch: propagate into a shadow state on the side stream, commit (state <- shadow, control <- new control) on
// the main stream once both are done.
__global__ void __launch_bounds__(128) k_propagate_shadow(int B, int v0, int nv, VehParams P, double dt, int nsub, const double* __restrict__ state, const double* __restrict__ control,
                                                          double* __restrict__ state_next, const uint8_t* hold) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    // runs beside the ADMM kernel, which moves hold[v] between 0 and 1 only: 2 (vehicle finished, set before the fork) is stable.  A vehicle
    // whose QP stays pending is propagated again in a later round from the same state and control: same result.
    if (hold && hold[v] == 2) return;
    double x[6];
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = state[i * B + v];
    double u[4] = {control[0 * B + v], control[1 * B + v] + control[2 * B + v], 0.0, 0.0};
    flow_rk4<MODEL_BICYCLE, 6, double>(P, x, dt, u, u, nsub);
#pragma unroll
    for (int i = 0; i < 6; i++) state_next[i * B + v] = x[i];
}
__global__ void k_commit_rollout(int B, int v0, int nv, const double* __restrict__ state_next, const double* __restrict__ new_control, double* __restrict__ state, double* __restrict__ control,
                                 const uint8_t* __restrict__ hold, int32_t* __restrict__ kstep) {
    const int iv = blockIdx.x * blockDim.x + threadIdx.x;
    if (iv >= nv) return;
    const int v = v0 + iv;
    if (hold) {
        if (hold[v]) return;              // the vehicle's step is not complete (QP pending) or it has none left
        kstep[v]++;
    }
#pragma unroll
    for (int i = 0; i < 6; i++) state[i * B + v] = state_next[i * B + v];
#pragma unroll
    for (int i = 0; i < 3; i++) control[i * B + v] = new_control[i * B + v];
}
void launch_propagate_shadow(pgn_handle* h, double dt, cudaStream_t side) {
    const int B = h->B;
    k_propagate_shadow<<<(h->nv + 127) / 128, 128, 0, side>>>(B, h->v0, h->nv, h->veh, dt, h->cfg.rk4_substeps, h->d_state, h->d_control, h->d_state_next,
                                                              h->hold_on ? h->d_hold : nullptr);
    h->launches++;
}
void launch_commit_rollout(pgn_handle* h) {
    const int B = h->B;
    k_commit_rollout<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(B, h->v0, h->nv, h->d_state_next, h->d_controls, h->d_state, h->d_control,
                                                                 h->hold_on ? h->d_hold : nullptr, h->d_kstep);
    h->launches++;
}
void launch_transpose_in(pgn_handle* h, const double* d_aos, double* d_soa, int k) {
    const int n = h->B * k;
    k_transpose_in<<<(n + 255) / 256, 256, 0, h->stream>>>(h->B, k, d_aos, d_soa);
    h->launches++;
}
void launch_transpose_out(pgn_handle* h, const double* d_soa, double* d_aos, int k) {
    const int n = h->B * k;
    k_transpose_out<<<(n + 255) / 256, 256, 0, h->stream>>>(h->B, k, d_soa, d_aos);
    h->launches++;
}
void launch_unpack_state(pgn_handle* h, int flags) {
    k_unpack_state<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, 0, h->B, flags, h->d_in, h->d_state, h->d_control, h->d_other, h->d_toff, h->d_t0, h->d_last_seg);
    h->launches++;
}
void launch_unpack_range(pgn_handle* h, const double* d_in, int flags) {
    k_unpack_state<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(h->B, h->v0, h->nv, flags, d_in, h->d_state, h->d_control, h->d_other, h->d_toff, h->d_t0, h->d_last_seg);
    h->launches++;
}
void launch_masked_reset(pgn_handle* h, const uint8_t* d_mask, int what) {
    k_masked_reset<<<h->B, 128, 0, h->stream>>>(h->B, d_mask, what, h->d_solved, h->d_ws_xz, h->d_ws_y, h->d_rho, h->tab.Nk, h->cfg.rho);
    h->launches++;
}
void launch_record(pgn_handle* h, int slot) {
    const size_t rec = (size_t)(13 + h->nx) * h->B;
    if (h->hold_on) k_record<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(h->B, h->v0, h->nv, h->nx, h->N, h->d_state, h->d_control, h->d_qs, h->d_ps, h->d_hist, h->d_hold, h->d_kstep, h->hist_stride, h->hist_cap);
    else k_record<<<(h->nv + 127) / 128, 128, 0, h->stream>>>(h->B, h->v0, h->nv, h->nx, h->N, h->d_state, h->d_control, h->d_qs, h->d_ps, h->d_hist + rec * slot, nullptr, nullptr, 1, 0);
    h->launches++;
}
// simulate loops with deferred solves: per vehicle, the time of its own next step and whether it takes part in this round
__global__ void k_round_begin(int n, const double* __restrict__ base, double dt, const int* __restrict__ target_p, const int32_t* __restrict__ kstep, uint8_t* __restrict__ hold, double* __restrict__ t0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = kstep[i], target = *target_p;              // the target lives in device memory: the round's graph does not change from call to call
    if (hold[i] != 1) hold[i] = k >= target ? 2 : 0;         // 1 = a solve continues: untouched
    t0[i] = __dadd_rn(base[i], __dmul_rn((double)k, dt));    // t_k = t0 + k*dt as the host computes it (no FMA contraction)
}
void launch_round_begin(pgn_handle* h, double dt) {
    const size_t o = (size_t)h->v0;
    k_round_begin<<<(h->nv + 255) / 256, 256, 0, h->stream>>>(h->nv, h->d_t0_base + o, dt, h->d_lag + 1, h->d_kstep + o, h->d_hold + o, h->d_t0 + o);
    h->launches++;
}
__global__ void k_fill_i32(int32_t* __restrict__ d, int value, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = value;
}
// profiling 3: a time stamp between the stages of a round (record: %globaltimer, 0, part | stage << 8 | 1 << 63)
__global__ void k_stamp(unsigned long long* __restrict__ trace, int* __restrict__ trace_n, int cap, unsigned long long meta) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const int slot = atomicAdd(trace_n, 1);
    if (slot < cap) { trace[3 * slot] = t; trace[3 * slot + 1] = 0; trace[3 * slot + 2] = meta; }
}
void launch_stamp(pgn_handle* h, int stage) {
    if (h->profiling != 3 || !h->d_trace) return;
    k_stamp<<<1, 1, 0, h->stream>>>(h->d_trace, h->d_trace_n, h->trace_cap, (unsigned long long)h->part | ((unsigned long long)stage << 8) | (1ull << 63));
}
void launch_fill_i32(pgn_handle* h, int32_t* d, int value, int n) {
    k_fill_i32<<<(n + 255) / 256, 256, 0, h->stream>>>(d, value, n);
    h->launches++;
}
// vehicles that have not reached `target` steps yet (the catch-up rounds of a simulate loop with deferred solves run until this is 0)
__global__ void k_count_lag(int B, const int32_t* __restrict__ kstep, int target, int* __restrict__ out) {
    __shared__ int s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    int c = 0;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < B; v += gridDim.x * blockDim.x) c += kstep[v] < target;
    if (c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0 && s) atomicAdd(out, s);
}
void launch_count_lag(pgn_handle* h, int target) {
    cudaMemsetAsync(h->d_lag, 0, sizeof(int), h->stream);
    k_count_lag<<<std::min(64, (h->B + 255) / 256), 256, 0, h->stream>>>(h->B, h->d_kstep, target, h->d_lag);
    h->launches++;
}
void launch_pack_out(pgn_handle* h, const double* d_soa, double* d_aos, int k) {
    const int n = h->nv * k;
    k_pack_out<<<(n + 255) / 256, 256, 0, h->stream>>>(h->B, h->v0, h->nv, k, d_soa, d_aos);
    h->launches++;
}
void launch_time_axpy(pgn_handle* h, const double* d_base, double k, double dt, double* d_v, int n) {
    k_time_axpy<<<(n + 255) / 256, 256, 0, h->stream>>>(n, d_base, k, dt, d_v);
    h->launches++;
}

}  // namespace pgn

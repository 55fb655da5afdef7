// oracle/mpc.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
//
// CPU restatement of the per-time-step MPC path of the reference:
//   /root/reference/src/model_predictive_control.jl:1-30 (MPCTimeSteps), :32-78 (TrajectoryTrackingMPC + step API), :80-100 (simulate)
//   /root/reference/src/coupled_lat_long.jl:23-38 (defaults), :62-142 (nodes), :197-313 (QP structure), :315-368 (update_QP!), :370-374
//   /root/reference/src/decoupled_lat_long.jl:18-28, :52-104, :134-226, :228-273, :275-278
// The QP is kept in the canonical form  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u  with variables and rows in the reference's
// construction order (documented in DESIGN.md); Parametron/MOI may permute or negate rows, which leaves the ADMM iterates
// unchanged up to rounding.
#pragma once
#include <cmath>
#include <memory>
#include <vector>
#include "hji.hpp"
#include "linearize.hpp"
#include "osqp_port.hpp"
#include "trajectory.hpp"
#include "vehicle.hpp"

namespace orc {

struct ControlParams {
    // CoupledControlParams (coupled_lat_long.jl:23-38); decoupled uses the subset (decoupled_lat_long.jl:18-28)
    double V_min = 1.0, V_max = 15.0, k_V = 10.0 / 4 / 100, k_s = 10.0 / 4 / 10000, delta_dot_max = 0.344;
    double Q_ds = 1.0, Q_dpsi = 1.0, Q_e = 1.0, W_beta = 50 / (10 * M_PI / 180), W_r = 50.0, W_HJI = 500.0;
    int N_HJI = 3;
    double R_delta = 0.0, R_ddelta = 0.1, R_Fx = 0.0, R_dFx = 0.5;
    static ControlParams coupled_defaults() { return ControlParams(); }
    static ControlParams decoupled_defaults() {
        ControlParams c;
        c.Q_dpsi = 1 / std::pow(10 * M_PI / 180, 2); c.Q_e = 1.0; c.R_delta = 0.0; c.R_ddelta = 0.01 / std::pow(10 * M_PI / 180, 2);
        return c;
    }
};
static const int CONTROL_PARAMS_LEN = 16;

// MPCTimeSteps (model_predictive_control.jl:1-30)
struct TimeSteps {
    int N_short, N_long; double dt_short, dt_long; bool use_correction_step;
    std::vector<double> ts, dt, prev_ts;
    TimeSteps(int ns, int nl, double ds, double dl, bool corr) : N_short(ns), N_long(nl), dt_short(ds), dt_long(dl), use_correction_step(corr) {
        int N = 1 + ns + nl;
        ts.resize(N); for (int i = 0; i < N; i++) ts[i] = i + 1;
        dt.assign(N - 1, 1.0); prev_ts = ts;
    }
    void compute(double t0) {
        prev_ts = ts;
        double t0_long = t0 + N_short * dt_short;
        if (use_correction_step) t0_long = dt_long * std::ceil((t0_long + dt_short) / dt_long - 1);
        for (int i = 0; i <= N_short; i++) ts[i] = t0 + dt_short * i;
        for (int i = 1; i <= N_long; i++) ts[N_short + i] = t0_long + dt_long * i;
        for (size_t i = 0; i + 1 < ts.size(); i++) dt[i] = ts[i + 1] - ts[i];
    }
};

enum MpcKind { MPC_COUPLED = 0, MPC_DECOUPLED = 1 };

class Mpc {
public:
    int kind;
    VehicleParams veh;
    ControlParams cp;
    TimeSteps TS;
    TrajectoryTube traj;
    double state[6] = {0}, control[3] = {0}, other_car[4] = {0};
    double time_offset = NAN;
    bool solved = false;
    HjiCache hji = placeholder_hji();
    double hji_eps = 0.05;
    bool use_hji_policy = false;   // use_HJI_policy[] of ros_integration.jl:47
    int N, T, Ns, nx, nu;          // nodes, intervals, short steps, state dim (6|4), QP control dim (2|1)
    std::vector<double> qs, us, ps;  // nodes: qs[N*nx], us[N*2], ps[N*4]
    // QP pieces (per interval t), kept for introspection / parity tests
    std::vector<DiscreteLin> lin;      // T
    std::vector<StableLimits> env;     // T
    std::vector<double> dmin, dmax, fxmax;  // T (normalised)
    double M_hji[2] = {0, 0}, b_hji = 1.0, hji_V = INFINITY, hji_gradV[7] = {0};
    double un[2];
    // canonical QP
    int n, m;
    std::vector<int> tr_r, tr_c; std::vector<double> tr_v;    // A triplets in construction order
    std::vector<int> tr_to_csc;
    Csc Pm, Am; std::vector<double> qv, lv, uv;
    OsqpSolver solver; OsqpSettings settings; bool solver_ready = false;

    Mpc(int kind_, const VehicleParams& v, const ControlParams& c, int ns, int nl, double dts, double dtl, bool corr, const OsqpSettings& s)
        : kind(kind_), veh(v), cp(c), TS(ns, nl, dts, dtl, corr), settings(s) {
        N = 1 + ns + nl; T = N - 1; Ns = ns;
        nx = kind == MPC_COUPLED ? 6 : 4; nu = kind == MPC_COUPLED ? 2 : 1;
        qs.assign(N * nx, 0); us.assign(N * 2, 0); ps.assign(N * 4, 0);
        lin.resize(T); env.resize(T); dmin.assign(T, 0); dmax.assign(T, 0); fxmax.assign(T, 0);
        un[0] = veh.delta_max; un[1] = std::max(-veh.Fx_min, veh.Fx_max);
        traj = straight_trajectory(30., 5.);
        build_pattern();
    }

    // ---- variable indexing ----
    int vq(int i, int t) const { return nx * t + i; }
    int vu(int i, int t) const { return nx * N + nu * t + i; }
    int vsig(int i, int t) const { return (nx + nu) * N + 2 * t + i; }
    int vsh(int t) const { return (nx + nu) * N + 2 * T + t; }                                     // coupled only
    int vdd(int t) const { return (nx + nu) * N + 2 * T + (kind == MPC_COUPLED ? Ns : 0) + t; }
    int vdf(int t) const { return (nx + nu) * N + 2 * T + Ns + T + t; }                           // coupled only

    void compute_time_steps(double t0) { TS.compute(t0); }

    // ---- compute_linearization_nodes! ----
    void compute_linearization_nodes() {
        if (kind == MPC_COUPLED) nodes_coupled(); else nodes_decoupled();
    }
    // ---- update_QP! ----
    void update_qp() {
        const std::vector<double>& dt = TS.dt;
        for (int t = 0; t < T; t++) {
            double up0[6] = {us[2 * t], us[2 * t + 1], ps[4 * t], ps[4 * t + 1], ps[4 * t + 2], ps[4 * t + 3]};
            double upf[6] = {us[2 * t + 2], us[2 * t + 3], ps[4 * t + 4], ps[4 * t + 5], ps[4 * t + 6], ps[4 * t + 7]};
            bool ramp = t >= Ns;
            if (kind == MPC_COUPLED) {
                lin[t] = linearize_flow(MODEL_TRACKING, veh, &qs[nx * t], dt[t], up0, upf, ramp, 2);
            } else {
                ContinuousLin CL = linearize_continuous(MODEL_LATERAL, veh, &qs[nx * t], up0);
                lin[t] = linearize_exact(CL, &qs[nx * t], dt[t], up0, upf, ramp, 1);
            }
            double Uxt = kind == MPC_COUPLED ? qs[nx * (t + 1) + 1] : ps[4 * (t + 1)];
            double Fxft, Fxrt;
            longitudinal_tire_forces<double>(veh, us[2 * (t + 1) + 1], Fxft, Fxrt);
            env[t] = stable_limits(veh, Uxt, Fxft, Fxrt);
            double dn = kind == MPC_COUPLED ? un[0] : 1.0;
            dmin[t] = std::max(env[t].delta_min, -veh.delta_max) / dn;
            dmax[t] = std::min(env[t].delta_max, veh.delta_max) / dn;
            fxmax[t] = std::min(veh.Px_max / Uxt, veh.Fx_max) / un[1];
        }
        if (kind == MPC_COUPLED) {
            double x7[7];
            hji_relative_state(state, other_car, x7);
            double uR[2] = {control[0], control[1] + control[2]};
            reachability_constraint(veh, hji, x7, hji_eps, uR, M_hji, b_hji, &hji_V, hji_gradV);
        }
        fill_values();
    }
    // ---- solve! ----
    int solve() {
        if (!solver_ready) { solver.setup(Pm, qv, Am, lv, uv, settings); solver_ready = true; }
        else solver.update(Pm.x.data(), Am.x.data(), qv.data(), lv.data(), uv.data());
        int st = solver.solve();
        solved = true;
        return st;
    }
    // Parametron.initialize! equivalent: fresh OSQP workspace (cold iterates, rho back to its setting)
    void reset_solver() { solver_ready = false; }
    // ---- get_next_control ----
    // use_hji_policy: the callback's override (ros_integration.jl:115-118): V <= HJI_eps => BicycleControl(LP, optimal_control(...))
    void get_next_control(double* out3) const {
        double d, Fx;
        if (kind == MPC_COUPLED && use_hji_policy && hji_V <= hji_eps) {
            double x7[7], u2[2];
            hji_relative_state(state, other_car, x7);
            optimal_control(veh, x7, hji_gradV, u2);
            d = u2[0]; Fx = u2[1];
        } else if (kind == MPC_COUPLED) { d = solver.sol_x[vu(0, 1)] * un[0]; Fx = solver.sol_x[vu(1, 1)] * un[1]; }
        else { d = solver.sol_x[vu(0, 1)]; Fx = us[2 * 1 + 1]; }
        out3[0] = d;
        longitudinal_tire_forces<double>(veh, Fx, out3[1], out3[2]);
    }
    // from_autobox_callback (ros_integration.jl:48-151) without the ROS plumbing.  Returns false where the callback returns early
    // (time outside the trajectory :77-80, Ux < pause_speed :84-87; the reference hard-codes 1 m/s, 0 disables the check here);
    // out5 = (delta, Fxf, Fxr, s, e); on an early return the control part is the current control.
    bool from_autobox(const double* q6, const double* u3, const double* other4, double stamp, double pause_speed, bool nan_fallback, double* out5) {
        for (int i = 0; i < 6; i++) state[i] = q6[i];
        for (int i = 0; i < 3; i++) control[i] = u3[i];
        if (other4) for (int i = 0; i < 4; i++) other_car[i] = other4[i];
        double s, e, tp;
        traj.path_coordinates(state[0], state[1], s, e, tp);
        out5[0] = control[0]; out5[1] = control[1]; out5[2] = control[2]; out5[3] = s; out5[4] = e;
        double t = tp;                                                  // path tracking mode (:72-75)
        if (!std::isnan(time_offset)) {
            t = stamp - time_offset;
            if (t < 0 || t > traj.t.back()) return false;
        }
        if (pause_speed > 0 && state[3] < pause_speed) return false;
        compute_time_steps(t);
        compute_linearization_nodes();
        update_qp();
        solve();
        double u[3];
        get_next_control(u);
        if (nan_fallback && (std::isnan(u[0]) || std::isnan(u[1]) || std::isnan(u[2]))) {      // :134-147
            reset_solver();
            solved = false;
            return true;
        }
        out5[0] = u[0]; out5[1] = u[1]; out5[2] = u[2];
        return true;
    }
    // one closed-loop step of simulate() (model_predictive_control.jl:87-98)
    void simulate_step(double t, double dt_sim) {
        compute_time_steps(t);
        compute_linearization_nodes();
        update_qp();
        solve();
        double u2[6] = {control[0], control[1] + control[2], 0, 0, 0, 0}, xn[6];
        flow_rk4<double>(MODEL_BICYCLE, veh, state, dt_sim, u2, u2, xn);
        for (int i = 0; i < 6; i++) state[i] = xn[i];
        get_next_control(control);
    }

private:
    void add(int r, int c) { tr_r.push_back(r); tr_c.push_back(c); }
    void build_pattern() {
        int r = 0;
        if (kind == MPC_COUPLED) {
            n = 8 * N + 4 * T + Ns;
            for (int t = 0; t < T; t++) for (int i = 0; i < 2; i++) add(r++, vsig(i, t));
            for (int t = 0; t < Ns; t++) add(r++, vsh(t));
            for (int t = 0; t < T; t++) { add(r, vu(0, t + 1)); add(r, vu(0, t)); add(r, vdd(t)); r++; }
            for (int t = 0; t < T; t++) { add(r, vu(1, t + 1)); add(r, vu(1, t)); add(r, vdf(t)); r++; }
            for (int t = 0; t < N; t++) add(r++, vq(1, t));
            for (int t = 0; t < N; t++) add(r++, vq(1, t));
            for (int t = 0; t < N; t++) add(r++, vu(1, t));
            for (int i = 0; i < 6; i++) add(r++, vq(i, 0));
            for (int i = 0; i < 2; i++) add(r++, vu(i, 0));
            for (int t = 0; t < Ns; t++) for (int i = 0; i < 6; i++) {
                for (int j = 0; j < 6; j++) add(r, vq(j, t));
                for (int k = 0; k < 2; k++) add(r, vu(k, t));
                add(r, vq(i, t + 1)); r++;
            }
            for (int t = 0; t < Ns; t++) { add(r, vu(0, t)); add(r, vu(1, t)); add(r, vsh(t)); r++; }
            for (int t = Ns; t < T; t++) for (int i = 0; i < 6; i++) {
                for (int j = 0; j < 6; j++) add(r, vq(j, t));
                for (int k = 0; k < 2; k++) add(r, vu(k, t));
                for (int k = 0; k < 2; k++) add(r, vu(k, t + 1));
                add(r, vq(i, t + 1)); r++;
            }
            for (int t = 0; t < T; t++) {
                add(r++, vu(0, t + 1)); add(r++, vu(0, t + 1)); add(r++, vu(1, t + 1));
                for (int k = 0; k < 4; k++) { add(r, vq(2, t + 1)); add(r, vq(3, t + 1)); add(r, vsig(k / 2, t)); r++; }
                add(r++, vdd(t)); add(r++, vdd(t));
            }
        } else {
            n = 5 * N + 3 * T;
            for (int t = 0; t < T; t++) for (int i = 0; i < 2; i++) add(r++, vsig(i, t));
            for (int t = 0; t < T; t++) { add(r, vu(0, t + 1)); add(r, vu(0, t)); add(r, vdd(t)); r++; }
            for (int i = 0; i < 4; i++) add(r++, vq(i, 0));
            add(r++, vu(0, 0));
            for (int t = 0; t < Ns; t++) for (int i = 0; i < 4; i++) {
                for (int j = 0; j < 4; j++) add(r, vq(j, t));
                add(r, vu(0, t));
                add(r, vq(i, t + 1)); r++;
            }
            for (int t = Ns; t < T; t++) for (int i = 0; i < 4; i++) {
                for (int j = 0; j < 4; j++) add(r, vq(j, t));
                add(r, vu(0, t)); add(r, vu(0, t + 1));
                add(r, vq(i, t + 1)); r++;
            }
            for (int t = 0; t < T; t++) {
                add(r++, vu(0, t + 1)); add(r++, vu(0, t + 1));
                for (int k = 0; k < 4; k++) { add(r, vq(0, t + 1)); add(r, vq(1, t + 1)); add(r, vsig(k / 2, t)); r++; }
                add(r++, vdd(t)); add(r++, vdd(t));
            }
        }
        m = r;
        tr_v.assign(tr_r.size(), 0.0);
        // triplets -> CSC (duplicates: the dynamics row i touches q(i,t+1) only once, and q(j,t) are distinct => none)
        Am.nrow = m; Am.ncol = n; Am.p.assign(n + 1, 0);
        for (size_t e = 0; e < tr_c.size(); e++) Am.p[tr_c[e] + 1]++;
        for (int j = 0; j < n; j++) Am.p[j + 1] += Am.p[j];
        Am.i.assign(tr_r.size(), 0); Am.x.assign(tr_r.size(), 0.0); tr_to_csc.assign(tr_r.size(), 0);
        std::vector<std::vector<std::pair<int, int>>> cols(n);
        for (size_t e = 0; e < tr_c.size(); e++) cols[tr_c[e]].push_back({tr_r[e], (int)e});
        for (int j = 0; j < n; j++) {
            std::sort(cols[j].begin(), cols[j].end());
            int pos = Am.p[j];
            for (auto& pr : cols[j]) { Am.i[pos] = pr.first; tr_to_csc[pr.second] = pos; pos++; }
        }
        Pm.nrow = Pm.ncol = n; Pm.p.resize(n + 1); Pm.i.resize(n); Pm.x.assign(n, 0.0);
        for (int j = 0; j < n; j++) { Pm.p[j] = j; Pm.i[j] = j; }
        Pm.p[n] = n;
        qv.assign(n, 0.0); lv.assign(m, 0.0); uv.assign(m, 0.0);
    }

    void fill_values() {
        const std::vector<double>& dt = TS.dt;
        const double INF = INFINITY;
        std::fill(Pm.x.begin(), Pm.x.end(), 0.0); std::fill(qv.begin(), qv.end(), 0.0);
        size_t e = 0; int r = 0;
        auto put = [&](double v) { tr_v[e++] = v; };
        if (kind == MPC_COUPLED) {
            for (int t = 0; t < T; t++) {
                Pm.x[vq(0, t + 1)] = 2 * cp.Q_ds * dt[t]; Pm.x[vq(4, t + 1)] = 2 * cp.Q_dpsi * dt[t]; Pm.x[vq(5, t + 1)] = 2 * cp.Q_e * dt[t];
                Pm.x[vu(0, t + 1)] = 2 * cp.R_delta * dt[t]; Pm.x[vu(1, t + 1)] = 2 * cp.R_Fx * dt[t];
                Pm.x[vdd(t)] = 2 * cp.R_ddelta / dt[t]; Pm.x[vdf(t)] = 2 * cp.R_dFx / dt[t];
                qv[vsig(0, t)] = cp.W_beta * dt[t]; qv[vsig(1, t)] = cp.W_r * dt[t];
            }
            for (int t = 0; t < Ns; t++) qv[vsh(t)] = t < cp.N_HJI ? cp.W_HJI : 0.0;
            for (int t = 0; t < T; t++) for (int i = 0; i < 2; i++) { put(1); lv[r] = 0; uv[r] = INF; r++; }
            for (int t = 0; t < Ns; t++) { put(1); lv[r] = 0; uv[r] = INF; r++; }
            for (int t = 0; t < T; t++) { put(1); put(-1); put(-1); lv[r] = 0; uv[r] = 0; r++; }
            for (int t = 0; t < T; t++) { put(1); put(-1); put(-1); lv[r] = 0; uv[r] = 0; r++; }
            for (int t = 0; t < N; t++) { put(1); lv[r] = cp.V_min; uv[r] = INF; r++; }
            for (int t = 0; t < N; t++) { put(1); lv[r] = -INF; uv[r] = cp.V_max; r++; }
            for (int t = 0; t < N; t++) { put(1); lv[r] = veh.Fx_min / un[1]; uv[r] = INF; r++; }
            for (int i = 0; i < 6; i++) { put(1); lv[r] = uv[r] = qs[i]; r++; }
            for (int i = 0; i < 2; i++) { put(1); lv[r] = uv[r] = us[i] / un[i]; r++; }
            for (int t = 0; t < Ns; t++) for (int i = 0; i < 6; i++) {
                for (int j = 0; j < 6; j++) put(lin[t].A[i * 6 + j]);
                for (int k = 0; k < 2; k++) put(lin[t].B0[i * 2 + k] * un[k]);
                put(-1); lv[r] = uv[r] = -lin[t].c[i]; r++;
            }
            for (int t = 0; t < Ns; t++) { put(M_hji[0] * un[0]); put(M_hji[1] * un[1]); put(1); lv[r] = -b_hji; uv[r] = INF; r++; }
            for (int t = Ns; t < T; t++) for (int i = 0; i < 6; i++) {
                for (int j = 0; j < 6; j++) put(lin[t].A[i * 6 + j]);
                for (int k = 0; k < 2; k++) put(lin[t].B0[i * 2 + k] * un[k]);
                for (int k = 0; k < 2; k++) put(lin[t].Bf[i * 2 + k] * un[k]);
                put(-1); lv[r] = uv[r] = -lin[t].c[i]; r++;
            }
            for (int t = 0; t < T; t++) {
                put(1); lv[r] = -INF; uv[r] = dmax[t]; r++;
                put(1); lv[r] = dmin[t]; uv[r] = INF; r++;
                put(1); lv[r] = -INF; uv[r] = fxmax[t]; r++;
                for (int k = 0; k < 4; k++) { put(env[t].H[k][0]); put(env[t].H[k][1]); put(-1); lv[r] = -INF; uv[r] = env[t].G[k]; r++; }
                put(1); lv[r] = -INF; uv[r] = cp.delta_dot_max * dt[t] / un[0]; r++;
                put(1); lv[r] = -cp.delta_dot_max * dt[t] / un[0]; uv[r] = INF; r++;
            }
        } else {
            for (int t = 0; t < T; t++) {
                Pm.x[vq(2, t + 1)] = 2 * cp.Q_dpsi * dt[t]; Pm.x[vq(3, t + 1)] = 2 * cp.Q_e * dt[t];
                Pm.x[vu(0, t + 1)] = 2 * cp.R_delta * dt[t]; Pm.x[vdd(t)] = 2 * cp.R_ddelta / dt[t];
                qv[vsig(0, t)] = cp.W_beta * dt[t]; qv[vsig(1, t)] = cp.W_r * dt[t];
            }
            for (int t = 0; t < T; t++) for (int i = 0; i < 2; i++) { put(1); lv[r] = 0; uv[r] = INF; r++; }
            for (int t = 0; t < T; t++) { put(1); put(-1); put(-1); lv[r] = 0; uv[r] = 0; r++; }
            for (int i = 0; i < 4; i++) { put(1); lv[r] = uv[r] = qs[i]; r++; }
            { put(1); lv[r] = uv[r] = us[0]; r++; }
            for (int t = 0; t < Ns; t++) for (int i = 0; i < 4; i++) {
                for (int j = 0; j < 4; j++) put(lin[t].A[i * 4 + j]);
                put(lin[t].B0[i]);
                put(-1); lv[r] = uv[r] = -lin[t].c[i]; r++;
            }
            for (int t = Ns; t < T; t++) for (int i = 0; i < 4; i++) {
                for (int j = 0; j < 4; j++) put(lin[t].A[i * 4 + j]);
                put(lin[t].B0[i]); put(lin[t].Bf[i]);
                put(-1); lv[r] = uv[r] = -lin[t].c[i]; r++;
            }
            for (int t = 0; t < T; t++) {
                put(1); lv[r] = -INF; uv[r] = dmax[t]; r++;
                put(1); lv[r] = dmin[t]; uv[r] = INF; r++;
                for (int k = 0; k < 4; k++) { put(env[t].H[k][0]); put(env[t].H[k][1]); put(-1); lv[r] = -INF; uv[r] = env[t].G[k]; r++; }
                put(1); lv[r] = -INF; uv[r] = cp.delta_dot_max * dt[t]; r++;
                put(1); lv[r] = -cp.delta_dot_max * dt[t]; uv[r] = INF; r++;
            }
        }
        for (size_t k = 0; k < tr_v.size(); k++) Am.x[tr_to_csc[k]] = tr_v[k];
    }

    void node_params(const TrajectoryNode& tj, double* p4) const { p4[0] = tj.V; p4[1] = tj.kappa; p4[2] = 0; p4[3] = 0; }

    void nodes_coupled() {
        const std::vector<double>&ts = TS.ts, &dt = TS.dt, &pts = TS.prev_ts;
        double s0, e0, t0;
        traj.path_coordinates(state[0], state[1], s0, e0, t0);
        TrajectoryNode tj = traj.at_s(s0);
        double ds = s0 - traj.at_time(ts[0]).s;
        double dpsi = adiff(state[2], tj.psi);
        double q[6] = {ds, state[3], state[4], state[5], dpsi, e0};
        double u[2] = {control[0], control[1] + control[2]};
        double p[4]; node_params(tj, p);
        auto store = [&](int i) { for (int k = 0; k < 6; k++) qs[6 * i + k] = q[k]; for (int k = 0; k < 2; k++) us[2 * i + k] = u[k]; for (int k = 0; k < 4; k++) ps[4 * i + k] = p[k]; };
        if (solved) {
            store(0);
            // update_interpolations!: knots = prev_ts, coefficients = previous QP solution
            const std::vector<double>& X = solver.sol_x;
            for (int i = 1; i < N; i++) {
                double t = ts[i];
                double tq = t < pts[N - 1] ? t : pts[N - 1];
                int k = TrajectoryTube::sslast(pts, tq);
                if (k < 1) k = 1; if (k > N - 1) k = N - 1;
                k -= 1;
                double w = (tq - pts[k]) / (pts[k + 1] - pts[k]);
                for (int c = 0; c < 6; c++) q[c] = (1 - w) * X[vq(c, k)] + w * X[vq(c, k + 1)];
                for (int c = 0; c < 2; c++) u[c] = ((1 - w) * X[vu(c, k)] + w * X[vu(c, k + 1)]) * un[c];
                double s = traj.at_time(t).s + q[0];
                tj = traj.at_s(s);
                node_params(tj, p);
                store(i);
            }
        } else {
            double s = s0;
            double sp = std::sin(dpsi), cpsi = std::cos(dpsi);
            double V = state[3] * cpsi - state[4] * sp;
            double beta0 = std::atan2(state[4], state[3]);
            double r0 = state[5], delta0 = control[0];
            double Fyf0, Fyr0;
            lateral_tire_forces_qu(veh, state, control, Fyf0, Fyr0);
            for (int i = 0; i < N; i++) {
                double tau = (i == N - 1) ? dt[i - 1] : dt[i];
                tj = traj.at_s(s);
                ds = s - traj.at_time(ts[i]).s;
                double A_des = tj.A + cp.k_V * (tj.V - V) / tau + (std::isnan(time_offset) ? 0.0 : -cp.k_s * ds / tau / tau);
                A_des = std::min(std::max(A_des, (cp.V_min - V) / tau), (cp.V_max - V) / tau);
                double A;
                if (i == 0) {
                    double u2[2] = {control[0], control[1] + control[2]}, p4[4] = {tj.psi, tj.kappa, tj.theta, tj.phi}, qd[6];
                    vehicle_model<double>(MODEL_BICYCLE, veh, state, u2, p4, qd);
                    A = (qd[3] - state[5] * state[4]) * cpsi - (qd[4] + state[5] * state[3]) * sp;
                } else if (i <= Ns) {
                    SteadyState est = steady_state_estimates(veh, V, A_des, tj.kappa, 1, r0, beta0, delta0, Fyf0);
                    q[0] = ds; q[1] = state[3]; q[2] = state[4]; q[3] = state[5]; q[4] = adiff(state[2], tj.psi); q[5] = e0;
                    u[0] = est.delta; u[1] = est.Fxf + est.Fxr;
                    node_params(tj, p);
                    A = est.A;
                } else {
                    SteadyState est = steady_state_estimates(veh, V, A_des, tj.kappa, 4, V * tj.kappa, 0.0, 0.0, 0.0);
                    q[0] = ds; q[1] = est.Ux; q[2] = est.Uy; q[3] = est.r; q[4] = -est.beta; q[5] = 0;
                    u[0] = est.delta; u[1] = est.Fxf + est.Fxr;
                    node_params(tj, p);
                    A = est.A;
                }
                store(i);
                if (i == N - 1) break;
                V = V + A * tau;
                s = s + V * tau + A * tau * tau / 2;
            }
        }
    }

    void nodes_decoupled() {
        const std::vector<double>&ts = TS.ts, &dt = TS.dt;
        double s, e0, t0;
        traj.path_coordinates(state[0], state[1], s, e0, t0);
        double V = std::hypot(state[3], state[4]);
        double beta0 = std::atan2(state[4], state[3]);
        double r0 = state[5], delta0 = control[0];
        double Fyf0, Fyr0;
        lateral_tire_forces_qu(veh, state, control, Fyf0, Fyr0);
        for (int i = 0; i < N; i++) {
            double tau = (i == N - 1) ? dt[i - 1] : dt[i];
            TrajectoryNode tj = traj.at_s(s);
            double kappa = tj.kappa;
            double A_des = tj.A + cp.k_V * (tj.V - V) / tau + (std::isnan(time_offset) ? 0.0 : cp.k_s * (traj.at_time(ts[i]).s - s) / tau / tau);
            A_des = std::min(std::max(A_des, (cp.V_min - V) / tau), (cp.V_max - V) / tau);
            double q[4], u[2], p[4], A;
            if (i == 0) {
                q[0] = state[4]; q[1] = state[5]; q[2] = adiff(state[2], tj.psi); q[3] = e0;
                u[0] = control[0]; u[1] = control[1] + control[2];
                p[0] = state[3]; p[1] = kappa; p[2] = 0; p[3] = 0;
                double p4[4] = {tj.psi, tj.kappa, tj.theta, tj.phi}, qd[6];
                vehicle_model<double>(MODEL_BICYCLE, veh, state, u, p4, qd);
                A = (qd[3] - state[5] * state[4]) * std::cos(beta0) + (qd[4] + state[5] * state[3]) * std::sin(beta0);
            } else if (i <= Ns) {
                q[0] = state[4]; q[1] = state[5]; q[2] = adiff(state[2], tj.psi); q[3] = e0;
                SteadyState est = steady_state_estimates(veh, V, A_des, kappa, 1, r0, beta0, delta0, Fyf0);
                u[0] = est.delta; u[1] = est.Fxf + est.Fxr;
                p[0] = est.Ux; p[1] = kappa; p[2] = 0; p[3] = 0;
                A = est.A;
            } else {
                SteadyState est = steady_state_estimates(veh, V, A_des, kappa, 4, V * kappa, 0.0, 0.0, 0.0);
                q[0] = est.Uy; q[1] = est.r; q[2] = -est.beta; q[3] = 0;
                u[0] = est.delta; u[1] = est.Fxf + est.Fxr;
                p[0] = est.Ux; p[1] = kappa; p[2] = 0; p[3] = 0;
                A = est.A;
            }
            for (int k = 0; k < 4; k++) qs[4 * i + k] = q[k];
            for (int k = 0; k < 2; k++) us[2 * i + k] = u[k];
            for (int k = 0; k < 4; k++) ps[4 * i + k] = p[k];
            if (i == N - 1) break;
            V = V + A * tau;
            s = s + V * tau + A * tau * tau / 2;
        }
    }
};

}  // namespace orc

// oracle/trajectory.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
//
// CPU restatement of /root/reference/src/trajectories.jl:1-105 (TrajectoryTube, lookup by time / arclength,
// path_coordinates), src/math.jl:1-9 (invcumtrapz, distance2), src/PigeonViz.jl:24-28 (adiff) and
// src/ros_integration.jl:13-16 (path message -> TrajectoryTube).
// Interpolations.jl 0.11.2 `Gridded(Linear())` + `Line()` extrapolation (not vendored in the reference tree) is
// restated as: interval i = clamp(searchsortedlast(knots, s), 1, n-1), weight unclamped.
#pragma once
#include <cmath>
#include <vector>
#include "vehicle.hpp"

namespace orc {

// mod2piF / adiff (PigeonViz.jl:24-28); Julia mod(x, y) for floats has the sign of y
inline double jl_mod(double x, double y) {
    double r = std::fmod(x, y);
    if (r == 0) return std::copysign(r, y);
    if ((r > 0) != (y > 0)) return r + y;
    return r;
}
inline double adiff(double x, double y) {
    double d = jl_mod(x - y, 2 * M_PI);
    return d <= M_PI ? d : d - 2 * M_PI;
}

struct TrajectoryNode { double t, s, V, A, E, N, psi, kappa, theta, phi, edge_L, edge_R; };

struct TrajectoryTube {
    std::vector<double> t, s, V, A, E, N, psi, kappa, theta, phi, edge_L, edge_R;
    int n() const { return (int)t.size(); }

    // Julia searchsortedfirst(v, x): first index (1-based) with v[i] >= x, n+1 if none. 0-based here: returns count of v[i] < x.
    static int ssfirst(const std::vector<double>& v, double x) {
        int lo = 0, hi = (int)v.size();
        while (lo < hi) { int mid = (lo + hi) / 2; if (v[mid] < x) lo = mid + 1; else hi = mid; }
        return lo;
    }
    // searchsortedlast: last index with v[i] <= x (1-based), 0 if none. 0-based count of v[i] <= x.
    static int sslast(const std::vector<double>& v, double x) {
        int lo = 0, hi = (int)v.size();
        while (lo < hi) { int mid = (lo + hi) / 2; if (v[mid] <= x) lo = mid + 1; else hi = mid; }
        return lo;
    }
    // interp_by_s (trajectories.jl:32-35): 8 spatial interpolants, gridded linear in s with Line() extrapolation
    void interp_by_s(double sq, TrajectoryNode& out) const {
        int nn = n();
        int i = sslast(s, sq);              // 1-based index of last knot <= sq
        if (i < 1) i = 1; if (i > nn - 1) i = nn - 1;
        int k = i - 1;                      // 0-based
        double w = (sq - s[k]) / (s[k + 1] - s[k]);
        auto L = [&](const std::vector<double>& c) { return (1 - w) * c[k] + w * c[k + 1]; };
        out.E = L(E); out.N = L(N); out.psi = L(psi); out.kappa = L(kappa); out.theta = L(theta); out.phi = L(phi);
        out.edge_L = L(edge_L); out.edge_R = L(edge_R);
    }
    // traj(t) (trajectories.jl:47-54)
    TrajectoryNode at_time(double tq) const {
        int nn = n();
        int i = ssfirst(t, tq) + 1 - 1;     // searchsortedfirst (1-based) - 1
        if (i < 1) i = 1; if (i > nn - 1) i = nn - 1;
        int k = i - 1;
        double A_ = (V[k + 1] - V[k]) / (t[k + 1] - t[k]);
        double dt = tq - t[k];
        TrajectoryNode o;
        o.t = tq; o.s = s[k] + V[k] * dt + A_ * dt * dt / 2; o.V = V[k] + A_ * dt; o.A = A_;
        interp_by_s(o.s, o);
        return o;
    }
    // traj[s] (trajectories.jl:55-68)
    TrajectoryNode at_s(double sq) const {
        int nn = n();
        int i = ssfirst(s, sq) + 1 - 1;
        if (i < 1) i = 1; if (i > nn - 1) i = nn - 1;
        int k = i - 1;
        double A_ = (V[k + 1] - V[k]) / (t[k + 1] - t[k]);
        double ds = sq - s[k];
        double dt;
        if (std::fabs(A_) < 1e-3 || sq > s[nn - 1]) dt = ds / V[k];
        else dt = (std::sqrt(2 * A_ * ds + V[k] * V[k]) - V[k]) / A_;
        TrajectoryNode o;
        o.t = t[k] + dt; o.s = sq; o.V = V[k] + A_ * dt; o.A = A_;
        interp_by_s(sq, o);
        return o;
    }
    // distance2 (math.jl:4-9)
    static double distance2(double ax, double ay, double bx, double by, double x, double y) {
        double vx = bx - ax, vy = by - ay;
        double lam = (vx * (x - ax) + vy * (y - ay)) / (vx * vx + vy * vy);
        lam = lam > 1 ? 1 : (lam < 0 ? 0 : lam);     // clamp(λ, 0, 1)
        double px = (1 - lam) * ax + lam * bx, py = (1 - lam) * ay + lam * by;
        return (px - x) * (px - x) + (py - y) * (py - y);
    }
    // path_coordinates (trajectories.jl:71-93). Deviation: sqrt argument floored at 0 (the reference would throw a
    // DomainError on a negative round-off).
    void path_coordinates(double x, double y, double& s_out, double& e_out, double& t_out) const {
        double d2min = INFINITY; int imin = 0;
        for (int i = 0; i < n() - 1; i++) {
            double d2 = distance2(E[i], N[i], E[i + 1], N[i + 1], x, y);
            if (d2 < d2min) { d2min = d2; imin = i; }
        }
        int i = imin;
        double vx = E[i + 1] - E[i], vy = N[i + 1] - N[i];
        double wx = x - E[i], wy = y - N[i];
        double arg = wx * wx + wy * wy - d2min;
        double ds = std::sqrt(arg > 0 ? arg : 0.0);
        s_out = s[i] + ds;
        e_out = std::sqrt(d2min) * signd(vx * wy - vy * wx);
        double A_ = (V[i + 1] - V[i]) / (t[i + 1] - t[i]);
        double dt;
        if (std::fabs(A_) < 1e-3) dt = ds / V[i];
        else dt = (std::sqrt(2 * A_ * ds + V[i] * V[i]) - V[i]) / A_;
        t_out = t[i] + dt;
    }
};

// invcumtrapz (math.jl:2) and TrajectoryTube(p::path) (ros_integration.jl:13-16): t = invcumtrapz(Ux_des, s), phi = 0
inline std::vector<double> invcumtrapz(const std::vector<double>& y, const std::vector<double>& x) {
    std::vector<double> r(x.size());
    double acc = 0; r[0] = 0;
    for (size_t i = 1; i < x.size(); i++) { acc += 2 * (x[i] - x[i - 1]) / (y[i - 1] + y[i]); r[i] = acc; }
    return r;
}
// straight_trajectory (trajectories.jl:96-105)
inline TrajectoryTube straight_trajectory(double len, double vel) {
    TrajectoryTube T;
    T.t = {0., len / vel}; T.s = {0., len}; T.V = {vel, vel}; T.A = {0., 0.}; T.E = {0., 0.}; T.N = {0., len};
    T.psi = {0., 0.}; T.kappa = {0., 0.}; T.theta = {0., 0.}; T.phi = {0., 0.}; T.edge_L = {4., 4.}; T.edge_R = {-4., -4.};
    return T;
}

}  // namespace orc

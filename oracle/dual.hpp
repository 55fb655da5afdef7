// oracle/dual.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never linked into the product library.
//
// Forward-mode dual numbers with ForwardDiff.jl 0.10 branch semantics (derivative of the taken branch),
// as used implicitly by the reference through LinearDynamicsModels.linearize and explicitly at
// /root/reference/src/vehicle_dynamics.jl:295 and src/HJI_computation.jl:167.
// PARITY UNPINNED: the reference ships no golden vectors (test/runtests.jl:5 is `@test 1 == 2`).
#pragma once
#include <cmath>

namespace orc {

template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; i++) d[i] = 0; }
    Dual(double x) : v(x) { for (int i = 0; i < N; i++) d[i] = 0; }
    static Dual seed(double x, int k) { Dual r(x); r.d[k] = 1.0; return r; }
};

// value(): strips the tangent (ForwardDiff.value)
inline double value(double x) { return x; }
template <int N> inline double value(const Dual<N>& x) { return x.v; }

template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; i++) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r; double inv = 1.0 / b.v; r.v = a.v * inv;
    for (int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
// mixed with double
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double a, const Dual<N>& b) { return b + a; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double a, const Dual<N>& b) { return (-b) + a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; i++) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double a, const Dual<N>& b) { return b * a; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { return a * (1.0 / b); }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }

// comparisons act on values
template <int N> inline bool operator<(const Dual<N>& a, const Dual<N>& b) { return a.v < b.v; }
template <int N> inline bool operator<(const Dual<N>& a, double b) { return a.v < b; }
template <int N> inline bool operator<(double a, const Dual<N>& b) { return a < b.v; }
template <int N> inline bool operator>(const Dual<N>& a, const Dual<N>& b) { return a.v > b.v; }
template <int N> inline bool operator>(const Dual<N>& a, double b) { return a.v > b; }
template <int N> inline bool operator>(double a, const Dual<N>& b) { return a > b.v; }
template <int N> inline bool operator<=(const Dual<N>& a, double b) { return a.v <= b; }
template <int N> inline bool operator>=(const Dual<N>& a, double b) { return a.v >= b; }
template <int N> inline bool operator>=(const Dual<N>& a, const Dual<N>& b) { return a.v >= b.v; }
template <int N> inline bool operator<=(const Dual<N>& a, const Dual<N>& b) { return a.v <= b.v; }

// elementary functions
using std::sin; using std::cos; using std::tan; using std::atan; using std::atan2; using std::sqrt; using std::fabs;
template <int N> inline Dual<N> sin(const Dual<N>& a) { Dual<N> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < N; i++) r.d[i] = c * a.d[i]; return r; }
template <int N> inline Dual<N> cos(const Dual<N>& a) { Dual<N> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < N; i++) r.d[i] = s * a.d[i]; return r; }
template <int N> inline Dual<N> tan(const Dual<N>& a) { Dual<N> r; r.v = std::tan(a.v); double g = 1.0 + r.v * r.v; for (int i = 0; i < N; i++) r.d[i] = g * a.d[i]; return r; }
template <int N> inline Dual<N> atan(const Dual<N>& a) { Dual<N> r; r.v = std::atan(a.v); double g = 1.0 / (1.0 + a.v * a.v); for (int i = 0; i < N; i++) r.d[i] = g * a.d[i]; return r; }
template <int N> inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
    Dual<N> r; r.v = std::atan2(y.v, x.v); double h = 1.0 / (x.v * x.v + y.v * y.v);
    for (int i = 0; i < N; i++) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * h;
    return r;
}
template <int N> inline Dual<N> atan2(const Dual<N>& y, double x) { return atan2(y, Dual<N>(x)); }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) { Dual<N> r; r.v = std::sqrt(a.v); double g = 0.5 / r.v; for (int i = 0; i < N; i++) r.d[i] = g * a.d[i]; return r; }
// abs(d) = signbit(value(d)) ? -d : d   (ForwardDiff)
inline double absd(double a) { return std::fabs(a); }
template <int N> inline Dual<N> absd(const Dual<N>& a) { return std::signbit(a.v) ? -a : a; }
// sign() carries no derivative
inline double signd(double a) { return (a > 0) - (a < 0); }
template <int N> inline double signd(const Dual<N>& a) { return (a.v > 0) - (a.v < 0); }

// Base.min(x,y) = ifelse(isless(y,x), y, x); Base.max(x,y) = ifelse(isless(y,x), x, y)   (generic Real fallback, used for Duals)
template <class T> inline T jl_min(const T& x, const T& y) { return (value(y) < value(x)) ? y : x; }
template <class T> inline T jl_max(const T& x, const T& y) { return (value(y) < value(x)) ? x : y; }
// Base.clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
template <class T> inline T jl_clamp(const T& x, double lo, double hi) { return (value(x) > hi) ? T(hi) : ((value(x) < lo) ? T(lo) : x); }

}  // namespace orc

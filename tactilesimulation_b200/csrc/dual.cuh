// Forward-mode scalar: value + ONE directional derivative.
//
// Design note (B200): an environment is owned by a tile of LPE lanes; lane k carries the
// tangent along reduced coordinate k.  The Newton matrix H = dg/dq1 and the adjoint blocks
// dg/dq0, dg/dqdot0 are therefore produced column-per-lane by evaluating the SAME residual
// code on Dual numbers -- no 12x12 maximal-coordinate blocks, no cross-lane reductions, and
// every lane runs the fp64 pipe on identical control flow.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN
#endif

struct Dual {
  double v, d;
  Dual() = default;
  HD Dual(double v_) : v(v_), d(0.0) {}
};

HD Dual mkdual(double v, double d) { Dual r; r.v = v; r.d = d; return r; }

HD Dual operator+(Dual a, Dual b) { return mkdual(a.v + b.v, a.d + b.d); }
HD Dual operator+(Dual a, double b) { return mkdual(a.v + b, a.d); }
HD Dual operator+(double a, Dual b) { return mkdual(a + b.v, b.d); }
HD Dual operator-(Dual a, Dual b) { return mkdual(a.v - b.v, a.d - b.d); }
HD Dual operator-(Dual a, double b) { return mkdual(a.v - b, a.d); }
HD Dual operator-(double a, Dual b) { return mkdual(a - b.v, -b.d); }
HD Dual operator-(Dual a) { return mkdual(-a.v, -a.d); }
HD Dual operator*(Dual a, Dual b) { return mkdual(a.v * b.v, a.v * b.d + a.d * b.v); }
HD Dual operator*(Dual a, double b) { return mkdual(a.v * b, a.d * b); }
HD Dual operator*(double a, Dual b) { return mkdual(a * b.v, a * b.d); }
HD Dual operator/(Dual a, Dual b) {
  double q = a.v / b.v;
  return mkdual(q, (a.d - q * b.d) / b.v);
}
HD Dual operator/(Dual a, double b) { return mkdual(a.v / b, a.d / b); }
HD Dual operator/(double a, Dual b) {
  double q = a / b.v;
  return mkdual(q, -q * b.d / b.v);
}
HD Dual& operator+=(Dual& a, Dual b) { a.v += b.v; a.d += b.d; return a; }
HD Dual& operator-=(Dual& a, Dual b) { a.v -= b.v; a.d -= b.d; return a; }
HD Dual& operator+=(Dual& a, double b) { a.v += b; return a; }

HD Dual dsqrt(Dual a) { double s = sqrt(a.v); return mkdual(s, a.d / (2.0 * s)); }
HD double dsqrt(double a) { return sqrt(a); }
HD Dual dabs(Dual a) { return a.v < 0.0 ? mkdual(-a.v, -a.d) : a; }
HD double dabs(double a) { return fabs(a); }
HD void dsincos(Dual a, Dual& s, Dual& c) {
  double sv = sin(a.v), cv = cos(a.v);
  s = mkdual(sv, cv * a.d);
  c = mkdual(cv, -sv * a.d);
}
HD void dsincos(double a, double& s, double& c) { s = sin(a); c = cos(a); }

HD double val(double a) { return a; }
HD double val(Dual a) { return a.v; }
HD double tan_of(double) { return 0.0; }
HD double tan_of(Dual a) { return a.d; }

template <class T> struct Lift;
template <> struct Lift<double> { static HD double mk(double v, double) { return v; } };
template <> struct Lift<Dual> { static HD Dual mk(double v, double d) { return mkdual(v, d); } };

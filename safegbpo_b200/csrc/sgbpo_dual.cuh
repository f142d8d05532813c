// Forward-mode dual numbers for the per-environment maps of the safeguarded rollout.
//
// Why duals: every per-environment function on the hot path (simulator step, linearisation,
// safe-set construction, projection / ray-mask map, observation, reward) is a tiny piecewise-smooth
// map R^(n+m) -> R^(n+o+1+m) with n <= 6, m <= 2.  The reverse pass needs its full Jacobian
// transpose times a cotangent, INCLUDING the second-order terms that appear because the safeguard
// consumes B = df/da (the reference differentiates through jacrev, safeguard.py:206-207).  Nesting
// Dual<Dual<T,K>,M> gives exactly that with zero hand-derived formulas; everything stays in registers.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif

namespace sgbpo {

template <class S, int K> struct Dual;

template <class S> struct Traits { using base = S; static constexpr int depth = 0; };
template <class S, int K> struct Traits<Dual<S, K>> {
  using base = typename Traits<S>::base;
  static constexpr int depth = Traits<S>::depth + 1;
};

// ---- base scalar functions ------------------------------------------------------------------
SG_HD float sg_sin(float x) { return sinf(x); }
SG_HD double sg_sin(double x) { return sin(x); }
SG_HD float sg_cos(float x) { return cosf(x); }
SG_HD double sg_cos(double x) { return cos(x); }
SG_HD float sg_atan2(float y, float x) { return atan2f(y, x); }
SG_HD double sg_atan2(double y, double x) { return atan2(y, x); }
SG_HD float sg_sqrt(float x) { return sqrtf(x); }
SG_HD double sg_sqrt(double x) { return sqrt(x); }
SG_HD float sg_tanh(float x) { return tanhf(x); }
SG_HD double sg_tanh(double x) { return tanh(x); }
SG_HD float sg_exp(float x) { return expf(x); }
SG_HD double sg_exp(double x) { return exp(x); }
SG_HD float sg_log(float x) { return logf(x); }
SG_HD double sg_log(double x) { return log(x); }
SG_HD float sg_abs(float x) { return fabsf(x); }
SG_HD double sg_abs(double x) { return fabs(x); }
SG_HD float primal(float x) { return x; }
SG_HD double primal(double x) { return x; }
// sine and cosine from ONE range reduction (every simulator needs both of the same angle)
SG_HD void sg_sincos(float x, float& s, float& c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, &s, &c);
#else
  s = sinf(x); c = cosf(x);
#endif
}
SG_HD void sg_sincos(double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(x, &s, &c);
#else
  s = sin(x); c = cos(x);
#endif
}

// ---- Dual -------------------------------------------------------------------------------------
template <class S, int K> struct Dual {
  using base = typename Traits<S>::base;
  S v;
  S d[K];
  SG_HD Dual() {}
  SG_HD Dual(base c) : v(c) {
#pragma unroll
    for (int k = 0; k < K; ++k) d[k] = S(base(0));
  }
  SG_HD static Dual lift(const S& x) {  // constant w.r.t. this level's seeds
    Dual r;
    r.v = x;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = S(base(0));
    return r;
  }
  SG_HD static Dual seed(const S& x, int which) {
    Dual r = lift(x);
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (k == which) r.d[k] = S(base(1));
    return r;
  }

  friend SG_HD Dual operator+(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v + b.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] + b.d[k];
    return r;
  }
  friend SG_HD Dual operator-(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v - b.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] - b.d[k];
    return r;
  }
  friend SG_HD Dual operator-(const Dual& a) {
    Dual r;
    r.v = -a.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = -a.d[k];
    return r;
  }
  friend SG_HD Dual operator*(const Dual& a, const Dual& b) {
    Dual r;
    r.v = a.v * b.v;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
  }
  friend SG_HD Dual operator/(const Dual& a, const Dual& b) {
    Dual r;
    S inv = S(base(1)) / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int k = 0; k < K; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    return r;
  }
  SG_HD Dual& operator+=(const Dual& b) { return *this = *this + b; }
  SG_HD Dual& operator-=(const Dual& b) { return *this = *this - b; }
  SG_HD Dual& operator*=(const Dual& b) { return *this = *this * b; }
};

template <class S, int K> SG_HD typename Traits<S>::base primal(const Dual<S, K>& x) { return primal(x.v); }

// same dual structure over double: the iterative program P2 (m = 2) is solved in double whatever the I/O type
template <class X> struct Widen { using type = double; };
template <class S, int K> struct Widen<Dual<S, K>> { using type = Dual<typename Widen<S>::type, K>; };
SG_HD double widen(float x) { return (double)x; }
SG_HD double widen(double x) { return x; }
template <class S, int K> SG_HD typename Widen<Dual<S, K>>::type widen(const Dual<S, K>& x) {
  typename Widen<Dual<S, K>>::type r;
  r.v = widen(x.v);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = widen(x.d[k]);
  return r;
}
SG_HD void narrow(double x, float& o) { o = (float)x; }
SG_HD void narrow(double x, double& o) { o = x; }
template <class SW, class S, int K> SG_HD void narrow(const Dual<SW, K>& x, Dual<S, K>& o) {
  narrow(x.v, o.v);
#pragma unroll
  for (int k = 0; k < K; ++k) narrow(x.d[k], o.d[k]);
}

template <class S, int K> SG_HD Dual<S, K> chain(const Dual<S, K>& x, const S& fv, const S& dfdx) {
  Dual<S, K> r;
  r.v = fv;
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = dfdx * x.d[k];
  return r;
}
template <class S, int K> SG_HD Dual<S, K> sg_sin(const Dual<S, K>& x) { return chain(x, sg_sin(x.v), sg_cos(x.v)); }
template <class S, int K> SG_HD Dual<S, K> sg_cos(const Dual<S, K>& x) { return chain(x, sg_cos(x.v), -sg_sin(x.v)); }
template <class S, int K> SG_HD void sg_sincos(const Dual<S, K>& x, Dual<S, K>& s, Dual<S, K>& c) {
  S sv, cv;
  sg_sincos(x.v, sv, cv);
  s = chain(x, sv, cv);
  c = chain(x, cv, -sv);
}
template <class S, int K> SG_HD Dual<S, K> sg_sqrt(const Dual<S, K>& x) {
  using B = typename Traits<S>::base;
  S r = sg_sqrt(x.v);
  return chain(x, r, S(B(0.5)) / r);
}
template <class S, int K> SG_HD Dual<S, K> sg_tanh(const Dual<S, K>& x) {
  using B = typename Traits<S>::base;
  S t = sg_tanh(x.v);
  return chain(x, t, S(B(1)) - t * t);
}
template <class S, int K> SG_HD Dual<S, K> sg_exp(const Dual<S, K>& x) {
  S e = sg_exp(x.v);
  return chain(x, e, e);
}
template <class S, int K> SG_HD Dual<S, K> sg_log(const Dual<S, K>& x) {
  using B = typename Traits<S>::base;
  return chain(x, sg_log(x.v), S(B(1)) / x.v);
}
template <class S, int K> SG_HD Dual<S, K> sg_abs(const Dual<S, K>& x) {
  using B = typename Traits<S>::base;
  B p = primal(x.v);
  B sgn = p > B(0) ? B(1) : (p < B(0) ? B(-1) : B(0));   // torch: d|x|/dx = sign(x)
  return chain(x, sg_abs(x.v), S(sgn));
}
template <class S, int K> SG_HD Dual<S, K> sg_atan2(const Dual<S, K>& y, const Dual<S, K>& x) {
  Dual<S, K> r;
  r.v = sg_atan2(y.v, x.v);
  S inv = S(typename Traits<S>::base(1)) / (x.v * x.v + y.v * y.v);
  S cy = x.v * inv, cx = -(y.v * inv);
#pragma unroll
  for (int k = 0; k < K; ++k) r.d[k] = cy * y.d[k] + cx * x.d[k];
  return r;
}

// ---- generic helpers (work for float/double/Dual) -----------------------------------------------
template <class S> SG_HD S C(double c) { return S(typename Traits<S>::base(c)); }
template <class S> SG_HD S sel(bool c, const S& a, const S& b) { return c ? a : b; }
template <class S> SG_HD S sg_min(const S& a, const S& b) { return primal(a) <= primal(b) ? a : b; }
template <class S> SG_HD S sg_max(const S& a, const S& b) { return primal(a) >= primal(b) ? a : b; }
// torch.clamp: gradient passes where lo <= x <= hi (inclusive), result is the (constant) bound otherwise
template <class S> SG_HD S sg_clamp_const(const S& x, double lo, double hi) {
  using B = typename Traits<S>::base;
  B p = primal(x);
  if (p < B(lo)) return C<S>(lo);
  if (p > B(hi)) return C<S>(hi);
  return x;
}
// The simulators wrap angles with atan2(sin x, cos x) (pendulum.py / cartpole.py / quadrotor.py): the representative of x in
// (-pi, pi], derivative 1.  Evaluated as x - 2 pi rint(x / 2 pi) with a two-term 2 pi (exact for |x| <= pi, within one rounding of
// the atan2 form elsewhere): the atan2(sin, cos) form costs two range reductions and an atan2 per dynamics evaluation, 180 of the
// ~1 100 instructions of a cartpole step.
SG_HD float wrap_angle(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  return fmaf(-k, -1.7484555e-7f, fmaf(-k, 6.2831854820251465f, x));
}
SG_HD double wrap_angle(double x) {
  const double k = rint(x * 0.15915494309189535);
  return fma(-k, 2.4492935982947064e-16, fma(-k, 6.283185307179586, x));
}
template <class S, int K> SG_HD Dual<S, K> wrap_angle(const Dual<S, K>& x) {
  Dual<S, K> r = x;
  r.v = wrap_angle(x.v);
  return r;
}

}  // namespace sgbpo

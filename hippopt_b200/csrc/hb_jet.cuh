// hb_jet.cuh -- truncated Taylor series ("jets") used by the smooth-terrain contact block.
//
//   BJ<N> : bivariate (x, y), total order N, NORMALISED coefficients
//           c[bidx(i, j)] = d^{i+j} f / (dx^i dy^j) / (i! j!).     N = 4 for the terrain surface T(x, y),
//           N = 3 for the unit normal, N = 2 for everything the constraints use.
//   TJ    : trivariate (x, y, z), order 2, normalised; order: 1, x, y, z, xx, xy, xz, yy, yz, zz.
//
// The reference gets the same derivatives from CasADi's AD of `smooth_terrain.py:201-227` and
// `terrain_descriptor.py:45-80`; the DCC margin (`complementarity.py:68-89`) contains the Jacobian of
// the normal, so its Hessian needs fourth derivatives of the surface -- hence order 4.
#pragma once
#include "hb_math.cuh"

namespace hb {

__host__ __device__ constexpr int bidx(int i, int j) { return (i + j) * (i + j + 1) / 2 + j; }

template <int N>
struct BJ {
  static constexpr int SIZE = (N + 1) * (N + 2) / 2;
  double c[SIZE];
};

template <int N>
__device__ __forceinline__ BJ<N> bj_const(double v) {
  BJ<N> r;
#pragma unroll
  for (int i = 0; i < BJ<N>::SIZE; ++i) r.c[i] = 0.0;
  r.c[0] = v;
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator+(const BJ<N>& a, const BJ<N>& b) {
  BJ<N> r;
#pragma unroll
  for (int i = 0; i < BJ<N>::SIZE; ++i) r.c[i] = a.c[i] + b.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator-(const BJ<N>& a, const BJ<N>& b) {
  BJ<N> r;
#pragma unroll
  for (int i = 0; i < BJ<N>::SIZE; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator-(const BJ<N>& a) {
  BJ<N> r;
#pragma unroll
  for (int i = 0; i < BJ<N>::SIZE; ++i) r.c[i] = -a.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator*(double s, const BJ<N>& a) {
  BJ<N> r;
#pragma unroll
  for (int i = 0; i < BJ<N>::SIZE; ++i) r.c[i] = s * a.c[i];
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator+(const BJ<N>& a, double s) {
  BJ<N> r = a;
  r.c[0] += s;
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> operator*(const BJ<N>& a, const BJ<N>& b) {
  BJ<N> r = bj_const<N>(0.0);
#pragma unroll
  for (int i1 = 0; i1 <= N; ++i1)
#pragma unroll
    for (int j1 = 0; j1 <= N; ++j1) {
      if (i1 + j1 > N) continue;
#pragma unroll
      for (int i2 = 0; i2 <= N; ++i2)
#pragma unroll
        for (int j2 = 0; j2 <= N; ++j2) {
          if (i1 + j1 + i2 + j2 > N) continue;
          r.c[bidx(i1 + i2, j1 + j2)] += a.c[bidx(i1, j1)] * b.c[bidx(i2, j2)];
        }
    }
  return r;
}
// f(a) with f[k] = f^(k)(a_0) / k!, k = 0..N (Horner in delta = a - a_0)
template <int N>
__device__ __forceinline__ BJ<N> bj_compose(const BJ<N>& a, const double* f) {
  BJ<N> d = a;
  d.c[0] = 0.0;
  BJ<N> r = bj_const<N>(f[N]);
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    r = r * d;
    r.c[0] += f[k];
  }
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N> bj_invsqrt(const BJ<N>& a) {
  const double s = a.c[0];
  const double f0 = 1.0 / sqrt(s), is = 1.0 / s;
  double f[5];
  f[0] = f0;
  f[1] = -0.5 * f[0] * is;
  f[2] = -0.75 * f[1] * is;          // (3/8) s^-5/2
  f[3] = -(5.0 / 6.0) * f[2] * is;   // -(5/16) s^-7/2
  f[4] = -0.875 * f[3] * is;         // (35/128) s^-9/2
  return bj_compose<N>(a, f);
}
template <int N, int M>
__device__ __forceinline__ BJ<M> bj_trunc(const BJ<N>& a) {
  static_assert(M <= N, "truncate only");
  BJ<M> r;
#pragma unroll
  for (int i = 0; i < BJ<M>::SIZE; ++i) r.c[i] = a.c[i];
  return r;
}
// partial derivatives lower the order by one
template <int N>
__device__ __forceinline__ BJ<N - 1> bj_ddx(const BJ<N>& a) {
  BJ<N - 1> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (i + j > N - 1) continue;
      r.c[bidx(i, j)] = (i + 1) * a.c[bidx(i + 1, j)];
    }
  return r;
}
template <int N>
__device__ __forceinline__ BJ<N - 1> bj_ddy(const BJ<N>& a) {
  BJ<N - 1> r;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (i + j > N - 1) continue;
      r.c[bidx(i, j)] = (j + 1) * a.c[bidx(i, j + 1)];
    }
  return r;
}

// ------------------------------------------------------------------------------------------------
struct TJ {
  double c[10];  // 1, x, y, z, xx, xy, xz, yy, yz, zz
};
__device__ __forceinline__ TJ tj_const(double v) {
  TJ r;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.c[i] = 0.0;
  r.c[0] = v;
  return r;
}
__device__ __forceinline__ TJ tj_from(const BJ<2>& a) {
  TJ r;
  r.c[0] = a.c[bidx(0, 0)];
  r.c[1] = a.c[bidx(1, 0)];
  r.c[2] = a.c[bidx(0, 1)];
  r.c[3] = 0.0;
  r.c[4] = a.c[bidx(2, 0)];
  r.c[5] = a.c[bidx(1, 1)];
  r.c[6] = 0.0;
  r.c[7] = a.c[bidx(0, 2)];
  r.c[8] = 0.0;
  r.c[9] = 0.0;
  return r;
}
__device__ __forceinline__ TJ operator+(const TJ& a, const TJ& b) {
  TJ r;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.c[i] = a.c[i] + b.c[i];
  return r;
}
__device__ __forceinline__ TJ operator-(const TJ& a, const TJ& b) {
  TJ r;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}
__device__ __forceinline__ TJ operator-(const TJ& a) {
  TJ r;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.c[i] = -a.c[i];
  return r;
}
__device__ __forceinline__ TJ operator*(double s, const TJ& a) {
  TJ r;
#pragma unroll
  for (int i = 0; i < 10; ++i) r.c[i] = s * a.c[i];
  return r;
}
__device__ __forceinline__ TJ operator+(const TJ& a, double s) {
  TJ r = a;
  r.c[0] += s;
  return r;
}
__device__ __forceinline__ TJ operator*(const TJ& a, const TJ& b) {
  TJ r;
  r.c[0] = a.c[0] * b.c[0];
  r.c[1] = a.c[0] * b.c[1] + a.c[1] * b.c[0];
  r.c[2] = a.c[0] * b.c[2] + a.c[2] * b.c[0];
  r.c[3] = a.c[0] * b.c[3] + a.c[3] * b.c[0];
  r.c[4] = a.c[0] * b.c[4] + a.c[4] * b.c[0] + a.c[1] * b.c[1];
  r.c[5] = a.c[0] * b.c[5] + a.c[5] * b.c[0] + a.c[1] * b.c[2] + a.c[2] * b.c[1];
  r.c[6] = a.c[0] * b.c[6] + a.c[6] * b.c[0] + a.c[1] * b.c[3] + a.c[3] * b.c[1];
  r.c[7] = a.c[0] * b.c[7] + a.c[7] * b.c[0] + a.c[2] * b.c[2];
  r.c[8] = a.c[0] * b.c[8] + a.c[8] * b.c[0] + a.c[2] * b.c[3] + a.c[3] * b.c[2];
  r.c[9] = a.c[0] * b.c[9] + a.c[9] * b.c[0] + a.c[3] * b.c[3];
  return r;
}
// f(a) with f0 = f(a_0), f1 = f'(a_0), f2 = f''(a_0) / 2
__device__ __forceinline__ TJ tj_compose(const TJ& a, double f0, double f1, double f2) {
  TJ r;
  r.c[0] = f0;
  r.c[1] = f1 * a.c[1];
  r.c[2] = f1 * a.c[2];
  r.c[3] = f1 * a.c[3];
  r.c[4] = f1 * a.c[4] + f2 * a.c[1] * a.c[1];
  r.c[5] = f1 * a.c[5] + 2.0 * f2 * a.c[1] * a.c[2];
  r.c[6] = f1 * a.c[6] + 2.0 * f2 * a.c[1] * a.c[3];
  r.c[7] = f1 * a.c[7] + f2 * a.c[2] * a.c[2];
  r.c[8] = f1 * a.c[8] + 2.0 * f2 * a.c[2] * a.c[3];
  r.c[9] = f1 * a.c[9] + f2 * a.c[3] * a.c[3];
  return r;
}
// second derivatives d2/dp_a dp_b, a <= b, in the order xx, xy, xz, yy, yz, zz
__device__ __forceinline__ double tj_hess(const TJ& a, int e) {
  return (e == 0 || e == 3 || e == 5) ? 2.0 * a.c[4 + e] : a.c[4 + e];
}

}  // namespace hb

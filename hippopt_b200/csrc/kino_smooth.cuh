// kino_smooth.cuh -- smooth-step terrain (BASELINE config 5) for the contact kernel.
//
// Restates, with runtime parameters instead of baked constants,
//   SmoothTerrain.create_height_function   utilities/smooth_terrain.py:201-227
//   SmoothTerrain.step                     utilities/smooth_terrain.py:266-336   (edge 5, side 10, yaw 0)
//   TerrainSum.create_height_function      utilities/terrain_sum.py:19-38
//   TerrainDescriptor normal / orientation utilities/terrain_descriptor.py:45-80
// and the terrain-dependent rows/costs of the contact block (complementarity.py:24-32, 68-89;
// contacts.py:24, 54-66, 158-166) including their first and second derivatives, obtained from
// truncated Taylor series of the surface (hb_jet.cuh) instead of CasADi's graph AD.
//
//   h(p) = z - sum_i exp(-g_i^20) height_i,   g_i = (2 (x - ox_i) / l_i)^10 + (2 (y - oy_i) / w_i)^10
// terrain parameters tp[10] = (l, w, height, ox, oy) x 2.
#pragma once
#include "hb_jet.cuh"

namespace hb {

__device__ __forceinline__ BJ<4> smooth_steps_surface(const double* __restrict__ tp, double x, double y) {
  BJ<4> T = bj_const<4>(0.0);
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const double l = tp[5 * s], w = tp[5 * s + 1], height = tp[5 * s + 2], ox = tp[5 * s + 3], oy = tp[5 * s + 4];
    const double dX = 2.0 / l, dY = 2.0 / w;
    const double X0 = dX * (x - ox), Y0 = dY * (y - oy);
    // X^10 and Y^10 expanded to fourth order: C(10,k) X0^(10-k) dX^k
    const double X2 = X0 * X0, X4 = X2 * X2, X6 = X4 * X2, Y2 = Y0 * Y0, Y4 = Y2 * Y2, Y6 = Y4 * Y2;
    BJ<4> g = bj_const<4>(X6 * X4 + Y6 * Y4);
    g.c[bidx(1, 0)] = 10.0 * X6 * X2 * X0 * dX;
    g.c[bidx(2, 0)] = 45.0 * X6 * X2 * dX * dX;
    g.c[bidx(3, 0)] = 120.0 * X6 * X0 * dX * dX * dX;
    g.c[bidx(4, 0)] = 210.0 * X6 * dX * dX * dX * dX;
    g.c[bidx(0, 1)] = 10.0 * Y6 * Y2 * Y0 * dY;
    g.c[bidx(0, 2)] = 45.0 * Y6 * Y2 * dY * dY;
    g.c[bidx(0, 3)] = 120.0 * Y6 * Y0 * dY * dY * dY;
    g.c[bidx(0, 4)] = 210.0 * Y6 * dY * dY * dY * dY;
    // G = g^20: C(20,k) g0^(20-k)
    const double g0 = g.c[0];
    // Far outside the step exp(-g^20) underflows to exactly 0 while the powers of g overflow; the
    // reference's graph then evaluates 0 * inf = NaN in the high derivatives.  The limit is 0, so the
    // step is skipped there (identical wherever the reference is finite, finite where it is not).
    if (g0 > 1.42) continue;  // g^20 > 1100 > -log(DBL_TRUE_MIN): exp(-G) == 0 in fp64
    const double g2 = g0 * g0, g4 = g2 * g2, g8 = g4 * g4, g16 = g8 * g8;
    double fp[5];
    fp[4] = 4845.0 * g16;
    fp[3] = 1140.0 * g16 * g0;
    fp[2] = 190.0 * g16 * g2;
    fp[1] = 20.0 * g16 * g2 * g0;
    fp[0] = g16 * g4;
    const BJ<4> G = bj_compose<4>(g, fp);
    // E = exp(-G)
    const double e0 = exp(-G.c[0]);
    const double fe[5] = {e0, e0, e0 / 2.0, e0 / 6.0, e0 / 24.0};
    const BJ<4> E = bj_compose<4>(-G, fe);
    T = T + height * E;
  }
  return T;
}

struct TFrame {
  TJ h, gx, gy;      // height and the x, y components of its gradient (the z component is 1)
  TJ n[3], xh[3], yh[3];
  TJ Dn[3][2];       // d n_a / d x, d n_a / d y  (d / d z = 0)
};

__device__ __forceinline__ void smooth_terrain_frame(const double* __restrict__ tp, double x, double y, double z,
                                                     TFrame& F) {
  const BJ<4> T = smooth_steps_surface(tp, x, y);
  const BJ<3> gx3 = -bj_ddx<4>(T), gy3 = -bj_ddy<4>(T);
  const BJ<3> inv3 = bj_invsqrt<3>(gx3 * gx3 + gy3 * gy3 + 1.0);
  BJ<3> n3[3];
  n3[0] = gx3 * inv3;
  n3[1] = gy3 * inv3;
  n3[2] = inv3;
  BJ<2> n2[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    F.Dn[a][0] = tj_from(bj_ddx<3>(n3[a]));
    F.Dn[a][1] = tj_from(bj_ddy<3>(n3[a]));
    n2[a] = bj_trunc<3, 2>(n3[a]);
    F.n[a] = tj_from(n2[a]);
  }
  // y0 = n x e_x, x0 = y0 x n, x = x0 / |x0|, y = n x x   (terrain_descriptor.py:66-71)
  BJ<2> x0[3];
  x0[0] = n2[2] * n2[2] + n2[1] * n2[1];
  x0[1] = -(n2[0] * n2[1]);
  x0[2] = -(n2[0] * n2[2]);
  const BJ<2> xi = bj_invsqrt<2>(x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2]);
  BJ<2> xh[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    xh[a] = x0[a] * xi;
    F.xh[a] = tj_from(xh[a]);
  }
  F.yh[0] = tj_from(n2[1] * xh[2] - n2[2] * xh[1]);
  F.yh[1] = tj_from(n2[2] * xh[0] - n2[0] * xh[2]);
  F.yh[2] = tj_from(n2[0] * xh[1] - n2[1] * xh[0]);
  F.h = tj_from(-bj_trunc<4, 2>(T));
  F.h.c[0] += z;
  F.h.c[3] = 1.0;
  F.gx = tj_from(bj_trunc<3, 2>(gx3));
  F.gy = tj_from(bj_trunc<3, 2>(gy3));
}

__device__ __forceinline__ TJ tj_dot(const TJ* a, D3 b) { return b.x * a[0] + b.y * a[1] + b.z * a[2]; }
__device__ __forceinline__ double comp(D3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// Values and derivatives of every terrain-dependent row / cost of one contact point.
struct SmoothPoint {
  // rows (Taylor series in the point position, all other variables frozen)
  TJ planar[3], dcc, fric, swing, N;
  TFrame F;
  TJ tau;
};

}  // namespace hb

// interp.cu -- initial-guess generation on the device.
//
// Replaces, for a batch of instances, `humanoid_state_interpolator`
// (/root/reference/src/hippopt/robot_planning/utilities/interpolators.py:396-448) and its callees:
//   linear_interpolator (:24-50), quaternion_slerp (:53-77), transform_interpolator (:80-103),
//   the per-point part of foot_contact_state_interpolator (:106-309: append_stance_phase :171-181,
//   append_swing_phase :183-229), FootContactState.set_from_parent_frame_transform (variables/contacts.py:103-107).
// The CONTROL FLOW of foot_contact_state_interpolator (which phase a point belongs to, how many points a
// swing half gets) depends on times only, not on the instance: the host (hippopt_b200/interpolators.py)
// turns it into a per-point schedule once, the kernel evaluates the transforms for every instance.
//
// One CTA per (point k, instance b); thread roles: 0..7 contact points, 8 base pose, 32.. joints and CoM.
// Writes 82 + n_joints doubles per (b, k) -- a pure HBM-write kernel.
#include <cstdint>

namespace hb {

// schedule entry (5 x int32 per foot and point)
enum { IS_KIND = 0, IS_A = 1, IS_B = 2, IS_J = 3, IS_N = 4, IS_STRIDE = 5 };
enum { IK_STANCE = 0, IK_SWING_UP = 1, IK_SWING_DOWN = 2 };
// phase record (17 doubles): transform position, quaternion (xyzw), mid-swing position, quaternion, force
enum { IP_POS = 0, IP_QUAT = 3, IP_MPOS = 7, IP_MQUAT = 10, IP_FORCE = 14, IP_STRIDE = 17 };
// state block (hippopt_b200/kino_layout.py ParamOffsets.st_pt / ST_*): 8 x (p, f, descriptor), base, joints, com
enum { IST_PB = 72, IST_Q = 75, IST_S = 79 };
// knot variables z (kino_layout.py:36-37)
enum { IZ_P = 6, IZ_F = 9, IZ_PB = 127, IZ_Q = 130, IZ_S = 157, IZ_COM = 180, IZ_N = 189 };

// (1 - t) a + t b with the reference's rounding (two products, one sum: no FMA contraction)
__device__ __forceinline__ double lerp(double a, double b, double t) {
  return __dadd_rn(__dmul_rn(1.0 - t, a), __dmul_rn(t, b));
}

// np.linspace(0, 1, n)[j] (interpolators.py:45): arange * step, last sample exactly 1
__device__ __forceinline__ double linspace01(int j, int n) {
  if (n <= 1 || j == 0) return 0.0;
  if (j == n - 1) return 1.0;
  return (double)j * (1.0 / (double)(n - 1));
}

// quaternion_slerp (interpolators.py:53-77) with liecasadi's Quaternion.slerp_step [ext]:
// (sin((1-t) a) q0 + sin(t a) q1) / sin(a), a = acos(q0 . q1); q0 when |a| <= 1e-6 (or a is NaN)
__device__ __forceinline__ void slerp(const double* q0, const double* q1, double t, double* out) {
  const double dot = q0[0] * q1[0] + q0[1] * q1[1] + q0[2] * q1[2] + q0[3] * q1[3];
  const double ang = acos(dot);
  if (fabs(ang) > 1e-6) {
    const double s0 = sin((1.0 - t) * ang), s1 = sin(t * ang), sa = sin(ang);
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = (s0 * q0[c] + s1 * q1[c]) / sa;
  } else {
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = q0[c];
  }
}

__global__ void __launch_bounds__(128)
interp_states_kernel(int n_points, int n_joints, const double* __restrict__ initial,
                     const double* __restrict__ final_, const int32_t* __restrict__ schedule,
                     const double* __restrict__ ph_l, long stride_l, const double* __restrict__ ph_r, long stride_r,
                     double* __restrict__ states, double* __restrict__ x, long x_stride, int knot0) {
  const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int ns = IST_S + n_joints + 3;
  const double* s0 = initial + (long)b * ns;
  const double* s1 = final_ + (long)b * ns;
  double* so = states ? states + ((long)b * n_points + k) * ns : nullptr;
  double* z = x ? x + (long)b * x_stride + (long)(knot0 + k) * IZ_N : nullptr;
  const double t = linspace01(k, n_points);

  if (tid < 8) {
    const int foot = tid >> 2;
    const int32_t* e = schedule + ((long)foot * n_points + k) * IS_STRIDE;
    const double* ph = foot ? ph_r + (long)b * stride_r : ph_l + (long)b * stride_l;
    const int kind = e[IS_KIND];
    const double* A = ph + (long)e[IS_A] * IP_STRIDE;
    const double* B = ph + (long)e[IS_B] * IP_STRIDE;
    double pos[3], q[4], f[3] = {0.0, 0.0, 0.0};
    if (kind == IK_STANCE) {
#pragma unroll
      for (int c = 0; c < 3; ++c) pos[c] = A[IP_POS + c], f[c] = A[IP_FORCE + c];
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = A[IP_QUAT + c];
    } else {
      // transform_interpolator (:80-103) over the half swing: translation linear, rotation slerp
      const double* p0 = kind == IK_SWING_UP ? A + IP_POS : A + IP_MPOS;
      const double* q0 = kind == IK_SWING_UP ? A + IP_QUAT : A + IP_MQUAT;
      const double* p1 = kind == IK_SWING_UP ? A + IP_MPOS : B + IP_POS;
      const double* q1 = kind == IK_SWING_UP ? A + IP_MQUAT : B + IP_QUAT;
      const double u = linspace01(e[IS_J], e[IS_N]);
#pragma unroll
      for (int c = 0; c < 3; ++c) pos[c] = lerp(p0[c], p1[c], u);
      slerp(q0, q1, u, q);
    }
    // p = translation + R(q) d, R = I + 2 w [v]x + 2 [v]x^2
    const double* d = s0 + 9 * tid + 6;
    const double d0 = d[0], d1 = d[1], d2 = d[2];
    const double c0 = q[1] * d2 - q[2] * d1, c1 = q[2] * d0 - q[0] * d2, c2 = q[0] * d1 - q[1] * d0;
    const double e0 = q[1] * c2 - q[2] * c1, e1 = q[2] * c0 - q[0] * c2, e2 = q[0] * c1 - q[1] * c0;
    const double p[3] = {pos[0] + d0 + 2.0 * q[3] * c0 + 2.0 * e0, pos[1] + d1 + 2.0 * q[3] * c1 + 2.0 * e1,
                         pos[2] + d2 + 2.0 * q[3] * c2 + 2.0 * e2};
    if (so) {
      double* o = so + 9 * tid;
      o[0] = p[0], o[1] = p[1], o[2] = p[2], o[3] = f[0], o[4] = f[1], o[5] = f[2], o[6] = d0, o[7] = d1, o[8] = d2;
    }
    if (z) {
      double* o = z + 15 * tid;
      o[IZ_P] = p[0], o[IZ_P + 1] = p[1], o[IZ_P + 2] = p[2], o[IZ_F] = f[0], o[IZ_F + 1] = f[1], o[IZ_F + 2] = f[2];
    }
  } else if (tid == 8) {
    // free_floating_object_state_interpolator (:340-366)
    double q[4], pb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) pb[c] = lerp(s0[IST_PB + c], s1[IST_PB + c], t);
    slerp(s0 + IST_Q, s1 + IST_Q, t, q);
    if (so) {
      so[IST_PB] = pb[0], so[IST_PB + 1] = pb[1], so[IST_PB + 2] = pb[2];
      so[IST_Q] = q[0], so[IST_Q + 1] = q[1], so[IST_Q + 2] = q[2], so[IST_Q + 3] = q[3];
    }
    if (z) {
      z[IZ_PB] = pb[0], z[IZ_PB + 1] = pb[1], z[IZ_PB + 2] = pb[2];
      z[IZ_Q] = q[0], z[IZ_Q + 1] = q[1], z[IZ_Q + 2] = q[2], z[IZ_Q + 3] = q[3];
    }
  } else if (tid >= 32) {
    // kinematic_tree_state_interpolator (:369-393) and the CoM (:419-423): contiguous in the state block
    for (int i = tid - 32; i < n_joints + 3; i += 96) {
      const double v = lerp(s0[IST_S + i], s1[IST_S + i], t);
      if (so) so[IST_S + i] = v;
      if (z) z[(i < n_joints ? IZ_S : IZ_COM - n_joints) + i] = v;
    }
  }
}

}  // namespace hb

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
// Work item = one (instance, point).  A warp takes IW_ITEMS = 3 consecutive items: lane = 9 * item + role,
// role 0..7 = contact point, role 8 = base pose, so that the one expensive path (slerp: acos, three sin,
// four divisions in fp64) runs once per warp with 27 lanes busy; the joints and the CoM are spread over all
// lanes afterwards.  Results are staged in shared memory and leave as coalesced 8-byte stores: the state
// blocks of the three items are contiguous in `states`, the decision-vector scatter writes runs of 6 / 7 / 26.
// (The first version -- one CTA per item, a thread per role, direct stores -- ran at 0.75 TB/s and was bound
// by the slerp instructions of its one busy warp: profiles/r01/ncu_interp_v10.txt.)
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
enum { IZ_P = 6, IZ_PB = 127, IZ_S = 157, IZ_N = 189, IZ_WRITTEN = 81 };
enum { IW_ITEMS = 3, IW_WARPS = 4 };

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

// 8 CTAs per SM (64 registers, 20 B of spills): the warps wait on their global loads (35 % long-scoreboard stalls at
// 6 CTAs per SM), 70 -> 61 us for 4 096 x 30 items
__global__ void __launch_bounds__(32 * IW_WARPS, 8)
interp_states_kernel(long total, int n_points, int n_joints, const double* __restrict__ initial,
                     const double* __restrict__ final_, const int32_t* __restrict__ schedule,
                     const double* __restrict__ ph_l, long stride_l, const double* __restrict__ ph_r, long stride_r,
                     double* __restrict__ states, double* __restrict__ x, long x_stride, int knot0) {
  extern __shared__ double ismem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ns = IST_S + n_joints + 3, nl = n_joints + 3;
  double* blk = ismem + warp * IW_ITEMS * ns;
  const long item0 = ((long)blockIdx.x * IW_WARPS + warp) * IW_ITEMS;
  if (item0 >= total) return;
  const int n_it = total - item0 < IW_ITEMS ? (int)(total - item0) : IW_ITEMS;

  // (instance, point) of the three items from ONE 64-bit division; the loops below select by item index
  long bs[IW_ITEMS];
  int ks[IW_ITEMS];
  bs[0] = item0 / n_points, ks[0] = (int)(item0 - bs[0] * n_points);
#pragma unroll
  for (int s = 1; s < IW_ITEMS; ++s) {
    const bool wrap = ks[s - 1] + 1 == n_points;
    bs[s] = bs[s - 1] + (wrap ? 1 : 0), ks[s] = wrap ? 0 : ks[s - 1] + 1;
  }
  const double step = n_points > 1 ? 1.0 / (double)(n_points - 1) : 0.0;  // linspace01 without its division
  double ts[IW_ITEMS];
#pragma unroll
  for (int s = 0; s < IW_ITEMS; ++s) ts[s] = ks[s] == 0 ? 0.0 : (ks[s] == n_points - 1 ? 1.0 : (double)ks[s] * step);

  const int sub = lane / 9, role = lane - 9 * sub;
  if (sub < n_it) {
    const long b = sub == 0 ? bs[0] : (sub == 1 ? bs[1] : bs[2]);
    const int k = sub == 0 ? ks[0] : (sub == 1 ? ks[1] : ks[2]);
    const double* s0 = initial + b * ns;
    const double* s1 = final_ + b * ns;
    const double *p0, *p1, *q0, *q1, *fz = nullptr;
    double u;
    if (role == 8) {
      // free_floating_object_state_interpolator (:340-366)
      p0 = s0 + IST_PB, p1 = s1 + IST_PB, q0 = s0 + IST_Q, q1 = s1 + IST_Q;
      u = sub == 0 ? ts[0] : (sub == 1 ? ts[1] : ts[2]);
    } else {
      const int foot = role >> 2;
      const int32_t* e = schedule + ((long)foot * n_points + k) * IS_STRIDE;
      const double* ph = foot ? ph_r + b * stride_r : ph_l + b * stride_l;
      const int kind = e[IS_KIND];
      const double* A = ph + (long)e[IS_A] * IP_STRIDE;
      const double* B = ph + (long)e[IS_B] * IP_STRIDE;
      if (kind == IK_STANCE) {
        // append_stance_phase: the phase's transform and force; u = 0 and equal end points make the
        // interpolation below return them unchanged
        p0 = p1 = A + IP_POS, q0 = q1 = A + IP_QUAT, fz = A + IP_FORCE, u = 0.0;
      } else {
        // transform_interpolator (:80-103) over a half swing: translation linear, rotation slerp, no force
        p0 = kind == IK_SWING_UP ? A + IP_POS : A + IP_MPOS;
        q0 = kind == IK_SWING_UP ? A + IP_QUAT : A + IP_MQUAT;
        p1 = kind == IK_SWING_UP ? A + IP_MPOS : B + IP_POS;
        q1 = kind == IK_SWING_UP ? A + IP_MQUAT : B + IP_QUAT;
        u = linspace01(e[IS_J], e[IS_N]);
      }
    }
    double pos[3], q[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) pos[c] = lerp(p0[c], p1[c], u);
    if (q0 != q1) {
      slerp(q0, q1, u, q);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = q0[c];
    }
    double* o = blk + sub * ns;
    if (role == 8) {
      o[IST_PB] = pos[0], o[IST_PB + 1] = pos[1], o[IST_PB + 2] = pos[2];
      o[IST_Q] = q[0], o[IST_Q + 1] = q[1], o[IST_Q + 2] = q[2], o[IST_Q + 3] = q[3];
    } else {
      // p = translation + R(q) d, R = I + 2 w [v]x + 2 [v]x^2
      const double* d = s0 + 9 * role + 6;
      const double d0 = d[0], d1 = d[1], d2 = d[2];
      const double c0 = q[1] * d2 - q[2] * d1, c1 = q[2] * d0 - q[0] * d2, c2 = q[0] * d1 - q[1] * d0;
      const double e0 = q[1] * c2 - q[2] * c1, e1 = q[2] * c0 - q[0] * c2, e2 = q[0] * c1 - q[1] * c0;
      o += 9 * role;
      o[0] = pos[0] + d0 + 2.0 * q[3] * c0 + 2.0 * e0;
      o[1] = pos[1] + d1 + 2.0 * q[3] * c1 + 2.0 * e1;
      o[2] = pos[2] + d2 + 2.0 * q[3] * c2 + 2.0 * e2;
      o[3] = fz ? fz[0] : 0.0, o[4] = fz ? fz[1] : 0.0, o[5] = fz ? fz[2] : 0.0;
      o[6] = d0, o[7] = d1, o[8] = d2;
    }
  }
  // kinematic_tree_state_interpolator (:369-393) and the CoM (:419-423): contiguous in the state block
  for (int idx = lane; idx < n_it * nl; idx += 32) {
    const int s = (idx >= nl) + (idx >= 2 * nl), i = idx - s * nl;
    const long b = s == 0 ? bs[0] : (s == 1 ? bs[1] : bs[2]);
    const double t = s == 0 ? ts[0] : (s == 1 ? ts[1] : ts[2]);
    blk[s * ns + IST_S + i] = lerp(initial[b * ns + IST_S + i], final_[b * ns + IST_S + i], t);
  }
  __syncwarp();
  if (states) {
    double* so = states + item0 * ns;
    for (int j = lane; j < n_it * ns; j += 32) so[j] = blk[j];
  }
  if (x) {
    // decision vector: per point p, f (6 of the 15 variables of the point), then base position + quaternion
    // (7 contiguous), joints + CoM (26 contiguous); host-checked: n_joints == 23
    for (int idx = lane; idx < n_it * IZ_WRITTEN; idx += 32) {
      const int s = idx / IZ_WRITTEN, j = idx - s * IZ_WRITTEN;
      int src, dst;
      if (j < 48) {
        const int pt = j / 6, c = j - 6 * pt;
        src = 9 * pt + c, dst = 15 * pt + IZ_P + c;
      } else if (j < 55) {
        src = IST_PB + (j - 48), dst = IZ_PB + (j - 48);
      } else {
        src = IST_S + (j - 55), dst = IZ_S + (j - 55);
      }
      const long b = s == 0 ? bs[0] : (s == 1 ? bs[1] : bs[2]);
      const long k = s == 0 ? ks[0] : (s == 1 ? ks[1] : ks[2]);
      x[b * x_stride + (knot0 + k) * IZ_N + dst] = blk[s * ns + src];
    }
  }
}

}  // namespace hb

// kino_kin.cu -- floating-base kinematics block of the kinodynamic NLP: one warp per (instance, knot).
//
// Rows evaluated here (reference file:line of the expression each one restates):
//   unit quaternion            planner.py:276-282
//   contact FK consistency     planner.py:590-632  (expressions/kinematics.py:217-308)
//   CoM consistency            planner.py:284-306  (expressions/kinematics.py:134-214)
//   angular momentum consist.  planner.py:308-339  (expressions/kinematics.py:11-131, quaternion.py:25-51)
//   minimum feet distance      planner.py:360-383  (expressions/kinematics.py:311-394)
// Costs: frame orientation (planner.py:449-477, kinematics.py:397-491), base quaternion (:479-491,
// quaternion.py:54-85), quaternion velocity (:493-503), joint regularisation (:505-520).
//
// Algorithm (not the reference's: CasADi differentiates an expression graph; here the tree structure
// is used directly).
//   1. primal forward kinematics by depth level (lane = body), world transforms, twists, world
//      inertias staged in shared memory; CoM / momentum sums by warp shuffles;
//   2. Jacobian by forward mode: lane j carries the unit tangent of variable j (4 quaternion + 23 joint
//      directions).  A direction rotates one sub-tree rigidly about its joint, so the tangents of the
//      contact points, of the CoM, of P_dot and of the angular momentum are closed-form in the staged
//      primal state and in the composite (sub-tree) moments -- column j of the kinematic rows, no sweep;
//   3. Hessian: the adjoints of the Lagrangian (true multipliers as seeds) once per warp, lane = body
//      (primal_adjoint_pass); then only the TANGENT of those adjoints along every direction, as
//      (direction, body) tasks packed on the lanes by a host-built schedule (kin_tangent_sweep_packed):
//      column j of the Lagrangian Hessian restricted to (q, s) x (vb, qd, sd, q, s).
//   The file also keeps the first algorithm for the Jacobian -- a row-per-lane adjoint sweep in fp64
//   (kin_backward: 32 rows = 24 FK, 3 CoM, 3 momentum, feet distance, frame-orientation trace) -- as an
//   option (hb_set_option) and cross-check, and the unpacked tangent sweep (HB_SWEEP_PACKED 0).
#include "kino_const.cuh"

namespace hb {

// per-body staging in shared memory (doubles): origin, CoM arm, twist, world axis, world inertia, rotation,
// I.w, I.hbar, rho = o - o_parent, and (Hessian pass) the primal adjoint totals (n, F, wbar, vbar) plus the
// local primal adjoints cbar, cdbar
enum {
  SB_O = 0, SB_D = 3, SB_W = 6, SB_V = 9, SB_AX = 12, SB_I = 15, SB_R = 21, SB_L = 30, SB_IH = 33, SB_RHO = 36,
  SB_ACC = 39, SB_STRIDE = 51,
  SB_CB = SB_R, SB_CDB = SB_R + 3  // alias the rotation: it is dead once the frames have been staged
};
enum { HSTAGE = 27 * 57 };  // per-warp staging of the Hessian columns (and, before that, the Jacobian rows)
#ifndef HB_KIN_STAGE
#define HB_KIN_STAGE 1   // Hessian kernel: stage + batched scatter (2.47 -> 2.25 ms on B200)
#endif
#ifndef HB_KIN_STAGE_F
#define HB_KIN_STAGE_F 0  // Jacobian-only kernel: staging costs a CTA per SM
#endif
#ifndef HB_FWD_JAC
#define HB_FWD_JAC 1      // Hessian kernel: Jacobian columns from forward tangents instead of a second sweep
#endif
#if HB_FWD_JAC && !HB_KIN_STAGE
#error "HB_FWD_JAC stages the Jacobian columns: it needs HB_KIN_STAGE"
#endif
#ifndef KIN_SCAT
#define KIN_SCAT 17  // scatter-map loads in flight per lane in the Hessian kernel's two scatters (8: 5.3 k -> see DESIGN.md)
#endif
#ifndef HB_SWEEP_PACKED
#define HB_SWEEP_PACKED 1  // Hessian kernel: (direction, body) tasks packed on the lanes (kin_tangent_sweep_packed)
#endif
#if HB_SWEEP_PACKED && !HB_KIN_STAGE
#error "HB_SWEEP_PACKED accumulates in the staged Hessian columns: it needs HB_KIN_STAGE"
#endif

struct KinSmem {
  int bodies, arms, fdu, fdal, fdar, G, z, gbuf, lamk, slot, stage, total;
};
__host__ __device__ inline KinSmem kin_smem_layout(int nb, int n_slots, bool with_hess) {
  KinSmem s;
  int o = 0;
  s.bodies = o;
  o += nb * SB_STRIDE;
  s.arms = o;
  o += 24;
  s.fdu = o;
  o += 3;
  s.fdal = o;
  o += 3;
  s.fdar = o;
  o += 3;
  s.G = o;
  o += 9;
  s.z = o;
  o += NZ + 1;
  s.gbuf = o;
  o += 58;
  s.lamk = o;  // multipliers of the 32 kinematic rows of the knot (Hessian variant)
  o += with_hess ? 32 : 0;
  s.slot = o;
  o += n_slots * 32 * 12;  // fp64 sweep: adjoints; Hessian sweep: their tangents (primal totals are per body)
  s.stage = o;
  // Values are staged per warp and scattered after the sweep with 8 map loads in flight per lane.
  // Measured on B200 (config 3): direct stores from the sweep 2.47 ms (17 % of the stall samples sat on
  // the store waiting for its own map load); staging with a naive scatter loop 2.54 ms; staging with
  // the batched scatter 2.25 ms.
  o += with_hess ? (HB_KIN_STAGE ? HSTAGE : 0) : (HB_KIN_STAGE_F ? 928 : 0);
  s.total = o;
  return s;
}

// direction data of one Hessian lane: every link state tangent is closed-form in these
struct Dir {
  D3 alpha, pi, wpi, vpi, u;
  unsigned mask;
};

template <class T>
struct St {
  V3<T> o, d, w, v, ax;
};

__device__ __forceinline__ St<double> load_state(const double* sb, int l, const Dir&, double) {
  const double* b = sb + l * SB_STRIDE;
  St<double> s;
  s.o = ld3(b + SB_O);
  s.d = ld3(b + SB_D);
  s.w = ld3(b + SB_W);
  s.v = ld3(b + SB_V);
  s.ax = ld3(b + SB_AX);
  return s;
}
__device__ __forceinline__ St<Dual> load_state(const double* sb, int l, const Dir& dir, Dual) {
  const double* b = sb + l * SB_STRIDE;
  const D3 o = ld3(b + SB_O), d = ld3(b + SB_D), w = ld3(b + SB_W), v = ld3(b + SB_V), ax = ld3(b + SB_AX);
  const double in = ((dir.mask >> l) & 1u) ? 1.0 : 0.0;
  const D3 a = scale(in, dir.alpha);
  const D3 uu = scale(in, dir.u);
  const D3 r = o - dir.pi;
  const D3 to = cross(a, r);
  const D3 tw = cross(a, w - dir.wpi) + uu;
  const D3 tv = cross(dir.wpi, to) + cross(a, (v - dir.vpi) - cross(dir.wpi, r)) + cross(uu, r);
  St<Dual> s;
  s.o = lift<Dual>(o, to);
  s.d = lift<Dual>(d, cross(a, d));
  s.w = lift<Dual>(w, tw);
  s.v = lift<Dual>(v, tv);
  s.ax = lift<Dual>(ax, cross(a, ax));
  return s;
}

// world inertia of body l applied to x
__device__ __forceinline__ D3 iapply(const double* sb, int l, const Dir&, D3 x) {
  return symmul(sb + l * SB_STRIDE + SB_I, x);
}
__device__ __forceinline__ V3<Dual> iapply(const double* sb, int l, const Dir& dir, V3<Dual> x) {
  const double* I = sb + l * SB_STRIDE + SB_I;
  const D3 xv = v3<double>(x.x.v, x.y.v, x.z.v), xd = v3<double>(x.x.d, x.y.d, x.z.d);
  const double in = ((dir.mask >> l) & 1u) ? 1.0 : 0.0;
  const D3 a = scale(in, dir.alpha);
  const D3 Ix = symmul(I, xv);
  // d(I^w) x = [a]x I x - I [a]x x
  const D3 t = symmul(I, xd) + cross(a, Ix) - symmul(I, cross(a, xv));
  return lift<Dual>(Ix, t);
}

// I^w w_l and I^w hbar with the primal products staged in shared memory (SB_L, SB_IH)
__device__ __forceinline__ D3 iw_omega(const double* sb, int l, const Dir&, const St<double>&) {
  return ld3(sb + l * SB_STRIDE + SB_L);
}
__device__ __forceinline__ V3<Dual> iw_omega(const double* sb, int l, const Dir& dir, const St<Dual>& s) {
  const double* I = sb + l * SB_STRIDE + SB_I;
  const D3 Lw = ld3(sb + l * SB_STRIDE + SB_L);
  const D3 wv = v3<double>(s.w.x.v, s.w.y.v, s.w.z.v), wd = v3<double>(s.w.x.d, s.w.y.d, s.w.z.d);
  const double in = ((dir.mask >> l) & 1u) ? 1.0 : 0.0;
  const D3 a = scale(in, dir.alpha);
  return lift<Dual>(Lw, symmul(I, wd - cross(a, wv)) + cross(a, Lw));
}
__device__ __forceinline__ D3 iw_hbar(const double* sb, int l, const Dir&, D3 hb, double) {
  return symmul(sb + l * SB_STRIDE + SB_I, hb);  // seeds differ per lane in the fp64 sweep
}
// rho_l = o_l - o_parent and the parent's angular velocity, without rebuilding the parent state
__device__ __forceinline__ void parent_link(const double* sb, int l, int p, const Dir&, D3& rho, D3& wp) {
  rho = ld3(sb + l * SB_STRIDE + SB_RHO);
  wp = ld3(sb + p * SB_STRIDE + SB_W);
}
__device__ __forceinline__ void parent_link(const double* sb, int l, int p, const Dir& dir, V3<Dual>& rho,
                                            V3<Dual>& wp) {
  const D3 r = ld3(sb + l * SB_STRIDE + SB_RHO), w = ld3(sb + p * SB_STRIDE + SB_W);
  const double in = ((dir.mask >> p) & 1u) ? 1.0 : 0.0;
  const D3 a = scale(in, dir.alpha);
  rho = lift<Dual>(r, cross(a, r));
  wp = lift<Dual>(w, cross(a, w - dir.wpi) + scale(in, dir.u));
}
// branch accumulators in shared memory, component-major and lane-minor
__device__ __forceinline__ void slot_get(D3& v, const double* sl, int c) {
  v.x += sl[(c)*32];
  v.y += sl[(c + 1) * 32];
  v.z += sl[(c + 2) * 32];
}
__device__ __forceinline__ void slot_get(V3<Dual>& v, const double* sl, int c) {
  v.x = v.x + mkdual(sl[(2 * c) * 32], sl[(2 * c + 1) * 32]);
  v.y = v.y + mkdual(sl[(2 * c + 2) * 32], sl[(2 * c + 3) * 32]);
  v.z = v.z + mkdual(sl[(2 * c + 4) * 32], sl[(2 * c + 5) * 32]);
}
__device__ __forceinline__ void slot_add(double* sl, int c, D3 v) {
  sl[(c)*32] += v.x;
  sl[(c + 1) * 32] += v.y;
  sl[(c + 2) * 32] += v.z;
}
__device__ __forceinline__ void slot_add(double* sl, int c, V3<Dual> v) {
  sl[(2 * c) * 32] += v.x.v;
  sl[(2 * c + 1) * 32] += v.x.d;
  sl[(2 * c + 2) * 32] += v.y.v;
  sl[(2 * c + 3) * 32] += v.y.d;
  sl[(2 * c + 4) * 32] += v.z.v;
  sl[(2 * c + 5) * 32] += v.z.d;
}
__device__ __forceinline__ V3<Dual> iw_hbar(const double* sb, int l, const Dir& dir, D3 hb, Dual) {
  const double* I = sb + l * SB_STRIDE + SB_I;
  const D3 Ih = ld3(sb + l * SB_STRIDE + SB_IH);
  const double in = ((dir.mask >> l) & 1u) ? 1.0 : 0.0;
  const D3 a = scale(in, dir.alpha);
  return lift<Dual>(Ih, cross(a, Ih) - symmul(I, cross(a, hb)));
}

template <class T>
struct Seeds {
  D3 wc;            // d Phi / d (sum_l m_l c_l)   (constant along every direction)
  D3 hb;            // d Phi / d h_ang             (constant along every direction)
  V3<T> footF[2];   // force / torque (about the foot body origin) applied on the two foot bodies
  V3<T> footN[2];
  V3<T> chestN;     // torque on the frame-orientation body
};

// Right-trivialised angular velocity map Omega(a, b) = 2(-b_w a_v + a_w b_v - b_v x a_v)
// (expressions/quaternion.py:42), bilinear in (quaternion a, its rate b).
template <class A, class B>
__device__ __forceinline__ V3<typename Prom<A, B>::type> omega_map(const A* a, const B* b) {
  typedef typename Prom<A, B>::type T;
  V3<A> av = v3<A>(a[0], a[1], a[2]);
  V3<B> bv = v3<B>(b[0], b[1], b[2]);
  V3<T> c = cross(bv, av);
  return v3<T>(2.0 * (a[3] * b[0] - b[3] * a[0] - c.x), 2.0 * (a[3] * b[1] - b[3] * a[1] - c.y),
               2.0 * (a[3] * b[2] - b[3] * a[2] - c.z));
}

// One adjoint sweep over the tree.  `emit.joint(l, sbar, sdbar)` receives d Phi / d s_{l-1} and
// d Phi / d s_dot_{l-1}; the totals at the root are returned for the base chain rule.
template <class T, class Emit>
__device__ __forceinline__ void kin_backward(const KinTopo& C, const double* sb, const double* zs, const Dir& dir,
                                             const Seeds<T>& S, const V3<T>& xc, const V3<T>& xd, double* slot,
                                             Emit& emit, V3<T>& n0, V3<T>& w0, V3<T>& v0) {
  const int nb = C.nb;
  const int lane = threadIdx.x & 31;
  constexpr int SW = sizeof(T) / sizeof(double);  // doubles per T
  for (int i = 0; i < C.n_slots * 12 * SW; ++i) slot[i * 32 + lane] = 0.0;
  V3<T> cn = vzero<T>(), cF = vzero<T>(), cw = vzero<T>(), cv = vzero<T>();
  V3<T> An = vzero<T>(), AF = vzero<T>(), Aw = vzero<T>(), Av = vzero<T>();
  for (int l = nb - 1; l >= 0; --l) {
    emit.prefetch(l);
    if (!C.carry[l]) {
      cn = vzero<T>();
      cF = vzero<T>();
      cw = vzero<T>();
      cv = vzero<T>();
    }
    if (l == 0) {
      cn = cn + An;
      cF = cF + AF;
      cw = cw + Aw;
      cv = cv + Av;
    }
    const St<T> s = load_state(sb, l, dir, T());
    // ---- local adjoints of body l
    {
      const double m = C.mass[l];
      const V3<T> c = s.o + s.d;
      const V3<T> cd = s.v + cross(s.w, s.d);
      const V3<T> cbar = scale(m, cross(cd - xd, S.hb) + S.wc);
      const V3<T> cdbar = scale(m, cross(S.hb, c - xc));
      const V3<T> Iw = iw_omega(sb, l, dir, s);
      const V3<T> Ih = iw_hbar(sb, l, dir, S.hb, T());
      cF = cF + cbar;
      cn = cn + cross(s.d, cbar) + cross(s.d, cross(cdbar, s.w)) + cross(Iw, S.hb) + cross(Ih, s.w);
      cv = cv + cdbar;
      cw = cw + cross(s.d, cdbar) + Ih;
      if (l == C.foot_body[0]) {
        cF = cF + S.footF[0];
        cn = cn + S.footN[0];
      }
      if (l == C.foot_body[1]) {
        cF = cF + S.footF[1];
        cn = cn + S.footN[1];
      }
      if (l == C.chest_body) cn = cn + S.chestN;
    }
    if (C.slot[l] >= 0) {
      const double* sl = slot + C.slot[l] * 12 * SW * 32 + lane;
      slot_get(cn, sl, 0);
      slot_get(cF, sl, 3);
      slot_get(cw, sl, 6);
      slot_get(cv, sl, 9);
    }
    if (l == 0) break;
    // ---- joint outputs
    emit.joint(l, dot(s.ax, cn), dot(s.ax, cw));
    // ---- contribution to the parent
    const int p = C.parent[l];
    V3<T> rho, wpar;
    parent_link(sb, l, p, dir, rho, wpar);
    const double sd = zs[Z_SD + l - 1];
    const V3<T> pn = cn + cross(rho, cF) + scale(sd, cross(s.ax, cw)) + cross(rho, cross(cv, wpar));
    const V3<T> pw = cw + cross(rho, cv);
    if (p == l - 1) {
      cn = pn;
      cw = pw;  // cF, cv carry unchanged
    } else if (p == 0) {
      An = An + pn;
      AF = AF + cF;
      Aw = Aw + pw;
      Av = Av + cv;
    } else {
      double* sl = slot + C.slot[p] * 12 * SW * 32 + lane;
      slot_add(sl, 0, pn);
      slot_add(sl, 3, cF);
      slot_add(sl, 6, pw);
      slot_add(sl, 9, cv);
    }
  }
  n0 = cn;
  w0 = cw;
  v0 = cv;
}

// ---------------------------------------------------------------------------------------------------
// Hessian pass, split in two so that no lane recomputes what is identical for all of them:
//   (1) primal_adjoint_pass: the adjoints of the Lagrangian (true multipliers as seeds) are the same
//       for every direction; they are computed ONCE per warp, lane = body, accumulated leaf-to-root by
//       depth level, and left in shared memory (SB_ACC totals, SB_CB / SB_CDB local terms);
//   (2) kin_tangent_sweep: each lane propagates only the TANGENT of those adjoints along its
//       direction (product rule with the staged primal values) -- 2/3 of the dual-number arithmetic
//       and half of its registers.
struct SeedsP {
  D3 wc, hb, footF[2], footN[2], chestN;
};
struct SeedsT {
  D3 footF[2], footN[2], chestN;  // wc and hb are constants: zero tangent
};

__device__ __forceinline__ void primal_adjoint_pass(const KinTopo& T, double* sb, const double* zs, const SeedsP& S,
                                                    D3 xc, D3 xd, int my_depth, int my_rank, int my_parent,
                                                    double my_mass) {
  const int lane = threadIdx.x & 31;
  const int nb = T.nb;
  if (lane < nb) {
    double* bl = sb + lane * SB_STRIDE;
    const D3 o = ld3(bl + SB_O), d = ld3(bl + SB_D), w = ld3(bl + SB_W), v = ld3(bl + SB_V);
    const double m = my_mass;
    const D3 c = o + d;
    const D3 cd = v + cross(w, d);
    const D3 cbar = scale(m, cross(cd - xd, S.hb) + S.wc);
    const D3 cdbar = scale(m, cross(S.hb, c - xc));
    const D3 Iw = ld3(bl + SB_L), Ih = ld3(bl + SB_IH);
    D3 F = cbar;
    D3 n = cross(d, cbar) + cross(d, cross(cdbar, w)) + cross(Iw, S.hb) + cross(Ih, w);
    const D3 wb = cross(d, cdbar) + Ih;
    if (lane == T.foot_body[0]) {
      F = F + S.footF[0];
      n = n + S.footN[0];
    }
    if (lane == T.foot_body[1]) {
      F = F + S.footF[1];
      n = n + S.footN[1];
    }
    if (lane == T.chest_body) n = n + S.chestN;
    st3(bl + SB_ACC, n);
    st3(bl + SB_ACC + 3, F);
    st3(bl + SB_ACC + 6, wb);
    st3(bl + SB_ACC + 9, cdbar);
    st3(bl + SB_CB, cbar);
    st3(bl + SB_CDB, cdbar);
  }
  __syncwarp();
  // (depth, sibling rank) steps that have at least one body, deepest first: children of one parent are
  // serialised by rank so that the read-modify-write of the parent's totals needs no atomics
  for (int st = 0; st < T.n_steps; ++st) {
    if (my_depth == T.step_depth[st] && my_rank == T.step_rank[st]) {
      const double* bl = sb + lane * SB_STRIDE;
      double* bp = sb + my_parent * SB_STRIDE;
      const D3 n = ld3(bl + SB_ACC), F = ld3(bl + SB_ACC + 3), wb = ld3(bl + SB_ACC + 6), vb = ld3(bl + SB_ACC + 9);
      const D3 ax = ld3(bl + SB_AX), rho = ld3(bl + SB_RHO), wpar = ld3(bp + SB_W);
      const double sd = zs[Z_SD + lane - 1];
      const D3 pn = n + cross(rho, F) + scale(sd, cross(ax, wb)) + cross(rho, cross(vb, wpar));
      const D3 pw = wb + cross(rho, vb);
      st3(bp + SB_ACC, ld3(bp + SB_ACC) + pn);
      st3(bp + SB_ACC + 3, ld3(bp + SB_ACC + 3) + F);
      st3(bp + SB_ACC + 6, ld3(bp + SB_ACC + 6) + pw);
      st3(bp + SB_ACC + 9, ld3(bp + SB_ACC + 9) + vb);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ D3 tangent_of(const V3<Dual>& a) { return v3<double>(a.x.d, a.y.d, a.z.d); }
__device__ __forceinline__ D3 primal_of(const V3<Dual>& a) { return v3<double>(a.x.v, a.y.v, a.z.v); }

template <class Emit>
__device__ __forceinline__ void kin_tangent_sweep(const KinTopo& C, const double* sb, const double* zs,
                                                  const Dir& dir, D3 hb, const SeedsT& S, D3 txc, D3 txd, double* slot,
                                                  Emit& emit, D3& n0, D3& w0, D3& v0) {
  const int nb = C.nb;
  const int lane = threadIdx.x & 31;
  const D3 zero = v3<double>(0.0, 0.0, 0.0);
  for (int i = 0; i < C.n_slots * 12; ++i) slot[i * 32 + lane] = 0.0;
  D3 cn = zero, cF = zero, cw = zero, cv = zero;
  D3 An = zero, AF = zero, Aw = zero, Av = zero;
  for (int l = nb - 1; l >= 0; --l) {
    if (!C.carry[l]) cn = cF = cw = cv = zero;
    if (l == 0) {
      cn = cn + An;
      cF = cF + AF;
      cw = cw + Aw;
      cv = cv + Av;
    }
    const double* bl = sb + l * SB_STRIDE;
    const St<Dual> s = load_state(sb, l, dir, Dual());
    const D3 d = primal_of(s.d), w = primal_of(s.w), ax = primal_of(s.ax);
    const D3 to = tangent_of(s.o), td = tangent_of(s.d), tw = tangent_of(s.w), tv = tangent_of(s.v),
             tax = tangent_of(s.ax);
    const double in = ((dir.mask >> l) & 1u) ? 1.0 : 0.0;
    const D3 a = scale(in, dir.alpha);
    {
      const double m = C.mass[l];
      const double* I = bl + SB_I;
      const D3 cbar = ld3(bl + SB_CB), cdbar = ld3(bl + SB_CDB), Ih = ld3(bl + SB_IH), Iw = ld3(bl + SB_L);
      const D3 tc = to + td;
      const D3 tcd = tv + cross(tw, d) + cross(w, td);
      const D3 tcbar = scale(m, cross(tcd - txd, hb));
      const D3 tcdbar = scale(m, cross(hb, tc - txc));
      const D3 tIw = symmul(I, tw - cross(a, w)) + cross(a, Iw);
      const D3 tIh = cross(a, Ih) - symmul(I, cross(a, hb));
      cF = cF + tcbar;
      cn = cn + cross(td, cbar) + cross(d, tcbar) + cross(td, cross(cdbar, w)) +
           cross(d, cross(tcdbar, w) + cross(cdbar, tw)) + cross(tIw, hb) + cross(tIh, w) + cross(Ih, tw);
      cv = cv + tcdbar;
      cw = cw + cross(td, cdbar) + cross(d, tcdbar) + tIh;
      if (l == C.foot_body[0]) {
        cF = cF + S.footF[0];
        cn = cn + S.footN[0];
      }
      if (l == C.foot_body[1]) {
        cF = cF + S.footF[1];
        cn = cn + S.footN[1];
      }
      if (l == C.chest_body) cn = cn + S.chestN;
    }
    if (C.slot[l] >= 0) {
      const double* sl = slot + C.slot[l] * 12 * 32 + lane;
      slot_get(cn, sl, 0);
      slot_get(cF, sl, 3);
      slot_get(cw, sl, 6);
      slot_get(cv, sl, 9);
    }
    if (l == 0) break;
    // primal totals of body l (all children included)
    const D3 np = ld3(bl + SB_ACC), Fp = ld3(bl + SB_ACC + 3), wp = ld3(bl + SB_ACC + 6), vp = ld3(bl + SB_ACC + 9);
    emit.joint_t(l, dot(tax, np) + dot(ax, cn), dot(tax, wp) + dot(ax, cw));
    const int p = C.parent[l];
    const D3 rho = ld3(bl + SB_RHO), wpar = ld3(sb + p * SB_STRIDE + SB_W);
    const double inp = ((dir.mask >> p) & 1u) ? 1.0 : 0.0;
    const D3 ap = scale(inp, dir.alpha);
    const D3 trho = cross(ap, rho);
    const D3 twpar = cross(ap, wpar - dir.wpi) + scale(inp, dir.u);
    const double sd = zs[Z_SD + l - 1];
    const D3 pn = cn + cross(trho, Fp) + cross(rho, cF) + scale(sd, cross(tax, wp) + cross(ax, cw)) +
                  cross(trho, cross(vp, wpar)) + cross(rho, cross(cv, wpar) + cross(vp, twpar));
    const D3 pw = cw + cross(trho, vp) + cross(rho, cv);
    if (p == l - 1) {
      cn = pn;
      cw = pw;
    } else if (p == 0) {
      An = An + pn;
      AF = AF + cF;
      Aw = Aw + pw;
      Av = Av + cv;
    } else {
      double* sl = slot + C.slot[p] * 12 * 32 + lane;
      slot_add(sl, 0, pn);
      slot_add(sl, 3, cF);
      slot_add(sl, 6, pw);
      slot_add(sl, 9, cv);
    }
  }
  n0 = cn;
  w0 = cw;
  v0 = cv;
}

// ---------------------------------------------------------------------------------------------------
// Packed tangent sweep.  In kin_tangent_sweep lane d walks all the bodies although direction d only
// moves its own sub-tree: of the 27 x 23 body steps 188 carry non-zero state tangents and 92 more are
// pure propagation through the ancestors of the joint.  Here those (direction, body) steps are TASKS,
// list-scheduled by the host (api.cu::build_sweep_schedule) on the 32 lanes in ~12 rounds instead of
// 23: in every round a lane executes the task sched[round][lane] -- whatever direction it belongs to.
//   * direction data (alpha, pivot, ...) and seed tangents stay in the registers of the OWNER lane d and
//     are fetched with warp shuffles by the lane that executes a task of direction d;
//   * a lane keeps the running adjoint of its chain in registers; where a chain ends it is added to a
//     shared-memory slot of the body it feeds (the schedule never lets two chains hit one slot in the
//     same round, so the order of the additions is fixed);
//   * the coupling of ALL bodies through the total momentum, -hb.(P x Pdot)/M, is what would make every
//     step non-zero; its Hessian is the rank-structured term -(1/M) hb.(dP_i x dPdot_j + dP_j x dPdot_i),
//     written into the stage beforehand (cross_terms), and the sweep runs with d xc = d xdot_c = 0.
struct DirData {
  D3 alpha, pi, wpi, vpi, u;
};
__device__ __forceinline__ D3 shfl3(const D3& v, int src) {
  return v3<double>(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                    __shfl_sync(0xffffffffu, v.z, src));
}

__device__ __forceinline__ void kin_tangent_sweep_packed(const KinTopo& T, const double* sb, const double* zs,
                                                         const double* qdat, const double* massv, const SeedsT& ownS,
                                                         D3 hb, double* pslot, double* stage) {
  const int lane = threadIdx.x & 31;
  const D3 zero = v3<double>(0.0, 0.0, 0.0);
  for (int i = lane; i < T.n_pslots * 12; i += 32) pslot[i] = 0.0;
  __syncwarp();
  D3 cn = zero, cF = zero, cw = zero, cv = zero;
  unsigned t_next = (unsigned)T.sched[lane];
  for (int r = 0; r < T.n_rounds; ++r) {
    const unsigned t = t_next;
    if (r + 1 < T.n_rounds) t_next = (unsigned)T.sched[(r + 1) * 32 + lane];
    const bool valid = (t & KT_VALID) != 0;
    const int l = (t >> KT_L_SHIFT) & 31, p = (t >> KT_P_SHIFT) & 31, d = (t >> KT_D_SHIFT) & 31;
    const int src = valid ? d : lane;
    const bool heavy = (T.round_heavy_mask >> r) & 1u;  // warp-uniform (constant bank): typed rounds, sweep_schedule.h
    SeedsT S;
    S.footF[0] = S.footF[1] = S.footN[0] = S.footN[1] = S.chestN = zero;
    if ((T.round_seed_mask >> r) & 1u) {  // warp-uniform
      S.footF[0] = shfl3(ownS.footF[0], src);
      S.footF[1] = shfl3(ownS.footF[1], src);
      S.footN[0] = shfl3(ownS.footN[0], src);
      S.footN[1] = shfl3(ownS.footN[1], src);
      S.chestN = shfl3(ownS.chestN, src);
    }
    if (valid) {
      if (t & KT_START) cn = cF = cw = cv = zero;
      const double* bl = sb + l * SB_STRIDE;
      const D3 ax = ld3(bl + SB_AX);
      D3 tax = zero, trho = zero, twpar = zero;
      const D3 rho = ld3(bl + SB_RHO), wpar = ld3(sb + p * SB_STRIDE + SB_W);
      if (heavy) {
        // direction data: a joint direction is the axis / origin / twist of its body (already in shared
        // memory); the four quaternion directions rotate everything about the base origin (qdat: g_a, u_a)
        DirData dir;
        {
          const double* bd = sb + (d < 4 ? 0 : d - 3) * SB_STRIDE;
          dir.wpi = ld3(bd + SB_W);
          dir.vpi = ld3(bd + SB_V);
          if (d < 4) {
            dir.alpha = ld3(qdat + 6 * d);
            dir.u = ld3(qdat + 6 * d + 3);
            dir.pi = zero;
          } else {
            dir.alpha = ld3(bd + SB_AX);
            dir.u = zero;
            dir.pi = ld3(bd + SB_O);
          }
        }
        const double m = massv[l];
        // with typed rounds every task of a heavy round lies in the sub-tree of its direction (KT_IN_L set); the
        // untyped fallback schedule mixes both kinds, hence the select
        const bool in = (t & KT_IN_L) != 0;
        const D3 a = in ? dir.alpha : zero, uu = in ? dir.u : zero;
        const D3 d_ = ld3(bl + SB_D), w = ld3(bl + SB_W);
        const D3 rr = ld3(bl + SB_O) - dir.pi;
        const D3 to = cross(a, rr);
        const D3 td = cross(a, d_);
        const D3 tw = cadd(uu, a, w - dir.wpi);
        const D3 tv = cadd(cadd(cross(dir.wpi, to), a, (ld3(bl + SB_V) - dir.vpi) - cross(dir.wpi, rr)), uu, rr);
        tax = cross(a, ax);
        if (in) {  // local adjoints of body l move with the sub-tree of the direction
          const double* I = bl + SB_I;
          const D3 cbar = ld3(bl + SB_CB), cdbar = ld3(bl + SB_CDB), Ih = ld3(bl + SB_IH), Iw = ld3(bl + SB_L);
          const D3 tc = to + td;
          const D3 tcd = cadd(cadd(tv, tw, d_), w, td);
          const D3 tcbar = scale(m, cross(tcd, hb));
          const D3 tcdbar = scale(m, cross(hb, tc));
          const D3 tIw = cadd(symmul(I, tw - cross(a, w)), a, Iw);
          const D3 tIh = csub(cross(a, Ih), I, cross(a, hb));
          cF = cF + tcbar;
          // two accumulators keep the dependent chains short
          D3 n1 = cadd(cadd(cadd(cn, td, cbar), d_, tcbar), td, cross(cdbar, w));
          D3 n2 = cadd(cadd(cadd(cross(d_, cadd(cross(tcdbar, w), cdbar, tw)), tIw, hb), tIh, w), Ih, tw);
          cn = n1 + n2;
          cv = cv + tcdbar;
          cw = cadd(cadd(cw, td, cdbar), d_, tcdbar) + tIh;
        }
        if (t & KT_IN_P) {  // the parent moves with the direction as well
          trho = cross(dir.alpha, rho);
          twpar = cadd(dir.u, dir.alpha, wpar - dir.wpi);
        }
      }
      // seed tangents (zero where they do not apply).  The feet-distance row couples the two feet: a
      // direction that moves one foot also changes the seed on the OTHER foot, which is why the schedule
      // visits the other leg for such a direction
      if (l == T.foot_body[0]) {
        cF = cF + S.footF[0];
        cn = cn + S.footN[0];
      }
      if (l == T.foot_body[1]) {
        cF = cF + S.footF[1];
        cn = cn + S.footN[1];
      }
      if (l == T.chest_body) cn = cn + S.chestN;
      const int ls = (t >> KT_LOAD_SHIFT) & 63;
      if (ls) {
        const double* sl = pslot + (ls - 1) * 12;
        cn = cn + ld3(sl);
        cF = cF + ld3(sl + 3);
        cw = cw + ld3(sl + 6);
        cv = cv + ld3(sl + 9);
      }
      if (l != 0) {
        double* col = stage + d * 57;
        const double sd = zs[Z_SD + l - 1];
        if (heavy) {
          const D3 np = ld3(bl + SB_ACC), Fp = ld3(bl + SB_ACC + 3), wp = ld3(bl + SB_ACC + 6), vp = ld3(bl + SB_ACC + 9);
          col[7 + l - 1] += dot(tax, wp) + dot(ax, cw);   // rows: vb3 qd4 sd23 q4 s23
          col[34 + l - 1] += dot(tax, np) + dot(ax, cn);
          D3 n1 = cadd(cadd(cn, trho, Fp), rho, cF);
          D3 n2 = cadd(scale(sd, cadd(cross(tax, wp), ax, cw)), trho, cross(vp, wpar));
          cn = cadd(n1 + n2, rho, cadd(cross(cv, wpar), vp, twpar));
          cw = cadd(cadd(cw, trho, vp), rho, cv);
        } else {
          // propagation through an ancestor of the joint: every state tangent is zero, the adjoint tangent moves
          // towards the root with the primal coefficients only
          col[7 + l - 1] += dot(ax, cw);
          col[34 + l - 1] += dot(ax, cn);
          cn = cadd(cadd(cn, rho, cF) + scale(sd, cross(ax, cw)), rho, cross(cv, wpar));
          cw = cadd(cw, rho, cv);
        }
      }
      const int fs = (t >> KT_FLUSH_SHIFT) & 63;
      if (fs) {
        double* sl = pslot + (fs - 1) * 12;
        if (t & KT_STORE) {
          st3(sl, cn);
          st3(sl + 3, cF);
          st3(sl + 6, cw);
          st3(sl + 9, cv);
        } else {
          st3(sl, ld3(sl) + cn);
          st3(sl + 3, ld3(sl + 3) + cF);
          st3(sl + 6, ld3(sl + 6) + cw);
          st3(sl + 9, ld3(sl + 9) + cv);
        }
      }
    }
    __syncwarp();
  }
}

// quaternion maps: g_a (rotation tangent of R(q/|q|) along q_a), u_a = d omega_0 / d q_a,
// w_a = d omega_0 / d q_dot_a.
template <class T>
__device__ __forceinline__ void quat_maps(const T* q, const double* qd, V3<T>* g, V3<T>* u, V3<T>* w) {
  const T r2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  const T r = dsqrt(r2);
  const T ir = 1.0 / r;
  T qh[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) qh[i] = q[i] * ir;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    T t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) t[i] = ((i == a ? 1.0 : 0.0) - qh[i] * qh[a]) * ir;
    g[a] = omega_map(qh, t);
    u[a] = omega_map(t, qd);
    double e[4] = {0.0, 0.0, 0.0, 0.0};
    e[a] = 1.0;
    w[a] = omega_map(qh, e);
  }
}

struct JacEmit {
  const KinoConst* C;
  const int* map;  // jk_map row of this knot
  double* jac;     // instance base
  double* gbuf;    // shared: grad_f accumulation for the kinematic variables
  int lane, k;
  double gscale;   // chest lane: 2 w_frame (phi - 3)
  bool write;
  // local bases inside JK (kino_layout.py::_enumerate_jk)
  __device__ __forceinline__ int base_q() const {
    if (lane < 24) return 4 + lane * 27;
    if (lane < 27) return 4 + 648 + (lane - 24) * 27;
    if (lane < 30) return 4 + 648 + 81 + (lane - 27) * 57 + 7;  // after vb(3), qd(4)
    return -1;
  }
  double* stage;   // non-null: values go to shared memory and are scattered after the sweep
  __device__ __forceinline__ void put(int e, double v) const {
    if (stage) {
      stage[e] = v;
      return;
    }
    const int slot = map[e];
    if (slot >= 0) jac[slot] = v;
  }
  // Scatter slots of the two entries the sweep emits at body l, loaded one step ahead (prefetch(l) is
  // called at the top of the iteration) so that the stores do not wait for their own map loads.
  int sbase, sdbase, pre_s, pre_sd;
  __device__ __forceinline__ void init_bases() {
    sdbase = -1;
    if (lane < 27) sbase = base_q() + 4;
    else if (lane < 30) {
      const int b = 4 + 648 + 81 + (lane - 27) * 57;
      sdbase = b + 11;
      sbase = b + 34;
    } else sbase = lane == 30 ? 4 + 648 + 81 + 171 : -1;
    pre_s = pre_sd = -1;
  }
  __device__ __forceinline__ void prefetch(int l) {
    if (stage || !write || l < 1) return;
    pre_s = sbase >= 0 ? map[sbase + l - 1] : -1;
    pre_sd = sdbase >= 0 ? map[sdbase + l - 1] : -1;
  }
  __device__ __forceinline__ void joint(int l, double sbar, double sdbar) const {
    const int j = l - 1;
    if (lane == 31) {
      gbuf[34 + j] += gscale * sbar;  // gbuf order: vb3 qd4 q4 sd23 s23 -> s at 34
      return;
    }
    if (!write) return;
    if (stage) {
      if (sbase >= 0) stage[sbase + j] = sbar;
      if (sdbase >= 0) stage[sdbase + j] = sdbar;
      return;
    }
    if (pre_s >= 0) jac[pre_s] = sbar;
    if (pre_sd >= 0) jac[pre_sd] = sdbar;
  }
};

struct HessEmit {
  const int* map;  // hk_map row of this knot
  double* hess;    // instance base
  int dirj;        // direction index 0..26, or -1 (idle lane)
  double add_sd, add_s;  // joint-regularisation terms on (sd_j, s_j), (s_j, s_j) for joint lanes
  double* stage;   // per-warp staging (HSTAGE doubles); scattered after the sweep
  bool acc = false;  // packed sweep: the stage already holds the cross terms, entries are added
  __device__ __forceinline__ void prefetch(int) {}
  __device__ __forceinline__ void put(int row, double v) const {
    if (stage) {
      if (acc) stage[dirj * 57 + row] += v;
      else stage[dirj * 57 + row] = v;
      return;
    }
    const int slot = map[dirj * 57 + row];
    if (slot >= 0) hess[slot] = v;
  }
  __device__ __forceinline__ void joint_t(int l, double tsbar, double tsdbar) const {
    if (dirj < 0) return;
    const int j = l - 1;
    const bool own = (dirj - 4 == j);
    put(7 + j, tsdbar + (own ? add_sd : 0.0));   // rows: vb3 qd4 sd23 q4 s23
    put(34 + j, tsbar + (own ? add_s : 0.0));
  }
  __device__ __forceinline__ void joint(int l, Dual sbar, Dual sdbar) const { joint_t(l, sbar.d, sdbar.d); }
};

// Measured on B200 (tools/time_kino.py): the dual sweep is fastest with the full 255 registers
// (2 CTAs/SM; 168 registers spill 1 KB/thread and lose 12%), the fp64-only variant with 168
// registers (3 CTAs/SM, -19%).
template <bool WITH_HESS>
#ifndef KIN_H_MIN_BLOCKS
#define KIN_H_MIN_BLOCKS 2
#endif
#ifndef HB_KIN_JAC_LIST
#define HB_KIN_JAC_LIST 1  // Jacobian scatter through the destination-sorted list (measured 1.145 ms against 1.162 ms in entry order)
#endif
#ifndef HB_KIN_HESS_LIST
#define HB_KIN_HESS_LIST 0  // 1: Hessian scatter through the destination-sorted list.  Measured: 1.161 ms against 1.145 ms in
                            // entry order -- the staged columns are read conflict-free in entry order, in destination order not
#endif
#ifndef KIN_H_THREADS
#define KIN_H_THREADS 128
#endif
__global__ void __launch_bounds__(WITH_HESS ? KIN_H_THREADS : 128, WITH_HESS ? KIN_H_MIN_BLOCKS : 3) kino_kin_kernel(
    const __grid_constant__ KinTopo T,
                                                       const KinoConst* __restrict__ Cp, unsigned mask,
                                                       const double* __restrict__ x, const double* __restrict__ p,
                                                       long p_stride, const double* __restrict__ lam,
                                                       const double* __restrict__ sigma, double* __restrict__ fpart,
                                                       double* __restrict__ grad_f, double* __restrict__ g,
                                                       double* __restrict__ jac, double* __restrict__ hess,
                                                       long batch) {
  extern __shared__ double smem[];
  const KinoConst& C = *Cp;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long wid = (long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int N = T.N;
  if (wid >= batch * N) return;
  const long b = wid / N;
  const int k = (int)(wid % N);
  const KinSmem L = kin_smem_layout(T.nb, T.n_slots, WITH_HESS);
  double* sm = smem + (size_t)warp * L.total;
  double* sb = sm + L.bodies;
  double* zs = sm + L.z;
  double* gbuf = sm + L.gbuf;
  double* lamk = sm + L.lamk;
  const double* xb = x + b * T.n_x + (long)k * T.x_stride;
  const double* pb_ = p + b * p_stride;
  const int nb = T.nb;
  const bool k1 = k >= T.cost_k0;  // knots on which the "apply_to_first_elements=False" expressions exist
  HB_PHASE_INIT

  // ---- every global input of this warp is requested here, addresses from the constant bank: the DRAM
  // latencies of x, the parameter slices, the multipliers and the joint frames overlap each other and
  // the forward kinematics instead of being exposed one by one at their points of use
  // (loads first, shared-memory stores last: a store that waits for its load would hold back, in program
  // order, every load behind it)
  // of the 189 variables of the knot this kernel reads the 24 contact-point positions (FK rows) and the robot block
  // [Z_VB, NZ): 93 doubles in 4 coalesced requests (the contact kernel owns the rest of the point variables)
  double xr[4];
  int xi[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    xi[u] = u == 0 ? (lane < 24 ? 15 * (lane / 3) + Z_P + lane % 3 : -1) : Z_VB + lane + 32 * (u - 1);
    if (xi[u] >= NZ) xi[u] = -1;
    xr[u] = 0.0;
    if (xi[u] >= 0) {
      if (T.zmap_identity) xr[u] = xb[xi[u]];
      else {
        const int zi = C.zmap[xi[u]];
        if (zi >= 0) xr[u] = xb[zi];
      }
    }
  }
  const double* rfq = pb_ + T.po_fq + T.ref_stride * k;
  const double* rbq = pb_ + T.po_bq + T.ref_stride * k;
  const double* rbqv = pb_ + T.po_bqv + T.ref_stride * k;
  const double* rjr = pb_ + T.po_jr + T.ref_stride * k;
  const D3 pre_desc = lane < 8 ? ld3(pb_ + T.po_desc0 + 24 * k + 3 * lane) : v3<double>(0.0, 0.0, 0.0);
  double pre_fq[4] = {0.0, 0.0, 0.0, 1.0};
  if (lane == 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) pre_fq[i] = rfq[i];
  }
  const double pre_bq[4] = {rbq[0], rbq[1], rbq[2], rbq[3]};
  const double pre_bqv = lane < 4 ? rbqv[lane] : 0.0;
  const double pre_jr = lane < HB_N_JOINTS ? rjr[lane] : 0.0;
  const double mass_p = pb_[T.po_mass];
  const bool with_l = WITH_HESS && (mask & HB_EVAL_HESS_L);
  double sg = 0.0, lam_reg = 0.0;
  if (with_l) {
    // multipliers of the 32 kinematic rows of this knot, one per lane, parked in shared memory
    int row;
    if (lane < 24) row = grow(C, (lane / 3) * HB_KF_PT_COUNT + HB_KF_PT_FK, k, lane % 3);
    else if (lane < 27) row = grow(T, HB_KF_COM_KIN, k, lane - 24);
    else if (lane < 30) row = grow(T, HB_KF_MOM_KIN, k, lane - 27);
    else row = grow(T, lane == 30 ? HB_KF_FEET_DIST : HB_KF_UNIT_QUAT, k, 0);
    lam_reg = row >= 0 ? lam[b * T.m + row] : 0.0;
    sg = sigma[b];
  }
  // joint frame constants of this lane's body
  const bool is_link = lane > 0 && lane < nb;
  const int my_depth = is_link ? C.body[lane].depth : -1;
  const int my_parent = is_link ? C.body[lane].parent : 0;
  const int my_rank = is_link ? C.body[lane].sib_rank : -1;
  const double my_mass = lane < nb ? C.body[lane].mass : 0.0;
  D3 b_com = v3<double>(0.0, 0.0, 0.0);
  double b_I[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (lane < nb) {
    b_com = ld3(C.body[lane].com);
#pragma unroll
    for (int i = 0; i < 6; ++i) b_I[i] = C.body[lane].inertia[i];
  }
  double bE[9], bEA[9], bEA2[9];
  D3 b_r = v3<double>(0.0, 0.0, 0.0), b_axis = b_r;
  if (is_link) {
    const BodyC& bc = C.body[lane];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      bE[i] = bc.E[i];
      bEA[i] = bc.EA[i];
      bEA2[i] = bc.EA2[i];
    }
    b_r = ld3(bc.r);
    b_axis = ld3(bc.axis);
  }
  const unsigned my_submask = lane < nb ? C.sub_mask[lane] : 0u;
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (xi[u] >= 0) zs[xi[u]] = xr[u];
  if (WITH_HESS) lamk[lane] = lam_reg;
  for (int i = lane; i < 58; i += 32) gbuf[i] = 0.0;
  __syncwarp();
  HB_PHASE(0, 0);  // inputs loaded and staged

  // ------------------------------------------------------------------ base
  const double q0 = zs[Z_Q], q1 = zs[Z_Q + 1], q2 = zs[Z_Q + 2], q3 = zs[Z_Q + 3];
  const double qq = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
  const double qn = sqrt(qq);
  const double qh[4] = {q0 / qn, q1 / qn, q2 / qn, q3 / qn};
  const double qdv[4] = {zs[Z_QD], zs[Z_QD + 1], zs[Z_QD + 2], zs[Z_QD + 3]};
  if (lane == 0) {
    // R = I + 2 w [v]x + 2 [v]x^2 (liecasadi SO3.from_quat().as_matrix() [ext])
    const double vx = qh[0], vy = qh[1], vz = qh[2], w = qh[3];
    double* R = sb + SB_R;
    R[0] = 1.0 - 2.0 * (vy * vy + vz * vz);
    R[1] = 2.0 * (vx * vy - w * vz);
    R[2] = 2.0 * (vx * vz + w * vy);
    R[3] = 2.0 * (vx * vy + w * vz);
    R[4] = 1.0 - 2.0 * (vx * vx + vz * vz);
    R[5] = 2.0 * (vy * vz - w * vx);
    R[6] = 2.0 * (vx * vz - w * vy);
    R[7] = 2.0 * (vy * vz + w * vx);
    R[8] = 1.0 - 2.0 * (vx * vx + vy * vy);
    st3(sb + SB_O, v3<double>(0.0, 0.0, 0.0));  // positions are relative to the base origin
    st3(sb + SB_W, omega_map(qh, qdv));
    st3(sb + SB_V, ld3(zs + Z_VB));
    st3(sb + SB_AX, v3<double>(0.0, 0.0, 0.0));
  }
  __syncwarp();
  // ------------------------------------------------------------------ forward kinematics by depth
  // the joint rotations do not depend on the parents: computed by all link lanes at once
  double Rl[9];
  double my_sd = 0.0;
  if (is_link) {
    const double s = zs[Z_S + lane - 1];
    my_sd = zs[Z_SD + lane - 1];
    double sn, cs;
    sincos(s, &sn, &cs);
    const double oc = 1.0 - cs;
#pragma unroll
    for (int i = 0; i < 9; ++i) Rl[i] = bE[i] + sn * bEA[i] + oc * bEA2[i];
  }
  for (int d = 1; d <= T.max_depth; ++d) {
    if (my_depth == d) {
      const double* bp = sb + my_parent * SB_STRIDE;
      double* bl = sb + lane * SB_STRIDE;
      const double sd = my_sd;
      const double* Rp = bp + SB_R;
      double R[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[3 * i + j] = Rp[3 * i] * Rl[j] + Rp[3 * i + 1] * Rl[3 + j] + Rp[3 * i + 2] * Rl[6 + j];
#pragma unroll
      for (int i = 0; i < 9; ++i) bl[SB_R + i] = R[i];
      const D3 op = ld3(bp + SB_O), wp = ld3(bp + SB_W), vp = ld3(bp + SB_V);
      const D3 rho = matvec(Rp, b_r);
      const D3 ax = matvec(R, b_axis);
      st3(bl + SB_O, op + rho);
      st3(bl + SB_AX, ax);
      st3(bl + SB_W, wp + scale(sd, ax));
      st3(bl + SB_V, vp + cross(wp, rho));
      st3(bl + SB_RHO, rho);
    }
    __syncwarp();
  }
  HB_PHASE(0, 1);  // forward kinematics by depth
  // ------------------------------------------------------------------ per-body derived quantities + sums
  D3 mc = v3<double>(0.0, 0.0, 0.0), mcd = mc, hl = mc;
  if (lane < nb) {
    double* bl = sb + lane * SB_STRIDE;
    const double* R = bl + SB_R;
    const D3 d = matvec(R, b_com);
    st3(bl + SB_D, d);
    // I^w = R I R^T
    double RI[9];
    const double* I = b_I;
    const double Im[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) RI[3 * i + j] = R[3 * i] * Im[j] + R[3 * i + 1] * Im[3 + j] + R[3 * i + 2] * Im[6 + j];
    double Iw[6];
    const int ii[6] = {0, 0, 0, 1, 1, 2}, jj[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; ++e)
      Iw[e] = RI[3 * ii[e]] * R[3 * jj[e]] + RI[3 * ii[e] + 1] * R[3 * jj[e] + 1] + RI[3 * ii[e] + 2] * R[3 * jj[e] + 2];
#pragma unroll
    for (int e = 0; e < 6; ++e) bl[SB_I + e] = Iw[e];
    const D3 o = ld3(bl + SB_O), w = ld3(bl + SB_W), v = ld3(bl + SB_V);
    const D3 c = o + d;
    const D3 cd = v + cross(w, d);
    mc = scale(my_mass, c);
    mcd = scale(my_mass, cd);
    const D3 Lw = symmul(Iw, w);
    st3(bl + SB_L, Lw);
    hl = cross(mc, cd) + Lw;
  }
  const double M = T.total_mass;
  const D3 Pm = v3<double>(warp_sum(mc.x), warp_sum(mc.y), warp_sum(mc.z));
  const D3 Pd = v3<double>(warp_sum(mcd.x), warp_sum(mcd.y), warp_sum(mcd.z));
  D3 hang = v3<double>(warp_sum(hl.x), warp_sum(hl.y), warp_sum(hl.z));
  hang = hang - scale(1.0 / M, cross(Pm, Pd));
  const D3 xc = scale(1.0 / M, Pm), xcd = scale(1.0 / M, Pd);
  // ------------------------------------------------------------------ frames
  if (lane < 8) {
    const int f = lane >> 2;
    const double* Rf = C.foot_R[f];
    const D3 r = pre_desc;
    const D3 bf = matvec(Rf, r) + ld3(C.foot_t[f]);
    const D3 a = matvec(sb + T.foot_body[f] * SB_STRIDE + SB_R, bf);
    st3(sm + L.arms + 3 * lane, a);
  }
  if (lane == 8) {
    // G = R_chest_body * (R_c * R(qd)^T), qd = desired frame quaternion (not normalised, kinematics.py:447)
    const double vx = pre_fq[0], vy = pre_fq[1], vz = pre_fq[2], w = pre_fq[3];
    double Rd[9];
    Rd[0] = 1.0 - 2.0 * (vy * vy + vz * vz);
    Rd[1] = 2.0 * (vx * vy - w * vz);
    Rd[2] = 2.0 * (vx * vz + w * vy);
    Rd[3] = 2.0 * (vx * vy + w * vz);
    Rd[4] = 1.0 - 2.0 * (vx * vx + vz * vz);
    Rd[5] = 2.0 * (vy * vz - w * vx);
    Rd[6] = 2.0 * (vx * vz - w * vy);
    Rd[7] = 2.0 * (vy * vz + w * vx);
    Rd[8] = 1.0 - 2.0 * (vx * vx + vy * vy);
    double Kc[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        Kc[3 * i + j] = C.chest_R[3 * i] * Rd[3 * j] + C.chest_R[3 * i + 1] * Rd[3 * j + 1] + C.chest_R[3 * i + 2] * Rd[3 * j + 2];
    const double* Rb = sb + T.chest_body * SB_STRIDE + SB_R;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) sm[L.G + 3 * i + j] = Rb[3 * i] * Kc[j] + Rb[3 * i + 1] * Kc[3 + j] + Rb[3 * i + 2] * Kc[6 + j];
  }
  if (lane == 9) {
    const double* RR = sb + T.foot_body[1] * SB_STRIDE + SB_R;
    const double* RL = sb + T.foot_body[0] * SB_STRIDE + SB_R;
    const double* Rf = C.foot_R[1];
    st3(sm + L.fdu, matvec(RR, v3<double>(Rf[1], Rf[4], Rf[7])));  // y axis of the reference sole frame
    st3(sm + L.fdal, matvec(RL, ld3(C.foot_t[0])));
    st3(sm + L.fdar, matvec(RR, ld3(C.foot_t[1])));
  }
  __syncwarp();
  const double* G = sm + L.G;
  const double phi = G[0] + G[4] + G[8];
  const D3 mG = v3<double>(G[5] - G[7], G[6] - G[2], G[1] - G[3]);
  const D3 fdu = ld3(sm + L.fdu), fdal = ld3(sm + L.fdal), fdar = ld3(sm + L.fdar);
  const D3 oL = ld3(sb + T.foot_body[0] * SB_STRIDE + SB_O), oR = ld3(sb + T.foot_body[1] * SB_STRIDE + SB_O);
  const D3 fdDelta = (oL + fdal) - (oR + fdar);
  const double feet_y = dot(fdu, fdDelta);

  HB_PHASE(0, 2);  // per-body quantities, sums, frames
  // ------------------------------------------------------------------ values: g rows, costs, grad_f (simple terms)
  const bool want_g = (mask & HB_EVAL_G) != 0;
  double* gb = g + b * T.m;
  if (want_g) {
    if (lane < 8 && k1) {
      const int f = lane >> 2;
      const D3 a = ld3(sm + L.arms + 3 * lane);
      const D3 o = ld3(sb + T.foot_body[f] * SB_STRIDE + SB_O);
      const int r0 = grow(C, lane * HB_KF_PT_COUNT + HB_KF_PT_FK, k, 0);  // lane-indexed family: global table
      const double* pp = zs + 15 * lane + Z_P;
      if (r0 >= 0) {
        gb[r0] = pp[0] - (zs[Z_PB] + o.x + a.x);
        gb[r0 + 1] = pp[1] - (zs[Z_PB + 1] + o.y + a.y);
        gb[r0 + 2] = pp[2] - (zs[Z_PB + 2] + o.z + a.z);
      }
    }
    if (lane == 8) {
      int r0 = grow(T, HB_KF_UNIT_QUAT, k, 0);
      if (r0 >= 0) gb[r0] = qq;
      r0 = grow(T, HB_KF_COM_KIN, k, 0);
      if (r0 >= 0) {
        gb[r0] = zs[Z_COM] - (zs[Z_PB] + xc.x);
        gb[r0 + 1] = zs[Z_COM + 1] - (zs[Z_PB + 1] + xc.y);
        gb[r0 + 2] = zs[Z_COM + 2] - (zs[Z_PB + 2] + xc.z);
      }
      r0 = grow(T, HB_KF_MOM_KIN, k, 0);
      if (r0 >= 0) {
        gb[r0] = zs[Z_H + 3] - hang.x / mass_p;
        gb[r0 + 1] = zs[Z_H + 4] - hang.y / mass_p;
        gb[r0 + 2] = zs[Z_H + 5] - hang.z / mass_p;
      }
      r0 = grow(T, HB_KF_FEET_DIST, k, 0);
      if (r0 >= 0) gb[r0] = feet_y;
    }
  }
  // base-quaternion error e = qd^-1 (x) q - (0,0,0,1) = A q - e4 (quaternion.py:64-70); columns of A
  double Aq[4][4];
  {
    const double dx = -pre_bq[0], dy = -pre_bq[1], dz = -pre_bq[2], dw = pre_bq[3];
    // (dw, dv) (x) (bw, bv): v = dw bv + bw dv + dv x bv ; w = dw bw - dv.bv ; column a: b = e_a
    // b = e_x
    Aq[0][0] = dw; Aq[1][0] = dz; Aq[2][0] = -dy; Aq[3][0] = -dx;
    Aq[0][1] = -dz; Aq[1][1] = dw; Aq[2][1] = dx; Aq[3][1] = -dy;
    Aq[0][2] = dy; Aq[1][2] = -dx; Aq[2][2] = dw; Aq[3][2] = -dz;
    Aq[0][3] = dx; Aq[1][3] = dy; Aq[2][3] = dz; Aq[3][3] = dw;
  }
  double eq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) eq[i] = Aq[i][0] * q0 + Aq[i][1] * q1 + Aq[i][2] * q2 + Aq[i][3] * q3 - (i == 3 ? 1.0 : 0.0);
  const double frame_cost = (phi - 3.0) * (phi - 3.0);
  const bool want_terms = (mask & HB_EVAL_COST_TERMS_BIT) != 0;
  if (mask & (HB_EVAL_F | HB_EVAL_GRAD_F | HB_EVAL_COST_TERMS_BIT)) {
    // quaternion-velocity cost (all knots) and, for k >= 1, base quaternion / joint / frame costs
    double c_bqv = 0.0, c_bq = 0.0, c_joint = 0.0, c_frame = 0.0;
    if (lane < 4) {
      const double e = qdv[lane] - pre_bqv;
      c_bqv = T.w_bqv * e * e;
      gbuf[3 + lane] += 2.0 * T.w_bqv * e;
      if (k1) {
        c_bq = T.w_bq * eq[lane] * eq[lane];
        double gq = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          double t = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) t += Aq[i][a] * eq[i];
          if (a == lane) gq = t;
        }
        gbuf[7 + lane] += 2.0 * T.w_bq * gq;
      }
    }
    if (lane < HB_N_JOINTS && k1) {
      const double sd = zs[Z_SD + lane], e = zs[Z_S + lane] - pre_jr;
      if (T.joint_cost_kind == 0) {
        // kinodynamic planner.py:506-520: sumsqr of the broadcast n x n matrix
        const double t = sd + C.wj[lane] * e;
        c_joint = T.w_joint * ((HB_N_JOINTS - 1) * sd * sd + t * t);
        gbuf[11 + lane] += T.w_joint * (2.0 * (HB_N_JOINTS - 1) * sd + 2.0 * t);
        gbuf[34 + lane] += T.w_joint * 2.0 * C.wj[lane] * t;
      } else {
        // pose finder planner.py:584-588: e^T diag(w) e
        c_joint = T.w_joint * (e * C.wj[lane]) * e;
        gbuf[34 + lane] += T.w_joint * 2.0 * C.wj[lane] * e;
      }
    }
    if (lane == 31 && k1) c_frame = T.w_frame * frame_cost;
    if (want_terms) {
      // solution report (hb_eval_cost_terms): one value per named cost expression of this knot
      double* ct = fpart + (b * N + k) * HB_COST_TERMS;
      const double t_frame = warp_sum(c_frame), t_bq = warp_sum(c_bq), t_bqv = warp_sum(c_bqv), t_j = warp_sum(c_joint);
      if (lane == 0) {
        ct[HB_CT_FRAME_QUAT] = t_frame;
        ct[HB_CT_BASE_QUAT] = t_bq;
        ct[HB_CT_BASE_QUAT_VEL] = t_bqv;
        ct[HB_CT_JOINTS] = t_j;
      }
    } else {
      double cost = (c_bqv + c_bq) + c_joint + c_frame;
      cost = warp_sum(cost);
      if (lane == 0 && (mask & HB_EVAL_F)) fpart[(b * N + k) * 2 + 1] = cost;
    }
  }
  __syncwarp();

  HB_PHASE(0, 3);  // g rows, costs
  const bool want_jac = (mask & HB_EVAL_JAC_G) != 0, want_grad = (mask & HB_EVAL_GRAD_F) != 0;
  double* slot = sm + L.slot;
  Dir nodir;
  nodir.mask = 0u;
  // ------------------------------------------------------------------ adjoint sweep, fp64: Jacobian rows
  // (Jacobian-only kernel.  The Hessian kernel gets the same values from forward tangents that its
  // direction lanes need anyway -- see "forward-mode Jacobian" below.)
  if ((want_jac || want_grad) && !(WITH_HESS && HB_FWD_JAC)) {
    Seeds<double> S;
    S.wc = S.hb = S.footF[0] = S.footF[1] = S.footN[0] = S.footN[1] = S.chestN = v3<double>(0.0, 0.0, 0.0);
    if (lane < 24) {
      const int pt = lane / 3, a = lane % 3;
      const D3 frc = v3<double>(a == 0 ? -1.0 : 0.0, a == 1 ? -1.0 : 0.0, a == 2 ? -1.0 : 0.0);
      const D3 trq = cross(ld3(sm + L.arms + 3 * pt), frc);
      if (pt < 4) {
        S.footF[0] = frc;
        S.footN[0] = trq;
      } else {
        S.footF[1] = frc;
        S.footN[1] = trq;
      }
    } else if (lane < 27) {
      const int a = lane - 24;
      S.wc = v3<double>(a == 0 ? -1.0 / M : 0.0, a == 1 ? -1.0 / M : 0.0, a == 2 ? -1.0 / M : 0.0);
    } else if (lane < 30) {
      const int a = lane - 27;
      const double sc = -1.0 / mass_p;
      S.hb = v3<double>(a == 0 ? sc : 0.0, a == 1 ? sc : 0.0, a == 2 ? sc : 0.0);
    } else if (lane == 30) {
      S.footF[0] = fdu;
      S.footN[0] = cross(fdal, fdu);
      S.footF[1] = -fdu;
      S.footN[1] = cross(fdu, fdDelta) - cross(fdar, fdu);
    } else {
      S.chestN = mG;
    }
    JacEmit em;
    em.C = Cp;
    const int2 jkm = *reinterpret_cast<const int2*>(&T.knot_maps[k].jk_base);  // {base, table offset}
    em.map = C.jk_map + jkm.y;
    em.jac = jac + b * T.nnz_j + jkm.x;
    em.gbuf = gbuf;
    em.lane = lane;
    em.k = k;
    em.gscale = k1 ? 2.0 * T.w_frame * (phi - 3.0) : 0.0;
    em.write = want_jac;
    em.stage = (WITH_HESS ? HB_KIN_STAGE : HB_KIN_STAGE_F) ? sm + L.stage : nullptr;
    em.init_bases();
    D3 n0, w0, v0;
    kin_backward<double>(T, sb, zs, nodir, S, xc, xcd, slot, em, n0, w0, v0);
    // base chain rule
    D3 gq[4], uq[4], wq[4];
    const double qraw[4] = {q0, q1, q2, q3};
    quat_maps<double>(qraw, qdv, gq, uq, wq);
    double dq[4], dqd[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      dq[a] = dot(n0, gq[a]) + dot(w0, uq[a]);
      dqd[a] = dot(w0, wq[a]);
    }
    if (lane == 31) {
#pragma unroll
      for (int a = 0; a < 4; ++a) gbuf[7 + a] += em.gscale * dq[a];
    } else if (want_jac) {
      if (lane < 27) {
        const int bq = em.base_q();
#pragma unroll
        for (int a = 0; a < 4; ++a) em.put(bq + a, dq[a]);
      } else if (lane < 30) {
        const int bm = 4 + 648 + 81 + (lane - 27) * 57;
        em.put(bm + 0, v0.x);
        em.put(bm + 1, v0.y);
        em.put(bm + 2, v0.z);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          em.put(bm + 3 + a, dqd[a]);
          em.put(bm + 7 + a, dq[a]);
        }
      }
      if (lane < 4) em.put(lane, 2.0 * qraw[lane]);  // unit quaternion row
    }
    __syncwarp();
    if (em.stage && want_jac) {
      // scatter the staged rows: coalesced map reads, no load->store dependency inside the sweep
      const double* st = sm + L.stage;
      const unsigned* lst = C.jk_list + jkm.y;  // destination-sorted: entry << 16 | slot
      const int n = T.knot_maps[k].jk_cnt;
      for (int eb = lane; eb < n; eb += 256) {  // 8 list loads in flight before the dependent stores
        unsigned w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = (eb + 32 * u) < n ? lst[eb + 32 * u] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (w[u] != 0xffffffffu) em.jac[w[u] & 0xffffu] = st[w[u] >> 16];
      }
      __syncwarp();
    }
    if (want_grad) {
      // gbuf order vb3 qd4 q4 sd23 s23 -> x offsets
      double* gf = grad_f + b * T.n_x + (long)k * T.x_stride;
      for (int i = lane; i < 57; i += 32) {
        int off;
        if (i < 3) off = Z_VB + i;
        else if (i < 7) off = Z_QD + i - 3;
        else if (i < 11) off = Z_Q + i - 7;
        else if (i < 34) off = Z_SD + i - 11;
        else off = Z_S + i - 34;
        if (C.zmap[off] >= 0) gf[C.zmap[off]] = gbuf[i];
      }
      if (lane < 3 && C.zmap[Z_PB + lane] >= 0) gf[C.zmap[Z_PB + lane]] = 0.0;
    }
  }
  if (!WITH_HESS) return;
  if (!with_l && !(HB_FWD_JAC && (want_jac || want_grad))) return;
  // ------------------------------------------------------------------ adjoint sweep, dual: Hessian columns
  {
    const int dirj = lane < 27 ? lane : -1;
    Dir dir;
    dir.mask = 0u;
    dir.alpha = dir.pi = dir.wpi = dir.vpi = dir.u = v3<double>(0.0, 0.0, 0.0);
    // only the primal maps are needed for the directions; the dual maps are rebuilt after the sweep
    // so that they are not live (144 registers) across it
    D3 wq_lane = v3<double>(0.0, 0.0, 0.0);  // d omega_0 / d qd_a for the velocity-direction lanes 3..6
    {
      D3 gq0[4], uq0[4], wq0[4];
      const double qraw[4] = {q0, q1, q2, q3};
      quat_maps<double>(qraw, qdv, gq0, uq0, wq0);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a == lane) {
          dir.alpha = gq0[a];
          dir.u = uq0[a];
        }
        if (a + 3 == lane) wq_lane = wq0[a];
      }
    }
    if (lane < 4) {
      dir.mask = 0xffffffffu;
      dir.wpi = ld3(sb + SB_W);
      dir.vpi = ld3(sb + SB_V);
    } else if (lane < 27) {
      const int l = lane - 4 + 1;
      dir.mask = C.sub_mask[l];
      const double* bl = sb + l * SB_STRIDE;
      dir.alpha = ld3(bl + SB_AX);
      dir.pi = ld3(bl + SB_O);
      dir.wpi = ld3(bl + SB_W);
      dir.vpi = ld3(bl + SB_V);
    }
    // ---- composite (sub-tree) moments, lane = body, accumulated leaf-to-root in the still unused stage:
    //   M, P = sum m c, Pd = sum m c_dot, H = sum (m c x c_dot + I w), J = sum (I + m (|c|^2 1 - c c^T))
    // A direction rotates one sub-tree rigidly about its pivot, so the tangents of the total P, P_dot
    // and angular momentum along it are closed-form in that sub-tree's moments (composite-rigid-body
    // argument): no lane loops over the bodies for them (that loop was 0.29 ms of 1.81 ms).
    enum { CM_M = 0, CM_P = 1, CM_PD = 4, CM_H = 7, CM_J = 10, CM_STRIDE = 17 };
    double* comp = sm + L.stage;
    if (lane < nb) {
      const double* bl = sb + lane * SB_STRIDE;
      const double m = my_mass;
      const D3 c = ld3(bl + SB_O) + ld3(bl + SB_D);
      const D3 cd = ld3(bl + SB_V) + cross(ld3(bl + SB_W), ld3(bl + SB_D));
      const double* I = bl + SB_I;
      double* cm = comp + lane * CM_STRIDE;
      cm[CM_M] = m;
      st3(cm + CM_P, scale(m, c));
      st3(cm + CM_PD, scale(m, cd));
      st3(cm + CM_H, cross(scale(m, c), cd) + ld3(bl + SB_L));
      cm[CM_J + 0] = I[0] + m * (c.y * c.y + c.z * c.z);
      cm[CM_J + 1] = I[1] - m * c.x * c.y;
      cm[CM_J + 2] = I[2] - m * c.x * c.z;
      cm[CM_J + 3] = I[3] + m * (c.x * c.x + c.z * c.z);
      cm[CM_J + 4] = I[4] - m * c.y * c.z;
      cm[CM_J + 5] = I[5] + m * (c.x * c.x + c.y * c.y);
    }
    __syncwarp();
    {
      // lane = body: sum the rows of its sub-tree (every lane reads the same row: broadcast loads, no
      // barriers; the leaf-to-root read-modify-write version of this pass took 12 % of the kernel)
      double acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.0;
      for (int l = 0; l < nb; ++l) {
        const double in = ((my_submask >> l) & 1u) ? 1.0 : 0.0;
        const double* cl = comp + l * CM_STRIDE;
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(in, cl[i], acc[i]);
      }
      __syncwarp();
      if (lane < nb) {
        double* cm = comp + lane * CM_STRIDE;
#pragma unroll
        for (int i = 0; i < 16; ++i) cm[i] = acc[i];
      }
      __syncwarp();
    }
    HB_PHASE(0, 4);  // composite moments
    // tangents of P, P_dot, h (about the base origin) along this lane's (q, s) direction
    D3 tP, tPd, th;
    {
      const double* cm = comp + (lane >= 4 && lane < 27 ? lane - 3 : 0) * CM_STRIDE;
      const double Ms = cm[CM_M];
      const D3 P = ld3(cm + CM_P), Pds = ld3(cm + CM_PD), Hs = ld3(cm + CM_H);
      const D3 a = dir.alpha, beta = cross(dir.alpha, dir.wpi);
      const D3 Pr = P - scale(Ms, dir.pi);
      tP = cross(a, Pr);
      tPd = cross(a, Pds - scale(Ms, dir.vpi)) - cross(beta, Pr) + cross(dir.u, P);
      th = cross(a, Hs) - cross(cross(a, dir.pi), Pds) - cross(P, cross(a, dir.vpi)) - symmul(cm + CM_J, beta) +
           cross(P, cross(beta, dir.pi)) + symmul(cm + CM_J, dir.u);
    }
    // velocity directions (lanes 0..29 = vb, qd, sd): w_l += in_l A, v_l += in_l A x (o_l - pivot) + B;
    // momentum is linear in the velocities, so its column is again closed-form in the moments
    D3 vel_dh = v3<double>(0.0, 0.0, 0.0);
    D3 vel_pv = vel_dh;  // d Pdot / d (velocity variable of this lane), for the cross terms of the Hessian
    if (lane < 30) {
      D3 A = v3<double>(0.0, 0.0, 0.0), B = A, piv = A;
      int lv = 0;
      if (lane < 3) B = v3<double>(lane == 0 ? 1.0 : 0.0, lane == 1 ? 1.0 : 0.0, lane == 2 ? 1.0 : 0.0);
      else if (lane < 7) A = wq_lane;
      else {
        lv = lane - 6;
        A = ld3(sb + lv * SB_STRIDE + SB_AX);
        piv = ld3(sb + lv * SB_STRIDE + SB_O);
      }
      const double* cm = comp + lv * CM_STRIDE;
      const double Ms = cm[CM_M];
      const D3 P = ld3(cm + CM_P);
      const D3 pv = cross(A, P - scale(Ms, piv)) + scale(Ms, B);
      const D3 hv = symmul(cm + CM_J, A) - cross(P, cross(A, piv)) + cross(P, B);
      vel_dh = scale(-1.0 / mass_p, hv - scale(1.0 / M, cross(Pm, pv)));
      vel_pv = pv;
    }
    __syncwarp();  // the stage is reused for the Jacobian columns below
    const V3<Dual> xcD = lift<Dual>(xc, scale(1.0 / M, tP));
    const V3<Dual> xdD = lift<Dual>(xcd, scale(1.0 / M, tPd));
    const double inL = ((dir.mask >> T.foot_body[0]) & 1u) ? 1.0 : 0.0;
    const double inR = ((dir.mask >> T.foot_body[1]) & 1u) ? 1.0 : 0.0;
    const double inC = ((dir.mask >> T.chest_body) & 1u) ? 1.0 : 0.0;
    // tangents of the feet distance u.(pL - pR) and of trace(G) along this direction
    double ty_lane, tphi_lane;
    {
      const D3 a = dir.alpha;
      const D3 tL = scale(inL, cross(a, (ld3(sb + T.foot_body[0] * SB_STRIDE + SB_O) - dir.pi) + fdal));
      const D3 tR = scale(inR, cross(a, (ld3(sb + T.foot_body[1] * SB_STRIDE + SB_O) - dir.pi) + fdar));
      ty_lane = dot(scale(inR, cross(a, fdu)), fdDelta) + dot(fdu, tL - tR);
      const D3 ac = scale(inC, a);
      tphi_lane = cross(ac, v3<double>(G[0], G[3], G[6])).x + cross(ac, v3<double>(G[1], G[4], G[7])).y +
                  cross(ac, v3<double>(G[2], G[5], G[8])).z;
    }
    HB_PHASE(0, 5);  // direction tangents of P, Pdot, h, feet distance, trace
#if HB_FWD_JAC
    // ---------------------------------------------------------------- forward-mode Jacobian
    // The direction lanes already hold the tangent of every link state along q_a / s_j, so column j of
    // the kinematic rows is a few products here; the velocity columns of the momentum rows (the map is
    // linear in the velocities) take one more pass over the bodies.  This replaces the row-per-lane
    // fp64 adjoint sweep of the Jacobian-only kernel (2.87 -> see DESIGN.md).
    if (want_jac || want_grad) {
      double* stg = sm + L.stage;
      const int bm = 4 + 648 + 81;  // momentum rows inside JK (kino_layout.py::_enumerate_jk)
      if (want_jac) {
        if (lane < 27) {
#pragma unroll
          for (int pt = 0; pt < 8; ++pt) {
            const int f = pt >> 2;
            const double in = f == 0 ? inL : inR;
            const D3 r = (ld3(sb + T.foot_body[f] * SB_STRIDE + SB_O) - dir.pi) + ld3(sm + L.arms + 3 * pt);
            const D3 t = scale(-in, cross(dir.alpha, r));
            stg[4 + (3 * pt) * 27 + lane] = t.x;
            stg[4 + (3 * pt + 1) * 27 + lane] = t.y;
            stg[4 + (3 * pt + 2) * 27 + lane] = t.z;
          }
          const D3 tc = scale(-1.0 / M, tP);
          stg[4 + 648 + lane] = tc.x;
          stg[4 + 648 + 27 + lane] = tc.y;
          stg[4 + 648 + 54 + lane] = tc.z;
          const D3 dh = scale(-1.0 / mass_p, th - scale(1.0 / M, cross(tP, Pd) + cross(Pm, tPd)));
          const int cq = lane < 4 ? 7 + lane : 34 + lane - 4;
          stg[bm + cq] = dh.x;
          stg[bm + 57 + cq] = dh.y;
          stg[bm + 114 + cq] = dh.z;
          if (lane < 4) stg[lane] = 2.0 * zs[Z_Q + lane];  // unit quaternion row
          else stg[bm + 171 + lane - 4] = ty_lane;          // feet distance row
        }
        if (lane < 30) {
          const int cv = lane < 7 ? lane : lane + 4;  // velocity columns: vb 0..2, qd 3..6, sd 11..33
          stg[bm + cv] = vel_dh.x;
          stg[bm + 57 + cv] = vel_dh.y;
          stg[bm + 114 + cv] = vel_dh.z;
        }
      }
      // frame-orientation cost: d/dz w (phi - 3)^2 = 2 w (phi - 3) dphi/dz
      if (want_grad && lane < 27 && k1)
        gbuf[lane < 4 ? 7 + lane : 34 + lane - 4] += 2.0 * T.w_frame * (phi - 3.0) * tphi_lane;
      HB_PHASE(0, 6);  // Jacobian columns staged
      __syncwarp();
      if (want_jac) {
        const int4 jkm = *reinterpret_cast<const int4*>(&T.knot_maps[k].jk_base);  // {base, offset, count, -}
        double* jb = jac + b * T.nnz_j + jkm.x;
#if HB_KIN_JAC_LIST
        const unsigned* lst = C.jk_list + jkm.y;  // destination-sorted: entry << 16 | slot
        const int n = jkm.z;
        // KIN_SCAT list loads in flight before the dependent stores: <= 927 entries in two trips
        for (int eb = lane; eb < n; eb += 32 * KIN_SCAT) {
          unsigned w[KIN_SCAT];
#pragma unroll
          for (int u = 0; u < KIN_SCAT; ++u) w[u] = (eb + 32 * u) < n ? lst[eb + 32 * u] : 0xffffffffu;
#pragma unroll
          for (int u = 0; u < KIN_SCAT; ++u)
            if (w[u] != 0xffffffffu) jb[w[u] & 0xffffu] = stg[w[u] >> 16];
        }
#else
        const int* jmap = C.jk_map + jkm.y;
        const int n = T.n_jk;
        for (int eb = lane; eb < n; eb += 32 * KIN_SCAT) {  // entry order: 927 entries in two trips
          int sl[KIN_SCAT];
#pragma unroll
          for (int u = 0; u < KIN_SCAT; ++u) sl[u] = (eb + 32 * u) < n ? jmap[eb + 32 * u] : -1;
#pragma unroll
          for (int u = 0; u < KIN_SCAT; ++u)
            if (sl[u] >= 0) jb[sl[u]] = stg[eb + 32 * u];
        }
#endif
      }
      if (want_grad) {
        // gbuf order vb3 qd4 q4 sd23 s23 -> x offsets
        double* gf = grad_f + b * T.n_x + (long)k * T.x_stride;
        for (int i = lane; i < 57; i += 32) {
          int off;
          if (i < 3) off = Z_VB + i;
          else if (i < 7) off = Z_QD + i - 3;
          else if (i < 11) off = Z_Q + i - 7;
          else if (i < 34) off = Z_SD + i - 11;
          else off = Z_S + i - 34;
          if (C.zmap[off] >= 0) gf[C.zmap[off]] = gbuf[i];
        }
        if (lane < 3 && C.zmap[Z_PB + lane] >= 0) gf[C.zmap[Z_PB + lane]] = 0.0;
      }
      __syncwarp();  // the Hessian columns reuse the stage
    }
#endif
    HB_PHASE(0, 7);  // Jacobian scatter, grad_f
    if (!with_l) return;  // Jacobian / gradient only: done
    Seeds<Dual> S;
    S.footF[0] = S.footF[1] = S.footN[0] = S.footN[1] = vzero<Dual>();
    {
      // multipliers from lamk (loaded at the top of the kernel; rows that do not exist hold 0)
      S.wc = scale(-1.0 / M, ld3(lamk + 24));
      S.hb = scale(-1.0 / mass_p, ld3(lamk + 27));
      // stage I^w hbar (same for every lane of the dual sweep)
      if (lane < nb) st3(sb + lane * SB_STRIDE + SB_IH, symmul(sb + lane * SB_STRIDE + SB_I, S.hb));
      __syncwarp();
    }
#pragma unroll
    for (int pt = 0; pt < 8; ++pt) {
      const int f = pt >> 2;
      const D3 frc = v3<double>(-lamk[3 * pt], -lamk[3 * pt + 1], -lamk[3 * pt + 2]);
      const D3 a = ld3(sm + L.arms + 3 * pt);
      const double in = f == 0 ? inL : inR;
      const V3<Dual> aD = lift<Dual>(a, scale(in, cross(dir.alpha, a)));
      S.footF[f] = S.footF[f] + lift<Dual>(frc, v3<double>(0.0, 0.0, 0.0));
      S.footN[f] = S.footN[f] + cross(aD, frc);
    }
    {
      const double kd = lamk[30];
      const St<Dual> sL = load_state(sb, T.foot_body[0], dir, Dual());
      const St<Dual> sR = load_state(sb, T.foot_body[1], dir, Dual());
      const V3<Dual> uD = lift<Dual>(fdu, scale(inR, cross(dir.alpha, fdu)));
      const V3<Dual> alD = lift<Dual>(fdal, scale(inL, cross(dir.alpha, fdal)));
      const V3<Dual> arD = lift<Dual>(fdar, scale(inR, cross(dir.alpha, fdar)));
      const V3<Dual> dD = (sL.o + alD) - (sR.o + arD);
      const V3<Dual> ku = scale(kd, uD);
      S.footF[0] = S.footF[0] + ku;
      S.footN[0] = S.footN[0] + cross(alD, ku);
      S.footF[1] = S.footF[1] - ku;
      S.footN[1] = S.footN[1] + scale(kd, cross(uD, dD)) - cross(arD, ku);
    }
    {
      // kappa = 2 sigma w (phi - 3) with its tangent; m(G) with its tangent (columns of G rotate with alpha)
      const D3 a = scale(inC, dir.alpha);
      double tG[9];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const D3 col = v3<double>(G[j], G[3 + j], G[6 + j]);
        const D3 t = cross(a, col);
        tG[j] = t.x;
        tG[3 + j] = t.y;
        tG[6 + j] = t.z;
      }
      const double tphi = tG[0] + tG[4] + tG[8];
      const double cw = k1 ? 2.0 * sg * T.w_frame : 0.0;
      const Dual kappa = mkdual(cw * (phi - 3.0), cw * tphi);
      const V3<Dual> mGD = lift<Dual>(mG, v3<double>(tG[5] - tG[7], tG[6] - tG[2], tG[1] - tG[3]));
      S.chestN = scale(kappa, mGD);
    }
    HessEmit em;
    em.map = C.hk_map + T.knot_maps[k].hk_off;
    em.hess = hess + b * T.nnz_h + T.knot_maps[k].hk_base;
    em.stage = HB_KIN_STAGE ? sm + L.stage : nullptr;
    em.dirj = dirj;
    em.add_sd = em.add_s = 0.0;
    if (lane >= 4 && lane < 27 && k1) {
      const double wjl = C.wj[lane - 4];
      em.add_sd = T.joint_cost_kind == 0 ? sg * T.w_joint * 2.0 * wjl : 0.0;
      em.add_s = T.joint_cost_kind == 0 ? sg * T.w_joint * 2.0 * wjl * wjl : sg * T.w_joint * 2.0 * wjl;
    }
    V3<Dual> n0, w0, v0;
    const double lu_pre = lamk[31];  // read now: the slots of the packed sweep alias gbuf + lamk + slot
#ifdef HB_DUAL_SWEEP
    // reference implementation: the whole adjoint sweep in dual arithmetic on every lane
    kin_backward<Dual>(T, sb, zs, dir, S, xcD, xdD, slot, em, n0, w0, v0);
#else
    {
      SeedsP SP;
      SeedsT ST;
      SP.wc = S.wc;
      SP.hb = S.hb;
#pragma unroll
      for (int f = 0; f < 2; ++f) {
        SP.footF[f] = primal_of(S.footF[f]);
        SP.footN[f] = primal_of(S.footN[f]);
        ST.footF[f] = tangent_of(S.footF[f]);
        ST.footN[f] = tangent_of(S.footN[f]);
      }
      SP.chestN = primal_of(S.chestN);
      ST.chestN = tangent_of(S.chestN);
      __syncwarp();
      HB_PHASE(0, 8);  // seeds
      primal_adjoint_pass(T, sb, zs, SP, xc, xcd, my_depth, my_rank, my_parent, my_mass);
      HB_PHASE(0, 9);  // primal adjoint pass
      D3 tn0, tw0, tv0;
#if HB_SWEEP_PACKED
      {
        // Hessian of -hb.(P x Pdot)/M (see kin_tangent_sweep_packed) for every (direction, row), which also
        // initialises the stage; rows: vb3 qd4 sd23 (velocity lane = row) q4 s23 (direction lane = row - 30)
        double* stg = sm + L.stage;
        double* pslot = sm + L.gbuf;  // gbuf + lamk + slot: 474 contiguous doubles, all dead by now
        // tangents of P (rows q, s) and of Pdot (all rows) parked in the slot area for the cross terms;
        // quaternion direction data and the masses in the part of z the kinematics no longer reads
        double* xt = pslot;           // [57][6]: dP_r (zero for velocity rows), dPdot_r
        double* qdat = zs;            // [4][6]: g_a, u_a
        double* massv = zs + 24;      // [nb]
        if (lane < 27) {
          st3(xt + (30 + lane) * 6, tP);
          st3(xt + (30 + lane) * 6 + 3, tPd);
        }
        if (lane < 30) {
          st3(xt + lane * 6, v3<double>(0.0, 0.0, 0.0));
          st3(xt + lane * 6 + 3, vel_pv);
        }
        if (lane < 4) {
          st3(qdat + 6 * lane, dir.alpha);
          st3(qdat + 6 * lane + 3, dir.u);
        }
        if (lane < nb) massv[lane] = my_mass;
        __syncwarp();
        if (lane < 27) {
          const D3 hbv = S.hb;
          const D3 a1 = cross(tPd, hbv), a2 = cross(hbv, tP);  // hb.(dP_r x dPdot_d) = dP_r.a1, hb.(dP_d x dPdot_r) = dPdot_r.a2
          double* col = stg + lane * 57;
          const double sM = -1.0 / M;
          // velocity rows (vb, qd, sd: 0..29) have dP_r = 0: half the loads and products
#pragma unroll 5
          for (int r = 0; r < 30; ++r) col[r] = sM * dot(ld3(xt + 6 * r + 3), a2);
#pragma unroll 3
          for (int r = 30; r < 57; ++r) col[r] = sM * (dot(ld3(xt + 6 * r), a1) + dot(ld3(xt + 6 * r + 3), a2));
          if (lane >= 4) {  // joint regularisation on (sd_j, s_j), (s_j, s_j)
            col[7 + lane - 4] += em.add_sd;
            col[34 + lane - 4] += em.add_s;
          }
        }
        __syncwarp();
        HB_PHASE(0, 10);  // cross terms
        em.acc = true;
        kin_tangent_sweep_packed(T, sb, zs, qdat, massv, ST, S.hb, pslot, stg);
        HB_PHASE(0, 11);  // packed tangent sweep
        const double* rs = pslot + (lane < 27 ? T.root_slot[lane] : 0) * 12;
        tn0 = ld3(rs);
        tw0 = ld3(rs + 6);
        tv0 = ld3(rs + 9);
      }
#else
      kin_tangent_sweep(T, sb, zs, dir, S.hb, ST, tangent_of(xcD), tangent_of(xdD), slot, em, tn0, tw0, tv0);
#endif
      n0 = lift<Dual>(ld3(sb + SB_ACC), tn0);
      w0 = lift<Dual>(ld3(sb + SB_ACC + 6), tw0);
      v0 = lift<Dual>(ld3(sb + SB_ACC + 9), tv0);
    }
#endif
    if (dirj >= 0) {
      em.put(0, v0.x.d);
      em.put(1, v0.y.d);
      em.put(2, v0.z.d);
      const double lu = lu_pre;
      // dual quaternion maps, rebuilt from shared memory after the sweep (see above)
      Dual qD[4];
      double qdv2[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        qD[a] = mkdual(zs[Z_Q + a], a == lane ? 1.0 : 0.0);
        qdv2[a] = zs[Z_QD + a];
      }
      V3<Dual> gq[4], uq[4], wq[4];
      quat_maps<Dual>(qD, qdv2, gq, uq, wq);
      // Hessian of w_bq |qd^-1 (x) q - 1|^2 is 2 w_bq A^T A = 2 w_bq |qd|^2 I (left-multiplication matrix)
      const double nqd = rbq[0] * rbq[0] + rbq[1] * rbq[1] + rbq[2] * rbq[2] +
                         rbq[3] * rbq[3];  // re-read (L1/L2 hit by now) rather than kept live across the sweep
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const Dual dq = dot(n0, gq[a]) + dot(w0, uq[a]);
        const Dual dqd = dot(w0, wq[a]);
        const double extra = (a == lane && k1) ? 2.0 * sg * T.w_bq * nqd + 2.0 * lu : 0.0;
        em.put(3 + a, dqd.d);
        em.put(30 + a, dq.d + extra);
      }
    }
    HB_PHASE(0, 12);  // root chain rule (dual quaternion maps)
    __syncwarp();
    if (em.stage) {
      // scatter the staged columns (27 directions x 57 rows) with coalesced map reads
      const double* st = sm + L.stage;
#if HB_KIN_HESS_LIST
      const int2 hkl = *reinterpret_cast<const int2*>(&T.knot_maps[k].hk_off);  // {offset, count}
      const unsigned* lst = C.hk_list + hkl.x;  // destination-sorted: entry << 16 | slot; mirrored pairs are absent
      const int n = hkl.y;
      for (int eb = lane; eb < n; eb += 32 * KIN_SCAT) {
        unsigned w[KIN_SCAT];
#pragma unroll
        for (int u = 0; u < KIN_SCAT; ++u) w[u] = (eb + 32 * u) < n ? lst[eb + 32 * u] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < KIN_SCAT; ++u)
          if (w[u] != 0xffffffffu) em.hess[w[u] & 0xffffu] = st[w[u] >> 16];
      }
#else
      for (int eb = lane; eb < HSTAGE; eb += 32 * KIN_SCAT) {  // entry order: 1539 entries in three trips
        int sl[KIN_SCAT];
#pragma unroll
        for (int u = 0; u < KIN_SCAT; ++u) sl[u] = (eb + 32 * u) < HSTAGE ? em.map[eb + 32 * u] : -1;
#pragma unroll
        for (int u = 0; u < KIN_SCAT; ++u)
          if (sl[u] >= 0) em.hess[sl[u]] = st[eb + 32 * u];
      }
#endif
    }
    // velocity-diagonal entries (HK2): quaternion-velocity cost, joint regularisation
    if (lane < 27) {
      const int2 hk2m = *reinterpret_cast<const int2*>(&T.knot_maps[k].hk2_base);
      const int slot2 = C.hk2_map[hk2m.y + lane];
      if (slot2 >= 0)
        (hess + b * T.nnz_h + hk2m.x)[slot2] = lane < 4 ? 2.0 * sg * T.w_bqv : sg * T.w_joint * 2.0 * HB_N_JOINTS;
    }
    HB_PHASE(0, 13);  // Hessian scatter
  }
}

template __global__ void kino_kin_kernel<true>(const __grid_constant__ KinTopo, const KinoConst*, unsigned, const double*,
                                               const double*, long, const double*, const double*, double*, double*,
                                               double*, double*, double*, long);
template __global__ void kino_kin_kernel<false>(const __grid_constant__ KinTopo, const KinoConst*, unsigned,
                                                const double*, const double*, long, const double*, const double*,
                                                double*, double*, double*, double*, double*, long);

}  // namespace hb

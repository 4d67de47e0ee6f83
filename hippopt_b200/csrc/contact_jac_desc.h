// contact_jac_desc.h -- value descriptors of the contact kernel's local Jacobian entries (kino_layout.py::_enumerate_jc).
//
// Every entry of that list is `coefficient x source`: the coefficient is one of a dozen per-warp scalars (+-1, +-dt/2,
// the robot mass, ...), the source is a double of the warp's shared-memory block -- a variable of the knot, or one of the
// few derived quantities the kernel stages behind them (tanh terms, p - com, the force sum).  The host turns the
// list into (destination slot, descriptor) pairs SORTED BY DESTINATION, so that the kernel's scatter is a coalesced
// stream of stores with no per-entry branching (the previous version walked the list in entry order: every store
// of a warp hit 32 different sectors and the per-entry decision tree diverged).
#pragma once

namespace hb {

// sources: index into the warp's block, [0, 189) = the knot's variables (zs), then the extras
enum {
  JX_ONE = 190,            // 1.0
  JX_TAU = 191,            // [8]  tanh(k_t p_z) per contact point
  JX_DTAU_U = 199,         // [16] dtau * u_x, dtau * u_y per point
  JX_KF = 215,             // [8]  k_bs f_z + fdot_z
  JX_KP = 223,             // [8]  k_bs p_z + v_z
  JX_PC = 231,             // [24] p - com per point
  JX_FSUM = 255,           // [3]  sum of the contact forces
  JX_COEF = 258,           // [16] coefficient table, JC_* below
  JX_END = 274
};
// coefficients
enum {
  JC_ONE = 0, JC_MONE, JC_HDT, JC_MHDT, JC_MASS, JC_QUARTER, JC_MQUARTER, JC_PERIODIC /* +1 at knot 0, -1 at the last */,
  JC_MTWO, JC_TWO_MU2, JC_COUNT
};
enum { JC_SKIP = -1 };  // entry written elsewhere (smooth-terrain point rows)

constexpr int jc_pack(int coef, int src) { return (coef << 16) | src; }

// Descriptor of local entry e; terrain: 0 planar, 1 smooth steps.  Offsets follow _enumerate_jc: C1 81 x 4 linear
// dynamics, C2 93 initial conditions, C3 81 final + 84 periodicity, C4 2 x 132 centroidal momentum, C5 per-point
// rows (29 or 58 per point), C6 constant robot rows.
inline int contact_jac_descriptor(int e, int terrain) {
  auto eps3 = [](int a, int b) { return ((b - a + 3) % 3 == 1) ? 1 : -1; };
  auto pair_bc = [](int pr, int& a, int& b, int& c) {  // (0,1),(0,2),(1,0),(1,2),(2,0),(2,1)
    a = pr >> 1;
    b = a == 0 ? (pr & 1) + 1 : (a == 1 ? ((pr & 1) ? 2 : 0) : (pr & 1));
    c = 3 - a - b;
  };
  if (e < 324) {
    const int t = e & 3;
    return jc_pack(t == 0 ? JC_ONE : (t == 2 ? JC_MONE : JC_MHDT), JX_ONE);
  }
  if (e < 417) return jc_pack((e - 324) < 87 ? JC_ONE : JC_MONE, JX_ONE);
  if (e < 498) return jc_pack(JC_ONE, JX_ONE);
  if (e < 582) return jc_pack(JC_PERIODIC, JX_ONE);
  if (e < 846) {
    const int side = (e - 582) / 132, q = (e - 582) % 132;
    if (q < 6) return jc_pack(side == 0 ? JC_ONE : JC_MONE, JX_ONE);
    if (q < 30) return jc_pack(JC_MHDT, JX_ONE);
    int a, b, c;
    if (q < 78) {  // -dt/2 eps f_c
      pair_bc((q - 30) % 6, a, b, c);
      return jc_pack(eps3(a, b) > 0 ? JC_MHDT : JC_HDT, 15 * ((q - 30) / 6) + 9 + c);
    }
    if (q < 126) {  // dt/2 eps (p - com)_c
      pair_bc((q - 78) % 6, a, b, c);
      return jc_pack(eps3(a, b) > 0 ? JC_HDT : JC_MHDT, JX_PC + 3 * ((q - 78) / 6) + c);
    }
    pair_bc(q - 126, a, b, c);
    return jc_pack(eps3(a, b) > 0 ? JC_HDT : JC_MHDT, JX_FSUM + c);
  }
  const int n_pt = terrain == 0 ? 29 : 58;
  const int base6 = 846 + 8 * n_pt;
  if (e < base6) {
    if (terrain != 0) return JC_SKIP;
    const int pt = (e - 846) / 29, u = (e - 846) % 29, o = 15 * pt;
    switch (u) {
      case 3: case 4: return jc_pack(JC_MONE, JX_TAU + pt);
      case 5: return jc_pack(JC_MONE, JX_ONE);
      case 6: return jc_pack(JC_MONE, JX_DTAU_U + 2 * pt);
      case 7: return jc_pack(JC_MONE, JX_DTAU_U + 2 * pt + 1);
      case 8: return jc_pack(JC_MONE, o + 9 + 2);    // -f_z
      case 9: return jc_pack(JC_MONE, o + 6 + 2);    // -p_z
      case 10: return jc_pack(JC_MONE, JX_KF + pt);
      case 11: return jc_pack(JC_MONE, JX_KP + pt);
      case 14: return jc_pack(JC_MTWO, o + 9);
      case 15: return jc_pack(JC_MTWO, o + 9 + 1);
      case 16: return jc_pack(JC_TWO_MU2, o + 9 + 2);
      case 20: case 21: case 22: return jc_pack(JC_MASS, JX_ONE);
      case 26: case 27: case 28: return jc_pack(JC_MONE, JX_ONE);
      default: return jc_pack(JC_ONE, JX_ONE);
    }
  }
  const int nch = terrain == 0 ? 1 : 3;
  const int r = e - base6;
  if (terrain != 0 && r >= 12 && r < 12 + nch) return JC_SKIP;  // smooth CoM-height row: terrain block
  if (r < 3) return jc_pack(JC_ONE, JX_ONE);
  if (r < 6) return jc_pack(JC_MONE, JX_ONE);
  if (r < 9) return jc_pack(JC_ONE, JX_ONE);
  if (r < 12) return jc_pack(JC_MASS, JX_ONE);
  if (r < 58 + nch) return jc_pack(JC_ONE, JX_ONE);
  return jc_pack((r - 58 - nch) < 4 ? JC_QUARTER : JC_MQUARTER, JX_ONE);
}

}  // namespace hb

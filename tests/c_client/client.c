/* A caller of libhippopt_b200.so that knows nothing of Python: loads a problem file written by hb_save, asks the
 * library for dimensions, patterns and bounds, evaluates one instance through the host-pointer entry point and
 * prints what it got (the GPU test compares with the Python evaluator's numbers).
 *   usage: client <problem.bin> <vectors.bin>     vectors.bin: doubles x[n_x] p[n_p] lam[m] sigma[1]            */
#include <stdio.h>
#include <stdlib.h>

#include "hippopt_b200.h"

#define CHECK(call)                                                      \
  do {                                                                   \
    int rc_ = (call);                                                    \
    if (rc_ != HB_OK) {                                                  \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hb_last_error()); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  hb_handle h = NULL;
  CHECK(hb_load(argv[1], &h));
  int64_t n_x, n_p, m, nnz_j, nnz_h;
  CHECK(hb_dims(h, &n_x, &n_p, &m, &nnz_j, &nnz_h));
  int64_t* jc = malloc(sizeof(int64_t) * (n_x + 1));
  int64_t* jr = malloc(sizeof(int64_t) * nnz_j);
  int64_t* hc = malloc(sizeof(int64_t) * (n_x + 1));
  int64_t* hr = malloc(sizeof(int64_t) * nnz_h);
  CHECK(hb_pattern_jac(h, jc, jr));
  CHECK(hb_pattern_hess(h, hc, hr));
  double *x, *p, *lam, *sigma, *f, *grad, *g, *jac, *hess;
  CHECK(hb_host_alloc((void**)&x, sizeof(double) * n_x));
  CHECK(hb_host_alloc((void**)&p, sizeof(double) * n_p));
  CHECK(hb_host_alloc((void**)&lam, sizeof(double) * m));
  CHECK(hb_host_alloc((void**)&sigma, sizeof(double)));
  CHECK(hb_host_alloc((void**)&f, sizeof(double)));
  CHECK(hb_host_alloc((void**)&grad, sizeof(double) * n_x));
  CHECK(hb_host_alloc((void**)&g, sizeof(double) * m));
  CHECK(hb_host_alloc((void**)&jac, sizeof(double) * nnz_j));
  CHECK(hb_host_alloc((void**)&hess, sizeof(double) * nnz_h));
  FILE* fv = fopen(argv[2], "rb");
  if (!fv || fread(x, sizeof(double), n_x, fv) != (size_t)n_x || fread(p, sizeof(double), n_p, fv) != (size_t)n_p ||
      fread(lam, sizeof(double), m, fv) != (size_t)m || fread(sigma, sizeof(double), 1, fv) != 1) {
    fprintf(stderr, "cannot read %s\n", argv[2]);
    return 1;
  }
  fclose(fv);
  double* lbg = malloc(sizeof(double) * m);
  double* ubg = malloc(sizeof(double) * m);
  CHECK(hb_bounds(h, p, lbg, ubg));
  CHECK(hb_host_set_parameters(h, p, 0, 1));
  CHECK(hb_eval_host(h, HB_EVAL_F | HB_EVAL_GRAD_F | HB_EVAL_G | HB_EVAL_JAC_G | HB_EVAL_HESS_L, x, lam, sigma, f, grad, g,
                     jac, hess, 1));
  double sg = 0, sj = 0, sh = 0, sgr = 0, slb = 0, sub = 0;
  int64_t n_eq = 0;
  for (int64_t i = 0; i < m; ++i) {
    sg += g[i] * (double)(i % 7 + 1);
    if (lbg[i] == ubg[i]) ++n_eq;
    if (lbg[i] > -1e300) slb += lbg[i];
    if (ubg[i] < 1e300) sub += ubg[i];
  }
  for (int64_t i = 0; i < nnz_j; ++i) sj += jac[i] * (double)(jr[i] % 5 + 1);
  for (int64_t i = 0; i < nnz_h; ++i) sh += hess[i] * (double)(hr[i] % 3 + 1);
  for (int64_t i = 0; i < n_x; ++i) sgr += grad[i] * (double)(i % 11 + 1);
  printf("dims %lld %lld %lld %lld %lld\n", (long long)n_x, (long long)n_p, (long long)m, (long long)nnz_j, (long long)nnz_h);
  printf("pattern %lld %lld %lld %lld\n", (long long)jc[n_x], (long long)jr[nnz_j - 1], (long long)hc[n_x], (long long)hr[nnz_h - 1]);
  printf("bounds %lld %.17g %.17g\n", (long long)n_eq, slb, sub);
  printf("values %.17g %.17g %.17g %.17g %.17g\n", f[0], sgr, sg, sj, sh);
  CHECK(hb_destroy(h));
  return 0;
}

/* oracle/sxvm.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A C restatement of the "SX virtual machine" CasADi uses to evaluate expanded (SX) functions [ext:
 * casadi is not vendored under /root/reference; SXFunction::eval is a switch-loop over an instruction
 * tape with a liveness-allocated work vector], extended with forward-mode directional sweeps so that
 * the same tape yields Jacobian / Hessian values (what nlp_jac_g / nlp_hess_l deliver after CasADi's
 * own AD).  The tape, its work-vector slots and the "has a tangent" flags are produced by oracle/cvm.py
 * from the graphs of oracle/sx.py.  Batched over instances with OpenMP: this is the CPU baseline the
 * benchmark reports next to the GPU numbers (bench.py cpu_baseline / --impl reference).
 *
 * build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libsxvm.so oracle/sxvm.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum {
  OP_CONST = 0, OP_SYM, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_SQ, OP_SQRT, OP_SIN, OP_COS, OP_TANH, OP_EXP,
  OP_POWC, OP_FABS
};

/* One instance: values in w[slot], tangents in t[slot * D .. slot * D + D). */
static void run_one(int n_instr, const int *op, const int *dst, const int *a, const int *b, const double *cval,
                    const unsigned char *has_t, int D, const double *x, const int *in_seed, double *w, double *t) {
  for (int i = 0; i < n_instr; ++i) {
    const int o = op[i], d = dst[i];
    double *td = t + (size_t)d * D;
    if (o == OP_CONST) {
      w[d] = cval[i];
      continue;
    }
    if (o == OP_SYM) {
      w[d] = x[a[i]];
      if (has_t[i]) {
        for (int k = 0; k < D; ++k) td[k] = 0.0;
        td[in_seed[a[i]]] = 1.0;
      }
      continue;
    }
    const double va = w[a[i]];
    const double *ta = t + (size_t)a[i] * D;
    const int ha = has_t[i] & 2, hb = has_t[i] & 4;
    double vb = 0.0;
    const double *tb = ta;
    if (b[i] >= 0) {
      vb = w[b[i]];
      tb = t + (size_t)b[i] * D;
    }
    double v, da = 0.0, db = 0.0;
    switch (o) {
      case OP_ADD: v = va + vb; da = 1.0; db = 1.0; break;
      case OP_SUB: v = va - vb; da = 1.0; db = -1.0; break;
      case OP_MUL: v = va * vb; da = vb; db = va; break;
      case OP_DIV: v = va / vb; da = 1.0 / vb; db = -v / vb; break;
      case OP_NEG: v = -va; da = -1.0; break;
      case OP_SQ: v = va * va; da = 2.0 * va; break;
      case OP_SQRT: v = sqrt(va); da = 0.5 / v; break;
      case OP_SIN: v = sin(va); da = cos(va); break;
      case OP_COS: v = cos(va); da = -sin(va); break;
      case OP_TANH: v = tanh(va); da = 1.0 - v * v; break;
      case OP_EXP: v = exp(va); da = v; break;
      case OP_POWC: v = pow(va, cval[i]); da = cval[i] * pow(va, cval[i] - 1.0); break;
      case OP_FABS: v = fabs(va); da = va >= 0 ? 1.0 : -1.0; break;
      default: v = 0.0;
    }
    w[d] = v;
    if (!(has_t[i] & 1)) continue;
    if (ha && hb) {
      for (int k = 0; k < D; ++k) td[k] = da * ta[k] + db * tb[k];
    } else if (ha) {
      for (int k = 0; k < D; ++k) td[k] = da * ta[k];
    } else {
      for (int k = 0; k < D; ++k) td[k] = db * tb[k];
    }
  }
}

/* X: batch x n_in; vals: batch x n_out; tang: batch x n_out x D (may be NULL when D == 0).
 * in_seed[j] = direction index of input j or -1; out_slot / out_has_t describe the outputs. */
int sxvm_run(int n_instr, const int *op, const int *dst, const int *a, const int *b, const double *cval,
             const unsigned char *has_t, int n_in, const int *in_seed, int n_out, const int *out_slot,
             const unsigned char *out_has_t, int n_slots, int D, int batch, const double *X, double *vals, double *tang,
             int n_threads) {
  int failed = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
  {
    double *w = (double *)malloc(sizeof(double) * (size_t)(n_slots + 1));
    double *t = (double *)malloc(sizeof(double) * ((size_t)n_slots * (D > 0 ? D : 1) + 1));
    if (!w || !t) {
#pragma omp atomic write
      failed = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int bi = 0; bi < batch; ++bi) {
        run_one(n_instr, op, dst, a, b, cval, has_t, D, X + (size_t)bi * n_in, in_seed, w, t);
        for (int j = 0; j < n_out; ++j) {
          vals[(size_t)bi * n_out + j] = w[out_slot[j]];
          if (D > 0 && tang) {
            double *to = tang + ((size_t)bi * n_out + j) * D;
            if (out_has_t[j]) memcpy(to, t + (size_t)out_slot[j] * D, sizeof(double) * D);
            else memset(to, 0, sizeof(double) * D);
          }
        }
      }
    }
    free(w);
    free(t);
  }
  return failed;
}

int sxvm_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

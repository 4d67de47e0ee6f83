"""A minimal recording stand-in for the `casadi` module: just the surface hippopt_b200.plugin's CasADi shim touches
(Sparsity, DM, MX.sym, Callback, Function, nlpsol, Opti), so that the shim's plumbing can be executed where CasADi
is not installed.  `nlpsol` does not optimise: it drives the user-supplied oracle functions in the order IPOPT does
(f, grad_f, g, jac_g at one x, then hess_l with multipliers) for a few iterates and records every call."""
import numpy as np

RECORD = []  # (event, payload) tuples, cleared by the tests


class Sparsity:
    def __init__(self, nrow, ncol, colind=None, row=None):
        self.nrow, self.ncol = int(nrow), int(ncol)
        if colind is None:
            colind = list(range(0, nrow * ncol + 1, nrow)) if nrow else [0] * (ncol + 1)
            row = list(range(nrow)) * ncol
        self._colind, self._row = list(colind), list(row)
        assert len(self._colind) == self.ncol + 1 and self._colind[-1] == len(self._row)

    @staticmethod
    def dense(r, c=1):
        return Sparsity(r, c)

    def nnz(self):
        return len(self._row)

    def numel(self):
        return self.nrow * self.ncol

    def size1(self):
        return self.nrow

    def size2(self):
        return self.ncol

    def colind(self):
        return self._colind

    def row(self):
        return self._row


class DM:
    def __init__(self, a, nz=None):
        if isinstance(a, Sparsity):
            self.sp, self.nz = a, np.asarray(nz, dtype=float).ravel()
            assert self.nz.size == a.nnz()
        else:
            arr = np.atleast_1d(np.asarray(a, dtype=float))
            arr = arr.reshape(arr.shape[0], -1)
            self.sp, self.nz = Sparsity.dense(*arr.shape), arr.ravel(order="F")

    def full(self):
        out = np.zeros((self.sp.nrow, self.sp.ncol))
        col = np.repeat(np.arange(self.sp.ncol), np.diff(self.sp.colind()))
        out[self.sp.row(), col] = self.nz
        return out

    def __array__(self, dtype=None, copy=None):
        return self.full() if dtype is None else self.full().astype(dtype)


class MX:
    def __init__(self, value=0, shape=(1, 1), name=None, fn=None):
        self.value, self.shape, self._name, self.fn = value, shape, name, fn

    @staticmethod
    def sym(name, n, m=1):
        return MX(None, (n, m), name)

    def numel(self):
        return self.shape[0] * self.shape[1]

    def name(self):
        return self._name


class _Call(MX):
    """symbolic application of a Callback"""

    def __init__(self, cb, args):
        super().__init__(None, (cb.get_sparsity_out(0).size1(), cb.get_sparsity_out(0).size2()))
        self.cb, self.args = cb, args


class Callback:
    def construct(self, name, opts):
        self.name_, self.opts_ = name, dict(opts)
        RECORD.append(("callback", name))

    def __call__(self, *args):
        if any(isinstance(a, MX) for a in args):
            return _Call(self, args)
        out = self.eval([a if isinstance(a, DM) else DM(a) for a in args])
        assert len(out) == self.get_n_out()
        for i, o in enumerate(out):  # a real Callback rejects outputs that do not fit the declared sparsity
            sp = self.get_sparsity_out(i)
            assert (o.sp.nrow, o.sp.ncol, o.sp.nnz()) == (sp.nrow, sp.ncol, sp.nnz()), (self.name_, i)
        return out if len(out) > 1 else out[0]


class Function:
    """Function(name, [symbols], [expressions]) over the fake Opti's python-callable expressions."""

    def __init__(self, name, ins, outs):
        self.name_, self.ins, self.outs = name, ins, outs

    def __call__(self, *args):
        env = {id(s): np.asarray(a, dtype=float).ravel() for s, a in zip(self.ins, args)}
        res = [DM(o.fn(env)) for o in self.outs]
        return res if len(res) > 1 else res[0]


class _Nlpsol:
    iterations = 3
    succeed = True

    def __init__(self, name, plugin, nlp, opts):
        self.nlp, self.opts, self.plugin = nlp, opts, plugin
        RECORD.append(("nlpsol", plugin, sorted(opts)))

    def __call__(self, x0, p, lbg, ubg, lbx=None, ubx=None, lam_g0=None):
        x = np.asarray(x0, dtype=float).ravel().copy()
        p = np.asarray(p, dtype=float).ravel()
        f_cb, g_cb = self.nlp["f"].cb, self.nlp["g"].cb
        m = g_cb.get_sparsity_out(0).size1()
        assert len(lbg) == m and len(ubg) == m and len(lbx) == len(x)
        lam = np.linspace(-1.0, 1.0, m) if lam_g0 is None else np.asarray(lam_g0, dtype=float).ravel()
        RECORD.append(("solve", dict(n=len(x), m=m, finite_lbx=int(np.isfinite(lbx).sum()))))
        for it in range(self.iterations):
            f = float(np.asarray(f_cb(x, p)).ravel()[0])
            f2, grad = self.opts["grad_f"](x, p)
            g = np.asarray(g_cb(x, p)).ravel()
            g2, jac = self.opts["jac_g"](x, p)
            assert float(np.asarray(f2).ravel()[0]) == f and np.array_equal(np.asarray(g2).ravel(), g)
            if "hess_lag" in self.opts:
                h = self.opts["hess_lag"](x, p, 1.0, lam)
                RECORD.append(("iter", it, f, float(np.abs(h.nz).sum())))
            else:
                RECORD.append(("iter", it, f, None))
            x = x - 1e-3 * np.asarray(grad).ravel()
        self._x, self._f, self._g, self._lam = x, f, g, lam
        return {"x": DM(x), "f": DM([f]), "g": DM(g), "lam_g": DM(lam), "lam_x": DM(np.zeros(len(x))), "lam_p": DM([0.0])}

    def stats(self):
        return {"success": self.succeed, "return_status": "Solve_Succeeded" if self.succeed else "Infeasible_Problem_Detected"}


def nlpsol(name, plugin, nlp, opts):
    return _Nlpsol(name, plugin, nlp, opts)


class _Debug:
    def __init__(self, opti):
        self.opti = opti

    def value(self, expr, *initial):
        return self.opti._value(expr)


class Opti:
    """Opti over python-callable expressions: variables / parameters are slices of x / p, `f` and `g` are given
    functions of (x, p) -- the canonical form Opti would have baked."""

    def __init__(self, f_fn, g_fn, x0, p, lbg, ubg):
        self._x0, self._p, self._lbg, self._ubg = (np.asarray(a, dtype=float) for a in (x0, p, lbg, ubg))
        self.nx, self.np, self.ng = len(self._x0), len(self._p), len(self._lbg)
        self.x, self.p, self.lam_g = MX.sym("x", self.nx), MX.sym("p", self.np), MX.sym("lam_g", self.ng)
        self.f = MX(fn=lambda env: f_fn(env[id(self.x)], env[id(self.p)]))
        self.g = MX(fn=lambda env: g_fn(env[id(self.x)], env[id(self.p)]), shape=(self.ng, 1))
        self.lbg, self.ubg = MX(name="lbg"), MX(name="ubg")
        self.debug = _Debug(self)
        self.minimized = None

    def variable_slice(self, lo, hi):
        v = MX.sym(f"x[{lo}:{hi}]", hi - lo)
        v.fn = lambda env: env[id(self.x)][lo:hi]
        return v

    def parameter_slice(self, lo, hi):
        v = MX.sym(f"p[{lo}:{hi}]", hi - lo)
        v.fn = lambda env: env[id(self.p)][lo:hi]
        v.par = (lo, hi)
        return v

    def constraint(self, lo, hi, flipped=False):
        c = MX(name=f"con[{lo}:{hi}]", shape=(hi - lo, 1))
        c.rows, c.flipped = (lo, hi), flipped
        return c

    def dual(self, con):
        lo, hi = con.rows
        sign = -1.0 if con.flipped else 1.0
        return MX(fn=lambda env: sign * env[id(self.lam_g)][lo:hi], shape=(hi - lo, 1))

    def expression(self, fn, n=1):
        return MX(fn=lambda env: fn(env[id(self.x)], env[id(self.p)]), shape=(n, 1))

    def minimize(self, cost):
        self.minimized = cost

    def initial(self):
        return "initial"

    def _value(self, expr):
        if expr is self.x:
            return self._x0
        if expr is self.p:
            return self._p
        if expr is self.lbg:
            return self._lbg
        if expr is self.ubg:
            return self._ubg
        if hasattr(expr, "par"):
            return self._p[expr.par[0]:expr.par[1]]
        raise KeyError(expr)


# ---- generated oracle functions (tools/dump_casadi_golden.py): nlpsol(...).get_function("nlp_jac_g") etc. ----------
class _OracleFunction:
    """What CasADi generates from the symbolic problem, backed here by an object with the oracle NLP's interface
    (eval_f / eval_grad_f / eval_g / eval_jac / eval_hess / jac_structure / hess_structure) attached to the Opti."""

    def __init__(self, kind, nlp, keep_rows=None, jac_keep=None, jac_sp=None):
        self.kind, self.nlp, self.keep_rows, self.jac_keep, self.jac_sp = kind, nlp, keep_rows, jac_keep, jac_sp

    def sparsity_out(self, i):
        n, m = self.nlp.n_x, self.nlp.m
        if self.kind == "nlp_jac_g" and i == 1:
            if self.jac_sp is not None:
                return self.jac_sp
            c, r = self.nlp.jac_structure()[:2]
            return Sparsity(m, n, list(c), list(r))
        if self.kind == "nlp_hess_l":
            c, r = self.nlp.hess_structure()[:2]
            return Sparsity(n, n, list(c), list(r))
        raise KeyError((self.kind, i))

    def __call__(self, *args):
        a = [np.asarray(v, dtype=float).ravel() for v in args]
        X, P = a[0][None], a[1][None]
        RECORD.append(("oracle_call", self.kind))
        if self.kind == "nlp_f":
            return DM([self.nlp.eval_f(X, P)[0]])
        if self.kind == "nlp_grad_f":
            return [DM([self.nlp.eval_f(X, P)[0]]), DM(self.nlp.eval_grad_f(X, P)[0])]
        if self.kind == "nlp_g":
            return DM(self.nlp.eval_g(X, P)[0])
        if self.kind == "nlp_jac_g":
            vals = self.nlp.eval_jac(X, P)[0]
            if self.jac_keep is not None:
                vals = vals[self.jac_keep]
            return [DM(self.nlp.eval_g(X, P)[0]), DM(self.sparsity_out(1), vals)]
        if self.kind == "nlp_hess_l":
            lam = a[3] if self.keep_rows is None else _scatter(a[3], self.keep_rows, self.nlp.m)
            return DM(self.sparsity_out(0), self.nlp.eval_hess(X, P, lam[None], np.asarray([a[2][0]]))[0])
        raise KeyError(self.kind)


def _scatter(v, rows, m):
    out = np.zeros(m)
    out[rows] = v
    return out


def _nlpsol_get_function(self, name):
    opti_nlp = self.nlp["f"].oracle if hasattr(self.nlp["f"], "oracle") else None
    assert opti_nlp is not None, "attach an oracle to the fake Opti (Opti.attach_oracle)"
    nlp, reduction = opti_nlp
    if self.opts.get("detect_simple_bounds") and reduction is not None:
        sp = Sparsity(reduction.m_reduced, nlp.n_x, list(reduction.jac_colind), list(reduction.jac_row))
        return _OracleFunction(name, nlp, reduction.general, reduction.jac_keep, sp)
    return _OracleFunction(name, nlp)


def _nlpsol_simple_bounds(self, p):
    nlp, reduction = self.nlp["f"].oracle
    lb, ub = nlp.eval_bounds(np.asarray(p, dtype=float)[None])
    _, _, lbx, ubx = reduction.bounds(lb[0], ub[0])
    return reduction.general, lbx, ubx


_Nlpsol.get_function = _nlpsol_get_function
_Nlpsol.simple_bounds = _nlpsol_simple_bounds


def _opti_attach_oracle(self, nlp, reduction=None):
    """make `opti.f` carry the oracle so that nlpsol({... "f": opti.f ...}).get_function works; lbg / ubg become
    functions of p"""
    self.f.oracle = (nlp, reduction)
    self.lbg.fn = lambda env: nlp.eval_bounds(env[id(self.p)][None])[0][0]
    self.ubg.fn = lambda env: nlp.eval_bounds(env[id(self.p)][None])[1][0]


Opti.attach_oracle = _opti_attach_oracle

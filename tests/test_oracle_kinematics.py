"""Independent cross-checks of the oracle's restatement of adam / liecasadi quantities (no install of
either exists here): first-principles sums over bodies with plain numpy, and
finite-difference checks of the oracle's AD.  CPU only."""
import numpy as np
import pytest

from oracle import expressions as ex
from oracle import kinodynamic as kd
from oracle import robot, sx


@pytest.fixture(scope="module")
def kin(model):
    pb, q, s = sx.syms("pb", 3), sx.syms("q", 4), sx.syms("s", 23)
    Hb = robot.base_pose(pb, robot.quat_normalize(q))
    AG = robot.centroidal_momentum_matrix(model, Hb, s)
    com = robot.com_position(model, Hb, s)
    Hs = robot.body_transforms(model, Hb, s)
    outs = list(AG.ravel()) + list(com) + [Hs[b][i, j] for b in range(model.n_bodies) for i in range(3) for j in range(4)]
    tape = sx.Tape(outs, list(pb) + list(q) + list(s))

    def ev(x):
        o = tape.eval(np.asarray(x)[None])[0]
        return o[:174].reshape(6, 29), o[174:177], o[177:].reshape(model.n_bodies, 3, 4)

    return ev


def test_quaternion_rotation_convention():
    q = np.array([0.1, -0.3, 0.2, 0.9])
    q /= np.linalg.norm(q)
    R = np.array([[float(v) for v in row] for row in robot.quat_to_rot(q)])
    assert R @ R.T == pytest.approx(np.eye(3), abs=1e-15)
    assert np.linalg.det(R) == pytest.approx(1.0)
    # rotation about z by 90 degrees: xyzw = (0, 0, sin 45, cos 45) maps e_x to e_y
    Rz = np.array([[float(v) for v in row] for row in robot.quat_to_rot([0, 0, np.sqrt(0.5), np.sqrt(0.5)])])
    assert Rz @ np.array([1.0, 0, 0]) == pytest.approx([0, 1.0, 0], abs=1e-15)


def test_centroidal_momentum_matches_first_principles(model, kin):
    """h_G = A_G nu against sum_l m (c_l - x) x c_dot_l + R I R^T w_l with body velocities from
    central differences of the forward kinematics."""
    rng = np.random.default_rng(1)
    x0 = np.concatenate([rng.normal(0, 1, 3), rng.normal(0, 1, 4), rng.uniform(-1, 1, 23)])
    A, c, Hm = kin(x0)
    vb, qd, sd = rng.normal(0, 1, 3), rng.normal(0, 1, 4), rng.normal(0, 1, 23)
    qh = x0[3:7] / np.linalg.norm(x0[3:7])
    qd -= qh * (qh @ qd)
    omega = 2 * (-qd[3] * qh[:3] + qh[3] * qd[:3] - np.cross(qd[:3], qh[:3]))
    h = A @ np.concatenate([vb, omega, sd])
    eps = 1e-6

    def poses(t):
        x = x0.copy()
        x[:3] += t * vb
        x[3:7] = qh + t * qd
        x[7:] += t * sd
        return kin(x)

    _, cp, Hp = poses(eps)
    _, cm, Hmm = poses(-eps)
    hl, ha = np.zeros(3), np.zeros(3)
    for b in range(model.n_bodies):
        R, o = Hm[b][:, :3], Hm[b][:, 3]
        cl = R @ model.com[b] + o
        cdot = ((Hp[b][:, :3] @ model.com[b] + Hp[b][:, 3]) - (Hmm[b][:, :3] @ model.com[b] + Hmm[b][:, 3])) / (2 * eps)
        W = (Hp[b][:, :3] - Hmm[b][:, :3]) / (2 * eps) @ R.T
        w = np.array([W[2, 1], W[0, 2], W[1, 0]])
        hl += model.mass[b] * cdot
        ha += model.mass[b] * np.cross(cl - c, cdot) + R @ model.inertia[b] @ R.T @ w
    assert h == pytest.approx(np.concatenate([hl, ha]), abs=1e-7)
    assert (cp - cm) / (2 * eps) * model.total_mass() == pytest.approx(h[:3], abs=1e-7)


def test_com_is_mass_weighted_mean(model, kin):
    rng = np.random.default_rng(3)
    x0 = np.concatenate([rng.normal(0, 1, 3), rng.normal(0, 1, 4), rng.uniform(-1, 1, 23)])
    _, c, Hm = kin(x0)
    acc = sum(model.mass[b] * (Hm[b][:, :3] @ model.com[b] + Hm[b][:, 3]) for b in range(model.n_bodies))
    assert c == pytest.approx(acc / model.total_mass(), abs=1e-14)


def test_kinodynamic_derivatives_against_finite_differences(model):
    nlp, lay = kd.build(model, kd.Settings(horizon=3, final_state_constraint=True, periodicity_constraint=True))
    rng = np.random.default_rng(0)
    B = 1
    X = rng.normal(0, 0.3, (B, nlp.n_x))
    P = rng.uniform(0.5, 1.5, (B, nlp.n_p))
    lam = rng.normal(0, 1, (B, nlp.m))
    J = nlp.dense_jac(X, P)
    Hd = nlp.dense_hess(X, P, lam, 0.7)
    gf = nlp.eval_grad_f(X, P)
    eps = 1e-6

    def gl(Xq):
        return 0.7 * nlp.eval_grad_f(Xq, P) + np.einsum("bm,bmn->bn", lam, nlp.dense_jac(Xq, P))

    for j in rng.choice(nlp.n_x, 12, replace=False):
        Xp, Xm = X.copy(), X.copy()
        Xp[:, j] += eps
        Xm[:, j] -= eps
        fdg = (nlp.eval_g(Xp, P) - nlp.eval_g(Xm, P)) / (2 * eps)
        assert np.abs(fdg - J[:, :, j]).max() < 1e-6 * max(1.0, np.abs(J[:, :, j]).max())
        fdf = (nlp.eval_f(Xp, P) - nlp.eval_f(Xm, P)) / (2 * eps)
        assert np.abs(fdf - gf[:, j]).max() < 1e-5 * max(1.0, np.abs(gf[:, j]).max())
        fdh = (gl(Xp) - gl(Xm)) / (2 * eps)
        assert np.abs(fdh - Hd[:, :, j]).max() < 1e-5 * max(1.0, np.abs(Hd[:, :, j]).max())

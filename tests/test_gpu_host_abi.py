"""hb_eval_host: the host-buffer form of the C ABI (what a CPU-side IPOPT shim calls).

Checked against the device-pointer form hb_eval, which the parity tests pin to the oracle: the host
pipeline only moves bytes, so its results must be bit-identical whatever the chunking."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _device_eval(ev, mask, x, p, lam, sigma):
    d = torch.device("cuda:0")
    t = [torch.tensor(np.ascontiguousarray(a), device=d) for a in (x, p, lam, sigma)]
    out = ev.eval(mask, *t)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy().copy() for k, v in out.items()}


@pytest.mark.parametrize("B,pinned", [(1, True), (130, True), (300, False)])
def test_host_pipeline_matches_device_call(model, built_library, B, pinned):
    """Ragged batches (one chunk, chunk + 2, two chunks + 44), pinned and pageable host memory."""
    from hippopt_b200.evaluator import ALL, HostPipeline, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    ev = KinoEvaluator(model, KinoSettings(horizon=3))
    x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=11, noise=0.1)
    ref = _device_eval(ev, ALL, x, p, lam, sigma)
    pipe = HostPipeline(ev, B, ALL, pinned=pinned)
    host = [HostPipeline.host_buffer(a.shape, pinned).copy_(torch.from_numpy(np.ascontiguousarray(a)))
            for a in (x, p, lam, sigma)]
    pipe.set_parameters(host[1])
    for _ in range(2):  # second call reuses the staging slabs
        out = pipe.run(host[0], host[2], host[3])
        for k in ref:
            assert np.array_equal(out[k].numpy(), ref[k]), k
    assert pipe.h2d_bytes == 8 * B * (ev.n_x + ev.m + 1)
    assert pipe.d2h_bytes == 8 * B * (1 + ev.n_x + ev.m + ev.nnz_j + ev.nnz_h)
    assert ev.last_launch_count() == 3 * -(-B // 128)


def test_host_pipeline_partial_mask_and_shared_parameters(model, built_library):
    from hippopt_b200.evaluator import F, G, JAC_G, HostPipeline, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    ev = KinoEvaluator(model, KinoSettings(horizon=3))
    x, p, lam, sigma = kino_batch(ev.layout, model, 7, seed=12, noise=0.1)
    p[:] = p[0]
    ref = _device_eval(ev, F | G | JAC_G, x, p, lam, sigma)
    pipe = HostPipeline(ev, 7, F | G | JAC_G)
    pipe.set_parameters(torch.from_numpy(p[0].copy()))  # one vector shared by all instances
    out = pipe.run(torch.from_numpy(x))
    assert set(out) == {"f", "g", "jac"}
    for k in ref:
        assert np.array_equal(out[k].numpy(), ref[k]), k
    assert pipe.h2d_bytes == 8 * 7 * ev.n_x


def test_host_abi_errors_and_pinned_allocator(model, built_library):
    from hippopt_b200 import _capi
    from hippopt_b200.evaluator import ALL, HostPipeline, ToyEvaluator

    L = _capi.lib()
    ev = ToyEvaluator(horizon=6, integrator="euler")
    pipe = HostPipeline(ev, 4, ALL)
    x = torch.zeros(4, ev.n_x, dtype=torch.float64)
    lam = torch.zeros(4, ev.m, dtype=torch.float64)
    sg = torch.ones(4, dtype=torch.float64)
    with pytest.raises(_capi.EvaluationError, match="hb_host_set_parameters first"):
        pipe.run(x, lam, sg)
    pipe.set_parameters(torch.tensor([[1.0, 2.0, 3.0]] * 2, dtype=torch.float64)[:, :ev.n_p].contiguous())
    with pytest.raises(_capi.EvaluationError, match="exceeds the batch"):
        pipe.run(x, lam, sg)
    with pytest.raises(ValueError):
        pipe.run(x[:, :-1].contiguous(), lam, sg)
    # toy problem through the host ABI with memory from hb_host_alloc
    n = 4 * ev.n_x
    ptr = ctypes.c_void_p()
    _capi.check(L.hb_host_alloc(ctypes.byref(ptr), 8 * n), "hb_host_alloc")
    xa = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(4, ev.n_x))
    xa[:] = np.random.default_rng(0).normal(size=(4, ev.n_x))
    pv = np.tile(np.arange(1.0, ev.n_p + 1.0), (4, 1))
    pipe.set_parameters(torch.from_numpy(pv))
    f = np.zeros(4)
    rc = L.hb_eval_host(ev._h, 1, ptr, None, None, f.ctypes.data_as(ctypes.c_void_p), None, None, None, None, 4)
    _capi.check(rc, "hb_eval_host")
    ref = _device_eval(ev, 1, xa.copy(), pv, np.zeros((4, ev.m)), np.ones(4))
    assert np.array_equal(f, ref["f"])
    _capi.check(L.hb_host_free(ptr), "hb_host_free")
    assert L.hb_host_alloc(None, 8) != 0 and L.hb_host_alloc(ctypes.byref(ptr), 0) != 0


def test_c_client_without_python(model, built_library, tmp_path):
    """(b) C-ABI completeness: a plain C program links libhippopt_b200.so, opens a problem file written by hb_save and
    gets dimensions, CCS patterns, bounds and all five evaluations without any Python-side layout code."""
    import shutil
    import subprocess

    from hippopt_b200.evaluator import ALL, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ev = KinoEvaluator(model, KinoSettings(horizon=4, final_state_constraint=True))
    lay = ev.layout
    x, p, lam, sigma = kino_batch(lay, model, 1, seed=12, noise=0.1)
    blob, vec, exe = (str(tmp_path / n) for n in ("problem.bin", "vectors.bin", "client"))
    ev.save(blob)
    np.concatenate([x[0], p[0], lam[0], sigma]).astype(np.float64).tofile(vec)
    libdir = os.path.join(root, "hippopt_b200")
    subprocess.run(["gcc", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c_client", "client.c"),
                    "-o", exe, "-L", libdir, "-l:libhippopt_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe, blob, vec], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.splitlines()}
    assert [int(v) for v in got["dims"]] == [ev.n_x, ev.n_p, ev.m, ev.nnz_j, ev.nnz_h]
    assert [int(v) for v in got["pattern"]] == [lay.jac_colind[-1], lay.jac_row[-1], lay.hess_colind[-1], lay.hess_row[-1]]
    lb, ub = lay.bounds(p)
    assert int(got["bounds"][0]) == int((lb[0] == ub[0]).sum())
    assert float(got["bounds"][1]) == pytest.approx(lb[0][np.isfinite(lb[0])].sum(), rel=1e-14)
    assert float(got["bounds"][2]) == pytest.approx(ub[0][np.isfinite(ub[0])].sum(), rel=1e-14)
    out = {k: v.cpu().numpy()[0] for k, v in ev.eval(ALL, *(torch.tensor(a, device="cuda:0") for a in (x, p, lam, sigma))).items()}
    want = [float(out["f"]), float((out["grad_f"] * (np.arange(ev.n_x) % 11 + 1)).sum()),
            float((out["g"] * (np.arange(ev.m) % 7 + 1)).sum()), float((out["jac"] * (lay.jac_row % 5 + 1)).sum()),
            float((out["hess"] * (lay.hess_row % 3 + 1)).sum())]
    assert [float(v) for v in got["values"]] == pytest.approx(want, rel=1e-12)


def test_casadi_external_function_abi(model, built_library):
    """the CasADi codegen ABI (`casadi.external("hb_nlp_jac_g", "libhippopt_b200.so")`) of the five nlpsol
    oracle functions, called the way CasADi calls an external function -- arg / res pointer arrays, sparsities as
    compact CCS vectors, work sizes -- and compared with hb_eval; f, grad_f, g, jac_g at one x cost one evaluation."""
    from hippopt_b200.evaluator import ALL, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    lib = built_library
    ev = KinoEvaluator(model, KinoSettings(horizon=3, final_state_constraint=True))
    lay = ev.layout
    x, p, lam, sigma = kino_batch(lay, model, 1, seed=31, noise=0.1)
    ref = {k: v.cpu().numpy()[0] for k, v in ev.eval(ALL, *(torch.tensor(a, device="cuda:0") for a in (x, p, lam, sigma))).items()}
    assert lib.hb_external_bind(ev._h) == 0
    cint, dp = ctypes.c_longlong, ctypes.POINTER(ctypes.c_double)

    def sparsity(fn, i):
        f = getattr(lib, fn)
        f.restype, f.argtypes = ctypes.POINTER(cint), [cint]
        s = f(i)
        nrow, ncol = s[0], s[1]
        colind = [s[2 + j] for j in range(ncol + 1)]
        rows = [s[2 + ncol + 1 + j] for j in range(colind[-1])]
        return nrow, ncol, colind, rows

    def call(name, ins, out_sizes):
        f = getattr(lib, name)
        f.restype = ctypes.c_int
        n_in, n_out = getattr(lib, name + "_n_in"), getattr(lib, name + "_n_out")
        n_in.restype = n_out.restype = cint
        assert (n_in(), n_out()) == (len(ins), len(out_sizes))
        sz = [cint() for _ in range(4)]
        assert getattr(lib, name + "_work")(*[ctypes.byref(v) for v in sz]) == 0 and sz[0].value >= len(ins)
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in ins]
        outs = [np.full(n, np.nan) for n in out_sizes]
        arg = (dp * len(arrs))(*[a.ctypes.data_as(dp) for a in arrs])
        res = (dp * len(outs))(*[o.ctypes.data_as(dp) for o in outs])
        assert f(arg, res, None, None, 0) == 0, lib.hb_last_error()
        return outs

    try:
        assert sparsity("hb_nlp_jac_g_sparsity_out", 1) == (ev.m, ev.n_x, list(lay.jac_colind), list(lay.jac_row))
        assert sparsity("hb_nlp_hess_l_sparsity_out", 0) == (ev.n_x, ev.n_x, list(lay.hess_colind), list(lay.hess_row))
        assert sparsity("hb_nlp_f_sparsity_in", 0)[:2] == (ev.n_x, 1) and sparsity("hb_nlp_f_sparsity_in", 1)[:2] == (ev.n_p, 1)
        nm = lib.hb_nlp_hess_l_name_in
        nm.restype, nm.argtypes = ctypes.c_char_p, [cint]
        assert [nm(i) for i in range(4)] == [b"x", b"p", b"lam_f", b"lam_g"]
        (f,) = call("hb_nlp_f", [x[0], p[0]], [1])
        f2, grad = call("hb_nlp_grad_f", [x[0], p[0]], [1, ev.n_x])
        (g,) = call("hb_nlp_g", [x[0], p[0]], [ev.m])
        g2, jac = call("hb_nlp_jac_g", [x[0], p[0]], [ev.m, ev.nnz_j])
        (hess,) = call("hb_nlp_hess_l", [x[0], p[0], sigma, lam[0]], [ev.nnz_h])
        assert f[0] == ref["f"] and f2[0] == ref["f"] and np.array_equal(grad, ref["grad_f"])
        assert np.array_equal(g, ref["g"]) and np.array_equal(g2, ref["g"]) and np.array_equal(jac, ref["jac"])
        assert np.array_equal(hess, ref["hess"])
        n1, n2 = ctypes.c_int64(), ctypes.c_int64()
        lib.hb_external_stats(ctypes.byref(n1), ctypes.byref(n2))
        assert (n1.value, n2.value) == (1, 1)  # four first-order calls at one x: ONE evaluation
        call("hb_nlp_g", [x[0] + 1e-3, p[0]], [ev.m])
        lib.hb_external_stats(ctypes.byref(n1), ctypes.byref(n2))
        assert n1.value == 2
    finally:
        assert lib.hb_external_bind(None) == 0
    arg = (dp * 2)(x[0].ctypes.data_as(dp), p[0].ctypes.data_as(dp))
    res = (dp * 1)(np.zeros(1).ctypes.data_as(dp))
    lib.hb_nlp_f.restype = ctypes.c_int
    assert lib.hb_nlp_f(arg, res, None, None, 0) != 0 and b"no problem bound" in lib.hb_last_error()


def test_repeated_single_chunk_calls_replay_a_graph(model, built_library):
    """What a CPU-side IPOPT does: the same pinned buffers, new contents, call after call.  From the third call on
    hb_eval_host replays a captured CUDA graph; every call must return what a plain device evaluation returns, also after
    the parameters or the mask change."""
    from hippopt_b200.evaluator import ALL, F, G, GRAD_F, JAC_G, HostPipeline, KinoEvaluator
    from hippopt_b200.kino_layout import KinoSettings
    from hippopt_b200.workloads import kino_batch

    ev = KinoEvaluator(model, KinoSettings(horizon=3))
    B = 2
    pipe = HostPipeline(ev, B, ALL)
    first = HostPipeline(ev, B, F | GRAD_F | G | JAC_G)
    xh, lh, sh = (HostPipeline.host_buffer(s) for s in ((B, ev.n_x), (B, ev.m), (B,)))
    for it in range(6):
        x, p, lam, sigma = kino_batch(ev.layout, model, B, seed=20 + it, noise=0.1)
        if it in (0, 4):  # new parameters mid-sequence (same size: the graphs stay valid, the values are read from d_p)
            pipe.set_parameters(torch.from_numpy(np.ascontiguousarray(p)))
            p_now = p
        xh.copy_(torch.from_numpy(x))
        lh.copy_(torch.from_numpy(lam))
        sh.copy_(torch.from_numpy(sigma))
        ref = _device_eval(ev, ALL, x, p_now, lam, sigma)
        out = pipe.run(xh, lh, sh)
        for k in ref:
            assert np.array_equal(out[k].numpy(), ref[k]), (it, k)
        out1 = first.run(xh)  # the other mask, interleaved: its own graph
        for k in ("f", "grad_f", "g", "jac"):
            assert np.array_equal(out1[k].numpy(), ref[k]), (it, k)
        assert ev.last_launch_count() == 3

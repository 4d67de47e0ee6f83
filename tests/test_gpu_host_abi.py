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

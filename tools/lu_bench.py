"""Developer benchmark of the batched LU kernels (csrc/lu.cu) against torch.linalg on the KKT stage-block size."""
import sys
import time

import torch

sys.path.insert(0, ".")
from hippopt_b200.kkt import lu_factor, lu_solve  # noqa: E402

d = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n, r = 337, 88
A = torch.randn(B, n, n, dtype=torch.float64, device=d) + 20 * torch.eye(n, dtype=torch.float64, device=d)
R = torch.randn(B, n, r, dtype=torch.float64, device=d)


def t(f, reps=5):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


F, piv, info = lu_factor(A.clone(), symmetric=True)
X = lu_solve(F, piv, R)
print(f"B={B} n={n}: residual {(torch.bmm(A.transpose(1, 2), X) - R).abs().max().item():.1e}")
tf = t(lambda: lu_factor(A.clone(), symmetric=True))
ts = t(lambda: lu_solve(F, piv, R))
print(f"hb    factor {tf:.3f} ms ({B * 2 / 3 * n ** 3 / tf / 1e9:.2f} TFLOP/s)  solve({r}) {ts:.3f} ms")
lu, p2, _ = torch.linalg.lu_factor_ex(A)
tf = t(lambda: torch.linalg.lu_factor_ex(A))
ts = t(lambda: torch.linalg.lu_solve(lu, p2, R))
print(f"torch factor {tf:.3f} ms ({B * 2 / 3 * n ** 3 / tf / 1e9:.2f} TFLOP/s)  solve({r}) {ts:.3f} ms")

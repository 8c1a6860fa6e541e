import sys, os
sys.path.insert(0, '/root/repo')
import torch, tensorly_b200 as tb
torch.manual_seed(0)
for n, p, cols in [(96, 16, 512), (512, 64, 4096), (512, 32, 4096), (200, 48, 1000)]:
    y = torch.rand(n, cols, device="cuda", dtype=torch.float64)
    G = y @ y.T
    u0 = tb.orthonormalize(torch.rand(n, p, device="cuda", dtype=torch.float64))
    for steps in (1, 3, 8):
        u = tb.subspace_iterate(G, u0.clone(), steps)
        d = float(torch.linalg.norm(u.T @ u - torch.eye(p, device="cuda", dtype=torch.float64)))
        # subspace quality vs exact eigenvectors
        w, v = torch.linalg.eigh(G)
        vt = v[:, -p:]
        q = torch.linalg.qr(u).Q
        res = float(torch.linalg.norm(q - vt @ (vt.T @ q)))
        print(n, p, steps, "orth defect %.2e" % d, "subspace residual %.2e" % res)

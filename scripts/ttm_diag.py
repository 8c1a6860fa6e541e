"""Locate wrong rows of the last-mode TTM on the tcgen05 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensorly_b200 as tb
g = torch.Generator(device="cuda").manual_seed(0)
for (L, J, I) in ((32768, 512, 64), (32768, 512, 64), (65536, 256, 64), (32768, 512, 32)):
    x = torch.randn(L, J, generator=g, device="cuda")
    m = torch.randn(I, J, generator=g, device="cuda")
    ref = (x.double() @ m.double().t())
    out = tb.mode_dot(x, m, 1)
    err_rows = (out.double() - ref).norm(dim=1) / ref.norm(dim=1)
    bad = (err_rows > 1e-4).nonzero().flatten()
    tiles = sorted(set((bad // 128).tolist()))
    print(f"L={L} J={J} I={I} path={tb.last_kernel_path()} total err {float((out.double()-ref).norm()/ref.norm()):.2e} bad rows {bad.numel()} bad tiles {len(tiles)} first {tiles[:12]}")
    if bad.numel():
        print("   bad rows (tile,row-in-tile):", [(int(b)//128, int(b)%128) for b in bad[:40]])
        r = int(bad[0]); cols = ((out[r].double() - ref[r]).abs() > 1e-3 * ref[r].abs().max()).nonzero().flatten().tolist()
        print("   first bad row", r, "bad cols", cols[:16], "n", len(cols), "out/ref", float(out[r, cols[0]]), float(ref[r, cols[0]]))

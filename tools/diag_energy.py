"""Diagnostic: per-sample energy ratio and error pattern of the 1024^2 forward call (unitary at the default optics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import style_transfer_based_holographic_imaging_b200 as pkg
from oracle import asm_oracle as ao
n, b = int(sys.argv[1]), int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = torch.Generator(device="cuda").manual_seed(1)
O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
z = (0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3
e_in = (O.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64)
first = None
for r in range(reps):
    U = pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, False)
    torch.cuda.synchronize()
    e = ((U.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64) / e_in - 1).abs()
    bad = torch.nonzero(e > 2e-6).flatten().tolist()
    same = None if first is None else bool(torch.equal(U, first))
    print(f"rep {r}: max energy dev {e.max().item():.3e}, bad samples {bad}, identical to rep 0: {same}")
    if first is None: first = U.clone()
    for s in bad[:2]:
        ref = ao.asm(O[s:s+1].cpu().numpy(), 532e-9, z[s:s+1].cpu().numpy(), 1.5e-6, False)[0, 0]
        d = np.abs(U[s, 0].cpu().numpy() - ref)
        rows = np.nonzero(d.max(axis=1) > 1e-3 * np.abs(ref).max())[0]
        cols = np.nonzero(d.max(axis=0) > 1e-3 * np.abs(ref).max())[0]
        print(f"   sample {s}: rel-L2 {ao.rel_l2(U[s,0].cpu().numpy(), ref):.3e}; bad rows {rows[:20]} ({len(rows)}), bad cols {cols[:20]} ({len(cols)})")

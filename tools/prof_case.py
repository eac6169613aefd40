"""Small driver for ncu: a few forward (+adjoint) calls at a given size.  python tools/prof_case.py N B [pad] [reps] [fwd_only]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import style_transfer_based_holographic_imaging_b200 as pkg
from style_transfer_based_holographic_imaging_b200 import _lib as L
n = int(sys.argv[1]); b = int(sys.argv[2]); pad = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
fwd_only = bool(int(sys.argv[5])) if len(sys.argv) > 5 else False
g = torch.Generator(device="cuda").manual_seed(0)
O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
I = torch.empty(b, 1, n, n, device="cuda")
A = torch.empty_like(O)
for _ in range(reps):
    pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out_mode=L.OUT_INTENSITY, out=I)
    if not fwd_only:
        pkg.asm_adjoint_raw(O, z, 532e-9, 1.5e-6, pad, out=A)
torch.cuda.synchronize()
print("ok", float(I.sum()), float(A.abs().sum()))

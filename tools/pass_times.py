"""Per-pass device time (CUDA events inside the library, serialised).  python tools/pass_times.py N B [pad]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import style_transfer_based_holographic_imaging_b200 as pkg
from style_transfer_based_holographic_imaging_b200 import _lib as L
n = int(sys.argv[1]); b = int(sys.argv[2]); pad = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
lib = L.load()
g = torch.Generator(device="cuda").manual_seed(0)
O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
I = torch.empty(b, 1, n, n, device="cuda")
for _ in range(2): pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out_mode=L.OUT_INTENSITY, out=I)
torch.cuda.synchronize()
lib.asm_b200_profile(1, None)
reps = 3
for _ in range(reps): pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out_mode=L.OUT_INTENSITY, out=I)
torch.cuda.synchronize()
ms = (ctypes.c_double * 3)(); lib.asm_b200_profile(0, ms)
print(f"N={n} B={b} pad={pad}: per image  rows_fwd {ms[0]/reps/b*1e3:.2f} us  cols {ms[1]/reps/b*1e3:.2f} us  rows_inv {ms[2]/reps/b*1e3:.2f} us")

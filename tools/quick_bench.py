"""Device-resident throughput of fwd (+adjoint) for a size; python tools/quick_bench.py N B [pad] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import style_transfer_based_holographic_imaging_b200 as pkg
from style_transfer_based_holographic_imaging_b200 import _lib as L
n = int(sys.argv[1]); b = int(sys.argv[2]); pad = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
g = torch.Generator(device="cuda").manual_seed(0)
O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
G = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
I = torch.empty(b, 1, n, n, device="cuda"); A = torch.empty_like(O)
def step():
    pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out_mode=L.OUT_INTENSITY, out=I)
    pkg.asm_adjoint_raw(G, z, 532e-9, 1.5e-6, pad, out=A)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ups = b / (ms * 1e-3)
print(f"N={n} B={b} pad={pad} lanes={os.environ.get('ASM_B200_LANES','def')} chunkMB={os.environ.get('ASM_B200_CHUNK_MB','def')}: "
      f"{ms:.3f} ms/step  {ups:.0f} units/s  {ups*28*n*n/1e9:.0f} GB/s algorithmic ({ups*28*n*n/1e9/6496.8*100:.1f}% of 6496.8)")

"""Device-resident throughput of the forward call and of the adjoint call separately; python tools/split_bench.py N B [pad]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import style_transfer_based_holographic_imaging_b200 as pkg
from style_transfer_based_holographic_imaging_b200 import _lib as L
n = int(sys.argv[1]); b = int(sys.argv[2]); pad = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
g = torch.Generator(device="cuda").manual_seed(0)
O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
I = torch.empty(b, 1, n, n, device="cuda"); A = torch.empty_like(O)
amp = torch.rand(b, 1, n, n, device="cuda", generator=g); ph = torch.rand(b, 1, n, n, device="cuda", generator=g)
cases = {
    "forward complex64 -> |U|^2": lambda: pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out_mode=L.OUT_INTENSITY, out=I),
    "forward complex64 -> complex64": lambda: pkg.asm_forward_raw(O, z, 532e-9, 1.5e-6, pad, out=A),
    "adjoint complex64 -> complex64": lambda: pkg.asm_adjoint_raw(O, z, 532e-9, 1.5e-6, pad, out=A),
}
for name, f in cases.items():
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"N={n} B={b} pad={pad} {name:34s}: {ms:.3f} ms/call  {ms / b * 1e3:.2f} us/sample")

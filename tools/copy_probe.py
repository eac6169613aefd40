"""L2 byte accounting probe: lts__t_bytes of a plain streaming copy (256 MB in, 256 MB out)."""
import torch
a = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
b = torch.ones_like(a)
for _ in range(3):
    torch.add(b, 1.0, out=a)
torch.cuda.synchronize()

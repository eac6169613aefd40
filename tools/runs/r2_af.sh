#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_unwrap.py tests/test_gpu_any_size.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2af_tests.log
python - > gpurun_out/r2af_time.log 2>&1 <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import style_transfer_based_holographic_imaging_b200 as pkg
from oracle import asm_oracle as ao
hg = pkg.Holo_Generator(ao.Optics()).cuda()
for b, n in [(5, 128), (64, 128), (64, 92), (8, 256)]:
    P = torch.rand(b, 1, n, n, device='cuda'); D = torch.rand(b, 1, 1, 1, device='cuda') * 0.5 + 0.3
    with torch.no_grad():
        a, p = hg(0.6, P, -D, return_field=True)
        for _ in range(2): pkg.unwrap(p)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): pkg.unwrap(p)
        torch.cuda.synchronize(); print(f"unwrap B={b} N={n}: {(time.perf_counter() - t) / 5 * 1e3:.3f} ms per call")
PY

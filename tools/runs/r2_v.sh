#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_any_size.py tests/test_gpu_unwrap.py tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/r2v_tests.log
python - > gpurun_out/r2v_time.log 2>&1 <<'PY'
import torch, time, sys
sys.path.insert(0, '.')
import style_transfer_based_holographic_imaging_b200 as pkg
from oracle import asm_oracle as ao
hg = pkg.Holo_Generator(ao.Optics()).cuda()
for b, n in [(64, 92), (64, 128), (5, 128)]:
    P = torch.rand(b, 1, n, n, device='cuda'); D = torch.rand(b, 1, 1, 1, device='cuda') * 0.5 + 0.3
    with torch.no_grad():
        for _ in range(3): hg(0.6, P, D)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): hg(0.6, P, D)
        torch.cuda.synchronize(); print(f"Holo_Generator B={b} N={n}: {(time.perf_counter() - t) / 10 * 1e3:.3f} ms per call")
        a, p = hg(0.6, P, -D, return_field=True)
        for _ in range(2): pkg.unwrap(p)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(5): pkg.unwrap(p)
        torch.cuda.synchronize(); print(f"unwrap B={b} N={n}: {(time.perf_counter() - t) / 5 * 1e3:.3f} ms per call")
PY

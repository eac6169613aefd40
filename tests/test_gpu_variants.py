"""GPU: the opt-in kernel variants of the FFT-size-1024 path (persistent dataflow kernel, cp.async row pipeline,
column-kernel flavours, generic 16-point kernels) must give the same answers as the default.  The variants are
selected by environment variables read once per process, so each runs in a subprocess."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
import style_transfer_based_holographic_imaging_b200 as pkg
from oracle import asm_oracle as ao
rng = np.random.default_rng(11)
worst = 0.0
for n, pad in [(1024, False), (512, True)]:
    b = 3
    O = (rng.standard_normal((b, 1, n, n)) + 1j * rng.standard_normal((b, 1, n, n))).astype(np.complex64)
    d = ((0.2 + 0.8 * rng.random((b, 1, 1, 1))) * 5e-3).astype(np.float32)
    x, z = torch.from_numpy(O).cuda(), torch.from_numpy(d).cuda()
    U = pkg.asm_forward_raw(x, z, 532e-9, 1.5e-6, pad)
    A = pkg.asm_adjoint_raw(x, z, 532e-9, 1.5e-6, pad)
    worst = max(worst, ao.rel_l2(U.cpu().numpy(), ao.asm(O, 532e-9, d, 1.5e-6, pad)),
                ao.rel_l2(A.cpu().numpy(), ao.asm_adjoint(O, 532e-9, d, 1.5e-6, pad)))
# autograd through the intensity model (amp/phase input, saved field, grad_z) on the 1024 FFT
hg = pkg.Holo_Generator(ao.Optics())
amp = (0.5 + 0.5 * rng.random((2, 1, 512, 512))).astype(np.float32)
ph = (6.28 * rng.random((2, 1, 512, 512))).astype(np.float32)
dd = np.array([0.4, 0.9], dtype=np.float32).reshape(2, 1, 1, 1)
w = rng.standard_normal((2, 1, 512, 512)).astype(np.float32)
A_, P_, D_ = (torch.from_numpy(t).cuda().requires_grad_(True) for t in (amp, ph, dd))
I = hg(A_, P_, D_)
gA, gP, gD = torch.autograd.grad(torch.sum(torch.from_numpy(w).cuda() * I), [A_, P_, D_])
ra, rp, rd = ao.holo_generator_vjp(amp, ph, dd, w, ao.Optics())
worst = max(worst, ao.rel_l2(I.detach().cpu().numpy(), ao.holo_generator(amp, ph, dd, ao.Optics())),
            ao.rel_l2(gA.cpu().numpy(), ra), ao.rel_l2(gP.cpu().numpy(), rp))
assert ao.rel_l2(gD.cpu().numpy().reshape(-1), rd) < 1e-3
print("WORST", worst)
assert worst < 1e-4
""" % ROOT

VARIANTS = {
    "default": {},
    "persistent_dataflow_kernel": {"ASM_B200_MEGA": "1"},
    "uniform_worker_flow_kernel": {"ASM_B200_FLOW": "1"},
    "uniform_worker_flow_kernel_tight_ring": {"ASM_B200_FLOW": "1", "ASM_B200_RING": "2", "ASM_B200_FLOW_RPT": "32"},
    "cp_async_row_pipeline": {"ASM_B200_ROWPIPE": "1"},
    "column_kernel_plain": {"ASM_B200_PIPE": "0"},
    "column_kernel_separate_landing_zone": {"ASM_B200_PIPE": "1"},
    "generic_16_point_kernels": {"ASM_B200_GENERIC10": "1"},
    "single_lane_small_ring": {"ASM_B200_LANES": "1", "ASM_B200_CHUNK_MB": "16"},
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_variant_matches_oracle(name):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    env = dict(os.environ)
    env.update(VARIANTS[name])
    r = subprocess.run([sys.executable, "-c", SNIPPET], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "WORST" in r.stdout

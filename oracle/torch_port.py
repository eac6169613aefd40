"""torch port of the reference's ASM op sequence, used ONLY as the timed baseline of bench.py when the reference's own
files are not staged (`cpu_baseline` leg, `--impl reference`, `--impl reference-cuda`) and cross-checked against
oracle/asm_oracle.py in tests.

TEST / BENCH INFRASTRUCTURE ONLY.  It keeps the reference's cost profile: the kz grid is rebuilt on the host in float64
for every call, repeated over the batch and moved to the field's device (utils/Angular_Spectrum_Method.py:13-26), the
transfer function is a complex128 exp (:29), the forward FFT runs in complex64 and the inverse in complex128 with
explicit fftshift / ifftshift copies (:38-48), and the intensity is |U|^2 cast to fp32 (utils/Forward_model.py:39).
`device` follows the input tensor (CPU = MKL FFT, CUDA = cuFFT), as in the reference (:25-26).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _kz_host(n: int, lamb: float, px: float, padded: bool, batch: int) -> torch.Tensor:
    m = 2 * n if padded else n
    ax = (np.arange(m) / 2 - n // 2) if padded else (np.arange(m) - n // 2)      # :13-17
    ax = ax / (n * px)                                                            # :19-20
    sq = ax ** 2
    g = 1 - lamb ** 2 * (sq[:, None] + sq[None, :])                               # :22
    kz = np.sqrt(np.where(g > 0, g, 0.0)) / lamb                                  # :23 (evanescent -> 0)
    return torch.from_numpy(np.repeat(kz[None, None], batch, axis=0))             # :23 repeat over batch


def asm_cpu(field: torch.Tensor, lamb: float, d: torch.Tensor, px: float, zero_padding: bool = False) -> torch.Tensor:
    b, _, n, _ = field.shape
    x = F.pad(field, (n // 2,) * 4, mode="replicate") if zero_padding else field   # :12
    kz = _kz_host(n, lamb, px, zero_padding, b).to(field.device)                   # :25-26 (H2D on every call)
    h = torch.exp(1j * 2 * math.pi * d * kz)                                       # :29 complex128
    spec = torch.fft.fftshift(torch.fft.fft2(x), dim=(-2, -1))                     # :38-42
    u = torch.fft.ifft2(torch.fft.ifftshift(h * spec, dim=(-2, -1)))               # :33, :44-48
    lo = (u.shape[-1] - n) // 2
    return u[:, :, lo:lo + n, lo:lo + n]                                           # :50-53


asm_torch = asm_cpu   # device-agnostic name


def forward_intensity_cpu(field, lamb, d, px, zero_padding=False) -> torch.Tensor:
    return torch.pow(torch.abs(asm_cpu(field, lamb, d, px, zero_padding)), 2).float()   # Forward_model.py:39


def adjoint_cpu(cot, lamb, d, px) -> torch.Tensor:
    """Unpadded adjoint = propagation by -d (exact)."""
    return asm_cpu(cot, lamb, -d, px, False)

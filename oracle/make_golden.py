"""Generate tests/golden/*.npz from the UNMODIFIED reference imported live (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/make_golden.py
  * tests/golden/mnist_holo.npz   -- the reference's bundled fixtures (test_data/*.pt, 20 batches x 5 samples,
    N=128): gt_phase, distance_content, distance_style and the stored hologram ``content_holo``.
    gt_amplitude is the constant 0.6 (asserted here) and is not stored.
  * tests/golden/ref_cases.npz    -- seeded inputs + outputs of the reference's ASM / Holo_Generator /
    Back_prop and the gradients PyTorch autograd derives from them, for the shapes/optics nothing in
    test_data pins (complex field, amp/phase, Back_prop, no-pad, evanescent optics, z up to 20 mm, negative z).
Seeds and shapes are fixed, so the files are reproducible with the same torch build.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from oracle.asm_oracle import Optics  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def mnist_fixtures():
    td = os.path.join(ref_import.CHECKOUT, "test_data")
    ld = lambda name, i: torch.load(os.path.join(td, f"test_{name}_{i}.pt"), map_location="cpu").numpy()
    phase, holo, dc, ds = [], [], [], []
    for i in range(20):
        amp = ld("gt_amplitude", i)
        assert np.all(amp == np.float32(0.6)), "gt_amplitude is expected to be the constant 0.6"
        phase.append(ld("gt_phase", i))
        holo.append(ld("content_holo", i))
        dc.append(ld("distance_content", i))
        ds.append(ld("distance_style", i))
    np.savez_compressed(os.path.join(OUT, "mnist_holo.npz"),
                        gt_phase=np.stack(phase).astype(np.float32),          # [20,5,1,128,128]
                        content_holo=np.stack(holo).astype(np.float32),       # [20,5,1,128,128]
                        distance_content=np.stack(dc).astype(np.float32),     # [20,5,1,1,1]
                        distance_style=np.stack(ds).astype(np.float32),
                        amplitude=np.float32(0.6))


def ref_cases():
    ASM, Holo_Generator, Back_prop = ref_import.load()
    g = torch.Generator().manual_seed(20261018)
    out = {}

    def rnd(*s):
        return torch.rand(*s, generator=g)

    def rndn(*s):
        return torch.randn(*s, generator=g)

    # ---- ASM, complex field in/out (utils/Angular_Spectrum_Method.py:7) ----
    cases = [  # name, B, N, pad, lamb, px, zmax
        ("asm_n32", 3, 32, False, 532e-9, 1.5e-6, 1e-3),
        ("asm_n32_pad", 3, 32, True, 532e-9, 1.5e-6, 1e-3),
        ("asm_n64_far", 2, 64, False, 532e-9, 1.5e-6, 20e-3),
        ("asm_n64_pad_far", 2, 64, True, 532e-9, 1.5e-6, 20e-3),
        ("asm_n64_evan", 2, 64, False, 532e-9, 0.2e-6, 50e-6),       # 55 % of bins evanescent -> kz clamp
        ("asm_n32_pad_evan", 2, 32, True, 532e-9, 0.2e-6, 50e-6),
        ("asm_n128_neg", 1, 128, False, 633e-9, 2.0e-6, -6e-3),      # negative z (back-focus)
        ("asm_n128_pad", 1, 128, True, 532e-9, 1.5e-6, 0.8e-3),
    ]
    for name, B, N, pad, lamb, px, zmax in cases:
        O = (rndn(B, 1, N, N) + 1j * rndn(B, 1, N, N)).to(torch.complex64)
        d = ((0.2 + 0.8 * rnd(B, 1, 1, 1)) * zmax).float()
        G = (rndn(B, 1, N, N) + 1j * rndn(B, 1, N, N)).to(torch.complex64)   # cotangent
        Oq = O.clone().requires_grad_(True)
        dq = d.clone().requires_grad_(True)
        U = ASM(Oq, lamb, dq, px, zero_padding=pad)
        # PyTorch convention: grad = d(Re<G,U>)/d(conj O) * 2 ... just take what autograd gives for the real
        # scalar L = Re sum(conj(G) * U); that is what a drop-in backward must reproduce.
        L = torch.real(torch.sum(torch.conj(G.to(U.dtype)) * U))
        gO, gd = torch.autograd.grad(L, [Oq, dq])
        out.update({f"{name}.O": O.numpy(), f"{name}.d": d.numpy(), f"{name}.G": G.numpy(),
                    f"{name}.U": U.detach().numpy().astype(np.complex64), f"{name}.gO": gO.numpy(),
                    f"{name}.gd": gd.numpy(),
                    f"{name}.meta": np.array([B, N, int(pad), lamb, px], dtype=np.float64)})

    # ---- Holo_Generator (utils/Forward_model.py:16-39) ----
    hg_cases = [  # name, B, N, optics kwargs, d range (normalised mm)
        ("hg_n32", 3, 32, dict(), (0.3, 0.9)),
        ("hg_n64_norm", 2, 64, dict(phase_normalize=2 * np.pi, distance_normalize=2.5, distance_normalize_constant=0.4), (-0.2, 0.6)),
        ("hg_n128_neg", 1, 128, dict(), (-0.8, -0.2)),
    ]
    for name, B, N, kw, (dlo, dhi) in hg_cases:
        args = Optics(**kw)
        hg = Holo_Generator(args)
        A = (0.5 + 0.5 * rnd(B, 1, N, N)).float()
        P = rnd(B, 1, N, N).float() * (1.0 if kw else 2 * np.pi)
        d = (dlo + (dhi - dlo) * rnd(B, 1, 1, 1)).float()
        W = rndn(B, 1, N, N).float()
        Aq, Pq, dq = A.clone().requires_grad_(True), P.clone().requires_grad_(True), d.clone().requires_grad_(True)
        I = hg(Aq, Pq, dq)
        gA, gP, gd = torch.autograd.grad(torch.sum(W * I), [Aq, Pq, dq])
        with torch.no_grad():
            amp, ph = hg(A, P, d, return_field=True)
            Uc = hg(A, P, d, complex_number=True)
        # gradients through the (abs, angle) outputs as well
        Wa, Wp = rndn(B, 1, N, N).float(), rndn(B, 1, N, N).float()
        a2, p2 = hg(Aq, Pq, dq, return_field=True)
        gA2, gP2, gd2 = torch.autograd.grad(torch.sum(Wa * a2) + torch.sum(Wp * p2), [Aq, Pq, dq])
        out.update({f"{name}.A": A.numpy(), f"{name}.P": P.numpy(), f"{name}.d": d.numpy(), f"{name}.W": W.numpy(),
                    f"{name}.I": I.detach().numpy(), f"{name}.amp": amp.numpy(), f"{name}.ph": ph.numpy(),
                    f"{name}.U": Uc.numpy().astype(np.complex64), f"{name}.gA": gA.numpy(), f"{name}.gP": gP.numpy(),
                    f"{name}.gd": gd.numpy(), f"{name}.Wa": Wa.numpy(), f"{name}.Wp": Wp.numpy(),
                    f"{name}.gA2": gA2.numpy(), f"{name}.gP2": gP2.numpy(), f"{name}.gd2": gd2.numpy(),
                    f"{name}.meta": np.array([B, N, args.wavelength, args.pixel_size, args.phase_normalize,
                                              args.distance_normalize, args.distance_normalize_constant])})

    # ---- Back_prop (utils/Forward_model.py:52-65) ----
    for name, B, N, kind, an in [("bp_n64_amp_pha", 2, 64, "amp_pha", 1.7), ("bp_n64_re_im", 2, 64, "real_imag", 0.5),
                                 ("bp_n32_amp_pha", 1, 32, "amp_pha", -1.0)]:
        args = Optics(amplitude_normalize=an, Holo_G_input=kind)
        bp = Back_prop(args)
        holo = (rnd(B, 1, N, N) * 1.5 + 0.05).float()
        d = (0.2 + 0.8 * rnd(B, 1, 1, 1)).float()
        with torch.no_grad():
            r = bp(holo, d)
        out.update({f"{name}.holo": holo.numpy(), f"{name}.d": d.numpy(), f"{name}.out": r.numpy().astype(np.float32),
                    f"{name}.meta": np.array([B, N, an, 1.0 if kind == "amp_pha" else 0.0])})

    np.savez_compressed(os.path.join(OUT, "ref_cases.npz"), **out)


if __name__ == "__main__":
    if not ref_import.available():
        sys.exit("reference checkout not present; golden vectors are generated in the build container only")
    os.makedirs(OUT, exist_ok=True)
    mnist_fixtures()
    ref_cases()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))

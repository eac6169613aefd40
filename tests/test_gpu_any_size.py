"""Field sizes that are not powers of two (matrix-product path, csrc/dft_any.cuh): the reference accepts any square N
unpadded and any even N with zero_padding=True (utils/Angular_Spectrum_Method.py:11-23); its MNIST loader makes 92 x 92
fields (utils/Data_loader.py:24: 28 + 2 * 32).  Oracle: oracle/asm_oracle.py (pinned to the live reference at these sizes
in tests/test_oracle.py::test_oracle_matches_reference_any_size)."""
import numpy as np
import pytest
import torch

from oracle import asm_oracle as ao

pytestmark = pytest.mark.gpu
LAMB, PX, TOL = 532e-9, 1.5e-6, 1e-4


def _field(rng, b, n):
    return (rng.standard_normal((b, 1, n, n)) + 1j * rng.standard_normal((b, 1, n, n))).astype(np.complex64)


@pytest.mark.parametrize("n,pad", [(92, True), (92, False), (100, True), (45, False), (184, False), (30, True), (250, True)])
def test_asm_forward_and_adjoint_any_size(n, pad):
    import style_transfer_based_holographic_imaging_b200 as pkg
    rng = np.random.default_rng(n)
    b = 3
    O, G = _field(rng, b, n), _field(rng, b, n)
    d = ((0.2 + 0.8 * rng.random((b, 1, 1, 1))) * 2e-3).astype(np.float32)
    x, g, z = torch.from_numpy(O).cuda(), torch.from_numpy(G).cuda(), torch.from_numpy(d).cuda()
    U = pkg.ASM(x, LAMB, z, PX, zero_padding=pad)
    A = pkg.asm_adjoint_raw(g, z, LAMB, PX, pad)
    e_u = ao.rel_l2(U.cpu().numpy(), ao.asm(O, LAMB, d, PX, pad))
    e_a = ao.rel_l2(A.cpu().numpy(), ao.asm_adjoint(G, LAMB, d, PX, pad))
    print(f"N={n} pad={pad}: forward {e_u:.2e} adjoint {e_a:.2e}")
    assert e_u < TOL and e_a < TOL
    # <A x, y> = <x, A^H y>
    lhs = torch.sum(U * torch.conj(g)).item()
    rhs = torch.sum(x * torch.conj(A)).item()
    assert abs(lhs - rhs) < 1e-4 * abs(lhs)


def test_mnist_loader_shape_holo_generator_and_grads():
    """92 x 92 (MNIST 28 + 2 * 32, utils/Data_loader.py:24-32): constant amplitude 0.6, padded, intensity + all gradients."""
    import style_transfer_based_holographic_imaging_b200 as pkg
    rng = np.random.default_rng(3)
    b, n = 4, 92
    amp = np.full((b, 1, n, n), 0.6, dtype=np.float32)
    ph = np.zeros((b, 1, n, n), dtype=np.float32)
    ph[:, :, 32:60, 32:60] = rng.random((b, 1, 28, 28)).astype(np.float32)
    d = (0.3 + 0.5 * rng.random((b, 1, 1, 1))).astype(np.float32)
    w = rng.standard_normal((b, 1, n, n)).astype(np.float32)
    args = ao.Optics()
    hg = pkg.Holo_Generator(args).cuda()
    A = torch.from_numpy(amp).cuda().requires_grad_(True)
    P = torch.from_numpy(ph).cuda().requires_grad_(True)
    D = torch.from_numpy(d).cuda().requires_grad_(True)
    I = hg(A, P, D)
    gA, gP, gD = torch.autograd.grad(torch.sum(torch.from_numpy(w).cuda() * I), [A, P, D])
    assert ao.rel_l2(I.detach().cpu().numpy(), ao.holo_generator(amp, ph, d, args)) < TOL
    ra, rp, rd = ao.holo_generator_vjp(amp, ph, d, w, args)
    assert ao.rel_l2(gA.cpu().numpy(), ra) < TOL and ao.rel_l2(gP.cpu().numpy(), rp) < TOL
    assert ao.rel_l2(gD.cpu().numpy().reshape(-1), rd) < 1e-3
    # the loaders' pair call with a scalar amplitude (IN_CONST_AMP_PHASE)
    with torch.no_grad():
        ha, hb = hg.forward_pair(0.6, P.detach(), torch.flip(P.detach(), dims=[3]), D.detach(), D.detach() * 0.5)
    assert ao.rel_l2(ha.cpu().numpy(), ao.holo_generator(amp, ph, d, args)) < TOL
    assert ao.rel_l2(hb.cpu().numpy(), ao.holo_generator(amp, ph[:, :, :, ::-1].copy(), d * 0.5, args)) < TOL


def test_back_prop_any_size():
    import style_transfer_based_holographic_imaging_b200 as pkg
    rng = np.random.default_rng(4)
    b, n = 2, 92
    holo = (0.2 + rng.random((b, 1, n, n))).astype(np.float32)
    d = (0.3 + 0.5 * rng.random((b, 1, 1, 1))).astype(np.float32)
    for mode in ("amp_pha", "real_imag"):
        args = ao.Optics(amplitude_normalize=1.7, Holo_G_input=mode)
        out = pkg.Back_prop(args).cuda()(torch.from_numpy(holo).cuda(), torch.from_numpy(d).cuda())
        ref = ao.back_prop(holo, d, args)
        if mode == "amp_pha":   # compare the complex field (the angle is ill-conditioned where |U| ~ 0)
            got = out[:, :1].cpu().numpy() * np.exp(1j * out[:, 1:].cpu().numpy())
            want = ref[:, :1] * np.exp(1j * ref[:, 1:])
            assert ao.rel_l2(got, want) < TOL
        else:
            assert ao.rel_l2(out.cpu().numpy(), ref) < TOL

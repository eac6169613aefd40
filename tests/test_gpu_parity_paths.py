"""GPU parity tests (-m gpu) for the exact kernel variants that are timed and shipped:

  (a) the aligned-buffer OUT_INTENSITY / OUT_COMPLEX fast path (TMA bulk-copy row kernels) at FFT 1024, and
      Holo_Generator under no_grad at N = 512 (bulk amplitude/phase in -> bulk intensity out);
  (b) the adjoint at FFT 2048 / 4096, padded and unpadded;
  (c) Back_prop (both modes), return_field=True and real-input ASM at N = 256 and N = 1024;
  (d) bench-scale batches (B = 512 at 1024^2, B = 4096 at 256^2): samples from different chunks / lanes vs the oracle.

Everything goes Python -> ctypes -> C ABI (include/asm_b200.h).  Reference lines pinned: utils/Forward_model.py:39,
:52-65, utils/Angular_Spectrum_Method.py:7-36.  Tolerance 1e-4 relative L2 (north_star); per-shape errors printed.
"""
import numpy as np
import pytest
import torch

from oracle import asm_oracle as ao

pytestmark = pytest.mark.gpu

TOL = 1e-4
LAMB, PX = 532e-9, 1.5e-6


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import style_transfer_based_holographic_imaging_b200 as m
    m._lib.load()
    return m


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _field(rng, b, n):
    return (rng.standard_normal((b, 1, n, n)) + 1j * rng.standard_normal((b, 1, n, n))).astype(np.complex64)


def _phase_err(a, b):
    return ao.rel_l2(np.exp(1j * a.astype(np.float64)), np.exp(1j * b.astype(np.float64)))


# ---------------------------------------------------------------------------------------------------
# (a) the headline call: aligned complex64 in, aligned fp32 |U|^2 out (what bench.py times)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,pad", [(1024, False), (512, True), (256, False), (2048, False)])
def test_headline_intensity_call_aligned_buffers(pkg, n, pad):
    from style_transfer_based_holographic_imaging_b200 import _lib as L
    rng = np.random.default_rng(1000 + n + int(pad))
    b = 3 if n <= 1024 else 1
    O = _field(rng, b, n)
    d = ((0.2 + 0.8 * rng.random((b, 1, 1, 1))) * 6e-3).astype(np.float32)
    x, z = _dev(O), _dev(d)
    I = torch.empty(b, 1, n, n, device="cuda", dtype=torch.float32)
    A = torch.empty(b, 1, n, n, device="cuda", dtype=torch.complex64)
    assert x.data_ptr() % 16 == 0 and I.data_ptr() % 16 == 0 and A.data_ptr() % 16 == 0
    out = pkg.asm_forward_raw(x, z, LAMB, PX, pad, out_mode=L.OUT_INTENSITY, out=I)     # bench.py's forward call
    assert out.data_ptr() == I.data_ptr()
    adj = pkg.asm_adjoint_raw(x, z, LAMB, PX, pad, out=A)                                # bench.py's adjoint call
    e_i = ao.rel_l2(I.cpu().numpy(), np.abs(ao.asm(O, LAMB, d, PX, pad)) ** 2)
    e_a = ao.rel_l2(adj.cpu().numpy(), ao.asm_adjoint(O, LAMB, d, PX, pad))
    print(f"headline call N={n} pad={pad}: |U|^2 {e_i:.3e}  adjoint {e_a:.3e}")
    assert e_i < TOL and e_a < TOL


@pytest.mark.parametrize("n", [128, 256, 512, 1024])
def test_holo_generator_no_grad_fast_path(pkg, n):
    """Holo_Generator under torch.no_grad(): amplitude/phase planes in -> |U|^2 out without a saved field
    (utils/Forward_model.py:16-39); N = 512 is the FFT-1024 bulk path, N = 1024 the FFT-2048 warp-pair path (intensity
    and complex outputs) and the generic row kernels around the 2 x 1024 column kernel (abs / angle outputs)."""
    rng = np.random.default_rng(2000 + n)
    b = 3 if n < 1024 else 2
    amp = (0.5 + 0.5 * rng.random((b, 1, n, n))).astype(np.float32)
    ph = (2 * np.pi * rng.random((b, 1, n, n))).astype(np.float32)
    d = (0.3 + 0.6 * rng.random((b, 1, 1, 1))).astype(np.float32)
    args = ao.Optics(phase_normalize=0.7, distance_normalize=1.5, distance_normalize_constant=0.05)
    hg = pkg.Holo_Generator(args)
    with torch.no_grad():
        I = hg(_dev(amp), _dev(ph), _dev(d))
        a2, p2 = hg(_dev(amp), _dev(ph), _dev(d), return_field=True)
        U = hg(_dev(amp), _dev(ph), _dev(d), complex_number=True)
    ref_u = ao.holo_generator(amp, ph, d, args, complex_number=True)
    e_i = ao.rel_l2(I.cpu().numpy(), ao.holo_generator(amp, ph, d, args))
    e_u = ao.rel_l2(U.cpu().numpy(), ref_u)
    e_abs = ao.rel_l2(a2.cpu().numpy(), np.abs(ref_u))
    e_ang = _phase_err(p2.cpu().numpy(), np.angle(ref_u))
    print(f"Holo_Generator no_grad N={n}: I {e_i:.3e} U {e_u:.3e} |U| {e_abs:.3e} angle {e_ang:.3e}")
    assert max(e_i, e_u, e_abs, e_ang) < TOL


# ---------------------------------------------------------------------------------------------------
# (b) adjoint at the large transforms
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,pad", [(2048, False), (1024, True), (4096, False), (2048, True)])
def test_adjoint_large_vs_oracle(pkg, n, pad):
    rng = np.random.default_rng(3000 + n + int(pad))
    G = _field(rng, 1, n)
    d = np.array([[[[5.1e-3]]]], dtype=np.float32)
    A = pkg.asm_adjoint_raw(_dev(G), _dev(d), LAMB, PX, pad)
    e = ao.rel_l2(A.cpu().numpy(), ao.asm_adjoint(G, LAMB, d, PX, pad))
    print(f"adjoint N={n} pad={pad}: rel-L2 {e:.3e}")
    assert e < TOL


# ---------------------------------------------------------------------------------------------------
# (c) Back_prop, return_field and real-input ASM above N = 128
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [256, 1024])
@pytest.mark.parametrize("mode", ["amp_pha", "real_imag"])
def test_back_prop_large(pkg, n, mode):
    rng = np.random.default_rng(4000 + n)
    b = 2
    holo = (0.2 + rng.random((b, 1, n, n))).astype(np.float32)
    d = (0.3 + 0.5 * rng.random((b, 1, 1, 1))).astype(np.float32)
    args = ao.Optics(amplitude_normalize=1.7, Holo_G_input=mode, distance_normalize=2.0, distance_normalize_constant=0.1)
    bp = pkg.Back_prop(args)
    with torch.no_grad():
        out = bp(_dev(holo), _dev(d)).cpu().numpy()
    ref = ao.back_prop(holo, d, args)
    assert out.shape == ref.shape == (b, 2, n, n)
    if mode == "amp_pha":
        e0, e1 = ao.rel_l2(out[:, :1], ref[:, :1]), _phase_err(out[:, 1:], ref[:, 1:])
    else:
        e0, e1 = ao.rel_l2(out[:, :1], ref[:, :1]), ao.rel_l2(out[:, 1:], ref[:, 1:])
    print(f"Back_prop N={n} {mode}: {e0:.3e} {e1:.3e}")
    assert e0 < TOL and e1 < TOL


@pytest.mark.parametrize("n,pad", [(256, False), (1024, False), (512, True), (2048, False), (1024, True)])
def test_real_input_asm_large(pkg, n, pad):
    rng = np.random.default_rng(5000 + n)
    b = 2 if n < 2048 else 1
    x = rng.standard_normal((b, 1, n, n)).astype(np.float32)
    d = ((0.2 + 0.8 * rng.random((b, 1, 1, 1))) * 4e-3).astype(np.float32)
    with torch.no_grad():
        U = pkg.ASM(_dev(x), LAMB, _dev(d), PX, zero_padding=pad)
    e = ao.rel_l2(U.cpu().numpy(), ao.asm(x.astype(np.complex64), LAMB, d, PX, pad))
    print(f"real-input ASM N={n} pad={pad}: {e:.3e}")
    assert e < TOL


@pytest.mark.parametrize("n", [256, 512, 1024])
def test_intensity_backward_large(pkg, n):
    """Training path above N = 128: saved field, grad_A / grad_phase (OUT_GRAD_AP epilogue) and grad_d."""
    rng = np.random.default_rng(6000 + n)
    b = 2
    amp = (0.5 + 0.5 * rng.random((b, 1, n, n))).astype(np.float32)
    ph = (2 * np.pi * rng.random((b, 1, n, n))).astype(np.float32)
    d = (0.3 + 0.6 * rng.random((b, 1, 1, 1))).astype(np.float32)
    w = rng.standard_normal((b, 1, n, n)).astype(np.float32)
    args = ao.Optics()
    hg = pkg.Holo_Generator(args)
    A, P, D = (_dev(t).requires_grad_(True) for t in (amp, ph, d))
    I = hg(A, P, D)
    gA, gP, gD = torch.autograd.grad(torch.sum(_dev(w) * I), [A, P, D])
    ra, rp, rd = ao.holo_generator_vjp(amp, ph, d, w, args)
    e = (ao.rel_l2(I.detach().cpu().numpy(), ao.holo_generator(amp, ph, d, args)), ao.rel_l2(gA.cpu().numpy(), ra),
         ao.rel_l2(gP.cpu().numpy(), rp), ao.rel_l2(gD.cpu().numpy().reshape(-1), rd))
    print(f"intensity fwd/bwd N={n}: I {e[0]:.3e} grad_A {e[1]:.3e} grad_phase {e[2]:.3e} grad_d {e[3]:.3e}")
    assert max(e[:3]) < TOL and e[3] < 1e-3


# ---------------------------------------------------------------------------------------------------
# (d) bench-scale batches: samples from different chunks / lanes against the oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,b,pad", [(1024, 512, False), (256, 4096, False), (128, 2048, True)])
def test_bench_scale_batch_samples_vs_oracle(pkg, n, b, pad):
    from style_transfer_based_holographic_imaging_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(4321)
    O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
    z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
    I = torch.empty(b, 1, n, n, device="cuda", dtype=torch.float32)
    pkg.asm_forward_raw(O, z, LAMB, PX, pad, out_mode=L.OUT_INTENSITY, out=I)
    A = pkg.asm_adjoint_raw(O, z, LAMB, PX, pad)
    torch.cuda.synchronize()
    idx = [0, 7, b // 3 + 1, b // 2 + 5, b - 10, b - 1]                 # first / middle / last chunks, all lanes
    o = O[idx].cpu().numpy()
    zz = z[idx].cpu().numpy()
    e_i = ao.rel_l2(I[idx].cpu().numpy(), np.abs(ao.asm(o, LAMB, zz, PX, pad)) ** 2)
    e_a = ao.rel_l2(A[idx].cpu().numpy(), ao.asm_adjoint(o, LAMB, zz, PX, pad))
    print(f"bench-scale N={n} B={b} pad={pad}: |U|^2 {e_i:.3e}  adjoint {e_a:.3e}  (samples {idx})")
    assert e_i < TOL and e_a < TOL
    # every sample: energy conservation (|H| = 1 at the default optics) as a size-independent checksum
    if not pad:
        e_in = (O.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64)
        assert torch.max(torch.abs(I.sum(dim=(1, 2, 3), dtype=torch.float64) / e_in - 1)).item() < 1e-5
        e_ad = (A.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64)
        assert torch.max(torch.abs(e_ad / e_in - 1)).item() < 1e-5


# ---------------------------------------------------------------------------------------------------
# broadcast inputs and the constant-amplitude input mode
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [64, 512])
def test_broadcast_inputs_and_constant_amplitude(pkg, n):
    """``amplitude*torch.exp(1j*phase)`` (utils/Forward_model.py:22) broadcasts: a [1,1,N,N] phase shared by the batch,
    a [1,1,N,N] amplitude against a [B,1,N,N] phase, and a scalar amplitude (python float / one-element tensor, the
    loaders' 0.6 of utils/Data_loader.py:25) must all work, with gradients reduced to the input shapes."""
    rng = np.random.default_rng(7000 + n)
    b = 3
    args = ao.Optics()
    hg = pkg.Holo_Generator(args)
    amp = (0.5 + 0.5 * rng.random((b, 1, n, n))).astype(np.float32)
    ph1 = (2 * np.pi * rng.random((1, 1, n, n))).astype(np.float32)
    d = (0.3 + 0.6 * rng.random((b, 1, 1, 1))).astype(np.float32)
    w = rng.standard_normal((b, 1, n, n)).astype(np.float32)
    # (1) broadcast phase that requires grad
    A, P, D = _dev(amp).requires_grad_(True), _dev(ph1).requires_grad_(True), _dev(d)
    I = hg(A, P, D)
    gA, gP = torch.autograd.grad(torch.sum(_dev(w) * I), [A, P])
    phb = np.broadcast_to(ph1, amp.shape)
    ra, rp, _ = ao.holo_generator_vjp(amp, phb, d, w, args)
    assert gP.shape == P.shape and gA.shape == A.shape
    e = (ao.rel_l2(I.detach().cpu().numpy(), ao.holo_generator(amp, phb, d, args)), ao.rel_l2(gA.cpu().numpy(), ra),
         ao.rel_l2(gP.cpu().numpy(), rp.sum(axis=0, keepdims=True)))
    print(f"broadcast phase N={n}: I {e[0]:.3e} grad_A {e[1]:.3e} grad_phase(summed) {e[2]:.3e}")
    assert max(e) < TOL
    # (2) broadcast amplitude [1,1,N,N] against a batched phase
    ph = (2 * np.pi * rng.random((b, 1, n, n))).astype(np.float32)
    with torch.no_grad():
        I2 = hg(_dev(amp[:1]), _dev(ph), D)
    assert ao.rel_l2(I2.cpu().numpy(), ao.holo_generator(np.broadcast_to(amp[:1], ph.shape), ph, d, args)) < TOL
    # (3) scalar amplitude: python float, 0-dim tensor, [1,1,1,1] tensor -> IN_CONST_AMP_PHASE
    ref = ao.holo_generator(np.full_like(ph, 0.6), ph, d, args)
    with torch.no_grad():
        for a in (0.6, torch.tensor(0.6, device="cuda"), torch.full((1, 1, 1, 1), 0.6, device="cuda")):
            I3 = hg(a, _dev(ph), D)
            assert ao.rel_l2(I3.cpu().numpy(), ref) < TOL
        amp3, ang3 = hg(0.6, _dev(ph), D, return_field=True)
    ref_u = ao.holo_generator(np.full_like(ph, 0.6), ph, d, args, complex_number=True)
    assert ao.rel_l2(amp3.cpu().numpy(), np.abs(ref_u)) < TOL and _phase_err(ang3.cpu().numpy(), np.angle(ref_u)) < TOL
    # gradients through a scalar amplitude and the phase
    a0 = torch.tensor(0.6, device="cuda", requires_grad=True)
    P3 = _dev(ph).requires_grad_(True)
    g0, gP3 = torch.autograd.grad(torch.sum(_dev(w) * hg(a0, P3, D)), [a0, P3])
    ra3, rp3, _ = ao.holo_generator_vjp(np.full_like(ph, 0.6), ph, d, w, args)
    assert g0.shape == a0.shape and abs(g0.item() - ra3.sum()) < 1e-3 * abs(ra3.sum()) + 1e-3 * np.abs(ra3).sum() / ra3.size ** 0.5
    assert ao.rel_l2(gP3.cpu().numpy(), rp3) < TOL
    # (4) the loader pair with a constant amplitude: no cat, no amplitude plane
    ph_b = (2 * np.pi * rng.random((b, 1, n, n))).astype(np.float32)
    d_b = (0.3 + 0.6 * rng.random((b, 1, 1, 1))).astype(np.float32)
    ha, hb = hg.forward_pair(0.6, _dev(ph), _dev(ph_b), D, _dev(d_b))
    assert ha.dtype == torch.float32 and not ha.requires_grad
    assert ao.rel_l2(ha.cpu().numpy(), ref) < TOL
    assert ao.rel_l2(hb.cpu().numpy(), ao.holo_generator(np.full_like(ph_b, 0.6), ph_b, d_b, args)) < TOL


def test_repeated_calls_replay_the_cached_launch_sequence(pkg):
    """Small transforms are launch bound: from the second identical call on (same buffers, same arguments) the library
    replays the call's launch sequence as a CUDA graph.  Replays must give the bits of the first call, follow changed
    input DATA (the graph holds pointers, not values), and a call with other arguments must not hit the cache."""
    from style_transfer_based_holographic_imaging_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(99)
    b, n = 640, 256
    O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
    z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 1e-3).float()
    I = torch.empty(b, 1, n, n, device="cuda", dtype=torch.float32)
    n0 = pkg._lib.load().asm_b200_launch_count()
    pkg.asm_forward_raw(O, z, LAMB, PX, False, out_mode=L.OUT_INTENSITY, out=I)
    first = I.clone()
    per_call = pkg._lib.load().asm_b200_launch_count() - n0
    for _ in range(4):                                                   # sighting 2 captures, 3+ replay
        I.zero_()
        pkg.asm_forward_raw(O, z, LAMB, PX, False, out_mode=L.OUT_INTENSITY, out=I)
        assert torch.equal(I, first)
    assert pkg._lib.load().asm_b200_launch_count() - n0 == 5 * per_call     # replays count their kernels
    idx = [0, 333, b - 1]
    ref = np.abs(ao.asm(O[idx].cpu().numpy(), LAMB, z[idx].cpu().numpy(), PX, False)) ** 2
    assert ao.rel_l2(I[idx].cpu().numpy(), ref) < TOL
    O.mul_(0.5)                                                          # new data in the same buffers
    pkg.asm_forward_raw(O, z, LAMB, PX, False, out_mode=L.OUT_INTENSITY, out=I)
    assert ao.rel_l2(I[idx].cpu().numpy(), 0.25 * ref) < TOL
    z2 = (z * 1.5).contiguous()                                          # other arguments: a different sequence
    pkg.asm_forward_raw(O, z2, LAMB, PX, False, out_mode=L.OUT_INTENSITY, out=I)
    ref2 = np.abs(ao.asm(O[idx].cpu().numpy(), LAMB, z2[idx].cpu().numpy(), PX, False)) ** 2
    assert ao.rel_l2(I[idx].cpu().numpy(), ref2) < TOL


def test_fft1024_repeated_calls_bit_identical(pkg):
    """Regression for a cross-proxy hazard found in round 2: the TMA bulk-copy row kernels re-request their landing line
    right after reading it; without a generic->async proxy fence a shared-memory load still queued behind a co-resident
    CTA's traffic could see the NEXT row (about one wrong row in 1e4 with two row CTAs per SM and three chunks in
    flight).  Repeated calls must be bit-identical and every sample must conserve energy (|H| = 1)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    n, b = 1024, 48
    O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
    z = (0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3
    e_in = (O.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64)
    first = None
    for rep in range(12):
        U = pkg.asm_forward_raw(O, z, LAMB, PX, False)
        dev = ((U.abs() ** 2).sum(dim=(1, 2, 3), dtype=torch.float64) / e_in - 1).abs().max().item()
        assert dev < 2e-6, (rep, dev)
        if first is None:
            first = U.clone()
        else:
            assert torch.equal(U, first), rep


@pytest.mark.parametrize("n,b,pad", [(1024, 40, False), (512, 64, True), (2048, 6, False), (1024, 6, True),
                                     (128, 96, True), (256, 64, True), (64, 128, True), (2048, 2, True)])
def test_repeated_calls_bit_identical_all_tma_paths(pkg, n, b, pad):
    """The TMA / mbarrier pipelines (k32t: landing = exchange = tile region, tensor stores and loads of the tiles, lines
    re-requested into the exchange line; k64: warp-pair rows; generic kernels: TMA column slabs) must be free of
    shared-memory races and of order-dependent sums (the folds of the padded adjoint are fixed-order shuffle reductions,
    not floating-point atomics): ten repetitions of the forward |U|^2, the forward field and the adjoint give the bits of
    the first, whatever the CTA / lane interleaving."""
    from style_transfer_based_holographic_imaging_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(7 + n + int(pad))
    O = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
    G = torch.view_as_complex(torch.randn(b, 1, n, n, 2, device="cuda", generator=g))
    z = ((0.2 + 0.8 * torch.rand(b, 1, 1, 1, device="cuda", generator=g)) * 6e-3).float()
    first = None
    for rep in range(10):
        I = pkg.asm_forward_raw(O, z, LAMB, PX, pad, out_mode=L.OUT_INTENSITY)
        U = pkg.asm_forward_raw(O, z, LAMB, PX, pad)
        A = pkg.asm_adjoint_raw(G, z, LAMB, PX, pad)
        if first is None:
            first = (I.clone(), U.clone(), A.clone())
            idx = [0, b // 2, b - 1]                                 # and the first result is the right one
            ref = ao.asm(O[idx].cpu().numpy(), LAMB, z[idx].cpu().numpy(), PX, pad)
            assert ao.rel_l2(U[idx].cpu().numpy(), ref) < TOL and ao.rel_l2(I[idx].cpu().numpy(), np.abs(ref) ** 2) < TOL
            assert ao.rel_l2(A[idx].cpu().numpy(), ao.asm_adjoint(G[idx].cpu().numpy(), LAMB, z[idx].cpu().numpy(), PX, pad)) < TOL
        else:
            assert torch.equal(I, first[0]) and torch.equal(U, first[1]) and torch.equal(A, first[2]), rep

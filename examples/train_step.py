"""a physics-loss training step around the drop-in forward model.

The reference ships no training script (and its AdaIN/VGG network is out of scope), so this uses a small stand-in
generator with the same interface as `Net.forward` (net.py:199-226): hologram -> (amplitude, phase) + a distance
head -> d.  What it exercises is OUR path inside a real optimisation step: Holo_Generator (intensity) forward,
and its backward w.r.t. amplitude, phase AND the predicted distance (utils/Forward_model.py:16-39 through autograd).

  python examples/train_step.py [--batch 64] [--size 256] [--steps 20]
  torchrun --nproc-per-node 8 examples/train_step.py      (DDP all-reduces the NETWORK grads; the ASM op needs no collective)
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import style_transfer_based_holographic_imaging_b200 as asm  # noqa: E402


class Args:
    wavelength, pixel_size = 532e-9, 1.5e-6
    phase_normalize, distance_normalize, distance_normalize_constant = 1.0, 1.0, 0.0


class TinyGenerator(nn.Module):
    """Stand-in for decoder/decoder_ph + Distance_G (net.py:33-74, :266-308): two maps and one scalar per sample."""

    def __init__(self, ch=16):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(1, ch, 3, padding=1), nn.ReLU(), nn.Conv2d(ch, ch, 3, padding=1), nn.ReLU())
        self.amp = nn.Conv2d(ch, 1, 3, padding=1)
        self.ph = nn.Conv2d(ch, 1, 3, padding=1)
        self.dist = nn.Sequential(nn.Linear(2 * ch, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())

    def forward(self, holo_amp):
        f = self.body(holo_amp)
        stats = torch.cat([f.mean(dim=(2, 3)), f.std(dim=(2, 3))], dim=1)
        return torch.sigmoid(self.amp(f)), torch.pi * torch.tanh(self.ph(f)), (0.2 + 0.8 * self.dist(stats)).view(-1, 1, 1, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(local)
    fwd_model = asm.Holo_Generator(Args()).to(dev)
    net = TinyGenerator().to(dev)
    if world > 1:
        net = nn.parallel.DistributedDataParallel(net, device_ids=[local])
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    b = a.batch // world
    # synthetic "content" holograms made by the forward model itself, as utils/Data_loader.py:24-32 does
    with torch.no_grad():
        gt_amp = torch.full((b, 1, a.size, a.size), 0.6, device=dev)
        gt_ph = torch.rand(b, 1, a.size, a.size, device=dev)
        d_true = 0.4 + 0.4 * torch.rand(b, 1, 1, 1, device=dev)
        target = torch.sqrt(fwd_model(gt_amp, gt_ph, d_true))

    def step():
        amp, ph, d = net(target)
        holo = fwd_model(amp, ph, d)                      # |F(A e^{i phi}, d)|^2 with a LEARNED distance
        loss = F.l1_loss(torch.sqrt(holo + 1e-8), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    losses = [step() for _ in range(a.steps)]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    # share of the step spent in the ASM op (forward + adjoint + derivative propagation), measured separately
    amp, ph, d = (t.detach().requires_grad_(True) for t in net(target))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(a.steps):
        holo = fwd_model(amp, ph, d)
        torch.autograd.grad(holo.sum(), [amp, ph, d])
    torch.cuda.synchronize()
    dt_asm = (time.perf_counter() - t1) / a.steps
    if int(os.environ.get("RANK", "0")) == 0:
        print(f"train step: global batch {a.batch} @ {a.size}^2 on {world} GPU(s): {dt * 1e3:.2f} ms/step "
              f"({a.batch / dt:.0f} samples/s), loss {losses[0].item():.4f} -> {losses[-1].item():.4f}; "
              f"ASM fwd+bwd (amp, phase, distance grads) {dt_asm * 1e3:.2f} ms = {100 * dt_asm / dt:.0f}% of the step")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

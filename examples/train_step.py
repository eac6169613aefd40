"""SURVEY.md section 8(f) row 1 / BASELINE.json configs[4]: the style-transfer training step (AdaIN network + ASM physics
loss) around the drop-in forward model, batch 64 at 256^2, data-parallel over the GPUs of one box.

The reference ships no training script; the step is reconstructed from Figures/training.png, `Net.forward`
(net.py:199-226) and `mnist_loader` (utils/Data_loader.py:10-36):

  synthesis (no grad)   A_s = sqrt(F(0.6 e^{i phi_s}, d_s)),  A_c = sqrt(F(0.6 e^{i phi_c}, d_c))     Data_loader.py:24-32
  network               loss_c, loss_s, A_t, phi_t, _, d~_c, d~_s = Net(A_c, A_s, 1.0, True, True)      net.py:199-226
  physics loss          L_phy = | A_c - sqrt(F(A_t e^{i phi_t}, d~_c - d~_s)) |_1   (learned distance -> grad_d is on the path)
  step                  (loss_c + w_s loss_s + w_p L_phy).backward(); Adam on decoder + Distance_G

The network is the REFERENCE'S OWN `net.Net` / `net.Distance_G` / `net.vgg[:31]` / `net.decoder` with random weights
(the trained blobs are not in the checkout), imported from the read-only checkout or from the copy staged by
`__graft_entry__.build()` under baseline/_ref (git-ignored).  If neither exists a small stand-in generator with the
same interface is used and the output says so.  `--forward reference` swaps in the reference's torch.fft
`Holo_Generator` (the "before" measurement); the default is the B200 drop-in.  DDP all-reduces the NETWORK gradients;
the ASM op itself needs no collective.

  python examples/train_step.py [--batch 64] [--size 256] [--steps 10] [--forward ours|reference]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
"""
import argparse
import copy
import json
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import warnings  # noqa: E402

import style_transfer_based_holographic_imaging_b200 as asm  # noqa: E402

warnings.filterwarnings("ignore", message="input's size at dim=0 does not match num_features")   # Distance_G's InstanceNorm1d on [B, 1024]


class Args:
    wavelength, pixel_size = 532e-9, 1.5e-6
    phase_normalize, distance_normalize, distance_normalize_constant = 1.0, 1.0, 0.0


class TinyNet(nn.Module):
    """Stand-in with `Net.forward`'s return signature, used only when the reference's net.py is not available."""

    def __init__(self, ch=16):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(1, ch, 3, padding=1), nn.ReLU(), nn.Conv2d(ch, ch, 3, padding=1), nn.ReLU())
        self.out = nn.Conv2d(ch, 2, 3, padding=1)
        self.dist = nn.Sequential(nn.Linear(2 * ch, 64), nn.ReLU(), nn.Linear(64, 1), nn.Sigmoid())

    def forward(self, content, style, alpha=1.0, field_retrieval=True, unkonwn_distance=True):
        fc, fs = self.body(content), self.body(style)
        g = self.out(fc)
        stat = lambda f: torch.cat([f.mean(dim=(2, 3)), f.std(dim=(2, 3))], dim=1)  # noqa: E731
        dc, ds = self.dist(stat(fc)), self.dist(stat(fs))
        zero = g.sum() * 0
        return zero, zero, g[:, :1], g[:, 1:], None, dc, ds


def build_network(dev):
    """The reference's Net with random-init weights, as test_field_retrieval_mnist.py:76-93 assembles it."""
    try:
        from oracle import ref_import
        ref_net = ref_import.load_net()
    except Exception as e:  # reference not staged on this box
        return TinyNet().to(dev), f"stand-in generator (reference net.py unavailable: {type(e).__name__})"
    decoder = copy.deepcopy(ref_net.decoder)
    decoder_ph = copy.deepcopy(ref_net.decoder)
    distance_g = ref_net.Distance_G()
    vgg = nn.Sequential(*list(copy.deepcopy(ref_net.vgg).children())[:31])
    network = ref_net.Net(vgg, decoder, decoder_ph, distance_g)
    for q in network.decoder_ph.parameters():      # Net.forward (net.py:199-226) never calls decoder_ph: keep DDP's
        q.requires_grad_(False)                    # reducer from waiting for gradients that do not exist
    return network.to(dev), "reference net.Net(vgg[:31], decoder, decoder_ph, Distance_G), random init"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64, help="global batch")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--forward", default="ours", choices=["ours", "reference"])
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)
    if a.forward == "ours":
        fwd_model = asm.Holo_Generator(Args()).to(dev)
    else:
        from oracle import ref_import
        fwd_model = ref_import.load()[1](Args()).to(dev)
    network, net_kind = build_network(dev)
    network.train()
    model = nn.parallel.DistributedDataParallel(network, device_ids=[local]) if world > 1 else network
    params = [p for p in network.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4)
    b, n = a.batch // world, a.size
    w_s, w_p = 10.0, 1.0

    def synthesize():
        """utils/Data_loader.py:24-32: constant amplitude 0.6, MNIST-like phase zero-padded by n/4, two distance lists."""
        with torch.no_grad():
            core = n // 2
            ph_s = F.pad(torch.rand(b, 1, core, core, device=dev), (n // 4,) * 4)
            ph_c = F.pad(torch.rand(b, 1, core, core, device=dev), (n // 4,) * 4)
            d_s = 0.2 + 0.8 * torch.rand(b, 1, 1, 1, device=dev)
            d_c = 0.2 + 0.8 * torch.rand(b, 1, 1, 1, device=dev)
            if a.forward == "ours":
                h_s, h_c = fwd_model.forward_pair(0.6, ph_s, ph_c, d_s, d_c)       # one scalar amplitude, no cat
            else:
                amp = torch.ones_like(ph_s) * 0.6
                h_s, h_c = fwd_model(amp, ph_s, d_s).float().detach(), fwd_model(amp, ph_c, d_c).float().detach()
            return torch.sqrt(h_s), torch.sqrt(h_c)

    def step(sync=True):
        style, content = synthesize()
        ctx = model.no_sync() if (world > 1 and not sync) else torch.enable_grad()
        with ctx:
            loss_c, loss_s, g_t, g_t_phase, _, d_c, d_s = model(content, style, 1.0, True, True)
            d = (d_c - d_s).view(-1, 1, 1, 1)
            holo = fwd_model(g_t, g_t_phase, d)                     # |F(A_t e^{i phi_t}, d~_c - d~_s)|^2, learned distance
            loss_p = F.l1_loss(torch.sqrt(holo.float() + 1e-8), content)
            loss = loss_c + w_s * loss_s + w_p * loss_p
            opt.zero_grad(set_to_none=True)
            loss.backward()
        opt.step()
        return loss

    def timed(fn, k):
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = [fn() for _ in range(k)]
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / k], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(3):
        step()
    ms_step, losses = timed(step, a.steps)
    ms_nosync = timed(lambda: step(sync=False), a.steps)[0] if world > 1 else ms_step
    # the ASM share, measured separately on this step's shapes: 2 syntheses (no grad) + 1 physics-loss forward + backward
    style, content = synthesize()
    with torch.no_grad():
        _, _, g_t, g_t_phase, _, d_c, d_s = network(content, style, 1.0, True, True)
    amp, ph, d = (t.detach().float().requires_grad_(True) for t in (g_t, g_t_phase, (d_c - d_s).view(-1, 1, 1, 1)))

    def asm_part():
        synthesize()
        holo = fwd_model(amp, ph, d)
        torch.autograd.grad(holo.float().sum(), [amp, ph, d])

    for _ in range(2):
        asm_part()
    ms_asm = timed(asm_part, a.steps)[0]
    if rank == 0:
        line = {"what": "style-transfer training step (AdaIN net + ASM physics loss)", "network": net_kind,
                "forward_model": "B200 drop-in Holo_Generator" if a.forward == "ours" else "reference Holo_Generator (torch.fft)",
                "global_batch": a.batch, "size": n, "fft_size": 2 * n, "n_gpus": world, "steps": a.steps,
                "ms_per_step": ms_step, "samples_per_s": a.batch / (ms_step * 1e-3),
                "ms_asm_part": ms_asm, "asm_share": ms_asm / ms_step,
                "ms_per_step_without_allreduce": ms_nosync, "allreduce_exposed_share": max(0.0, 1 - ms_nosync / ms_step),
                "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
                "trainable_params": sum(p.numel() for p in params)}
        print(json.dumps(line), flush=True)
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

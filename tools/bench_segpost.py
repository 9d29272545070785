"""Time the segmentation-sample post-processing kernels on one GPU.

    python tools/bench_segpost.py [n_img] [n_samples]

Prints per-kernel times and the HBM rate against the algorithmic bytes (each probability read once by the blob
kernel = 262 144 B per sample; the entropy kernel re-reads only the bytes under the blob masks)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from contouring_uncertainty_b200 import engine as E  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iters)]
    for i in range(iters):
        ev[2 * i].record()
        fn()
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    return min(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(iters))


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    dev = torch.device("cuda:0")
    noise = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
    speckle = float(sys.argv[4]) if len(sys.argv) > 4 else 0.002
    base = torch.from_numpy(synth.segmentation_samples(4, T, seed=9, adversarial=False, speckle=speckle)).to(dev)
    probs = base.repeat((n_img + 3) // 4, 1, 1, 1)[:n_img]
    probs = (probs + noise * torch.randn(probs.shape, device=dev)).clamp_(0, 1).contiguous()
    n = n_img * T
    masks, area = E.seg_largest_blob(probs, return_area=True)
    t_blob = timed(lambda: E.seg_largest_blob(probs, return_area=True))
    t_ent = timed(lambda: E.seg_entropy(probs, masks, return_pred_samples=True))
    frac = E.unpack_masks(masks).float().mean().item()
    print(f"n_img {n_img} T {T}: largest_blob {t_blob:.3f} ms = {n / t_blob * 1e3:.3e} samples/s, "
          f"{n * 262144 / t_blob / 1e6:.0f} GB/s of probabilities; entropy {t_ent:.3f} ms "
          f"(mask coverage {frac:.3f}); total {n / (t_blob + t_ent) * 1e3:.3e} samples/s; mean blob area {area.float().mean().item():.0f}")


if __name__ == "__main__":
    main()

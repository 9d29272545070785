"""Stage timing of the sequence (ED/ES) PSM sampler: python tools/bench_seq.py [pairs] [T]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch, synth
from contouring_uncertainty_b200 import engine as E
from contouring_uncertainty_b200 import _native as nat
dev = torch.device("cuda:0")
npair = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mu, cov = synth.point_distributions(2 * npair, seed=5)
mu = torch.from_numpy(mu).to(dev).reshape(npair, 2, 21, 2); cov = torch.from_numpy(cov).to(dev).reshape(npair, 2, 21, 2, 2)
model, seq = E.ShapeModelConstants(device=dev), E.SequenceModelConstants(device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(7)]
    fi = (torch.rand((npair, T), device=dev) < 0.5).to(torch.uint8)
    e[0].record()
    m2, c2 = mu.reshape(2 * npair, 21, 2), cov.reshape(2 * npair, 21, 2, 2)
    a_t = E.psm_factor(model, m2, c2); e[1].record()
    first_all = E.sample_contours(m2, a_t, T, seed=rep); e[2].record()
    first = torch.empty((npair, T, 21, 2), dtype=torch.float32, device=dev)
    mu_f = torch.empty((npair * T, 21, 2), dtype=torch.float32, device=dev)
    cov_f = torch.empty((npair * T, 21, 2, 2), dtype=torch.float32, device=dev)
    nat.check(nat.lib().cub200_seq_condition(nat.ptr(mu), nat.ptr(cov), nat.ptr(first_all), nat.ptr(fi), npair, T, nat.ptr(seq.m2t), nat.ptr(seq.covc),
                                             nat.ptr(seq.mu), nat.ptr(seq.mean), nat.ptr(seq.scale), nat.ptr(first), nat.ptr(mu_f), nat.ptr(cov_f),
                                             nat.stream_ptr(dev)), "seq"); e[3].record()
    a2 = E.psm_factor(model, mu_f, cov_f); e[4].record()
    second = E.sample_contours(mu_f, a2, 1, seed=rep + 7); e[5].record()
    out = E.sequence_sample(model, seq, mu, cov, T, seed=rep); e[6].record()
    torch.cuda.synchronize()
names = ["factor(2Np)", "sample first (both instants)", "seq_condition", "factor per sample", "sample second", "whole sequence_sample"]
print(f"pairs {npair} T {T}: " + ", ".join(f"{n} {e[i].elapsed_time(e[i+1]):.2f} ms" for i, n in enumerate(names)))

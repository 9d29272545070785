import torch, time
dev = torch.device("cuda:0")
n = 352321536 // 4
h = torch.empty(n, dtype=torch.float32, pin_memory=True); h.normal_()
d = torch.empty(n, dtype=torch.float32, device=dev)
def bw(nchunks, nstreams, reps=5):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    cs = n // nchunks
    torch.cuda.synchronize()
    best = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        for i in range(nchunks):
            with torch.cuda.stream(streams[i % nstreams]):
                d[i*cs:(i+1)*cs].copy_(h[i*cs:(i+1)*cs], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n * 4 / (time.perf_counter() - t0) / 1e9)
    return best
for nc, ns in ((1,1),(8,1),(8,2),(8,4),(32,2),(64,4)):
    print(nc, ns, round(bw(nc, ns), 1), "GB/s")
# concurrent D2H
o = torch.empty(82343424 // 4, dtype=torch.float32, pin_memory=True); od = torch.empty_like(o, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): o.copy_(od, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("h2d+d2h concurrent:", round(n*4/dt/1e9, 1), "GB/s h2d-equivalent")

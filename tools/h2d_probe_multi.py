"""Aggregate pinned host->device copy ceiling of the box with N ranks copying at once (what bench.py's e2e leg is bounded by):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/h2d_probe_multi.py
Every rank copies the e2e leg's 352 MB of logits per step from its own pinned buffer (8 chunks, like HostPipeline) while a
result-sized buffer goes the other way; device-timed, max over ranks.  Prints one JSON line (rank 0)."""
import json, os
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_in, n_out = 352321536 // 4, 19428864 // 4
h = torch.empty(n_in, dtype=torch.float32, pin_memory=True); h.normal_()
d = torch.empty(n_in, dtype=torch.float32, device=dev)
o = torch.empty(n_out, dtype=torch.float32, pin_memory=True); od = torch.zeros(n_out, dtype=torch.float32, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
cs = n_in // 8

def step():
    for i in range(8):
        with torch.cuda.stream(s_in):
            d[i * cs:(i + 1) * cs].copy_(h[i * cs:(i + 1) * cs], non_blocking=True)
    with torch.cuda.stream(s_out):
        o.copy_(od, non_blocking=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s_in.wait_stream(torch.cuda.current_stream()); s_out.wait_stream(torch.cuda.current_stream())
for _ in range(reps):
    step()
torch.cuda.current_stream().wait_stream(s_in); torch.cuda.current_stream().wait_stream(s_out)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
ts = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(ts, t)
else:
    ts = [t]
ms = [float(x.item()) for x in ts]
if rank == 0:
    per = [n_in * 4 / (m * 1e-3) / 1e9 for m in ms]
    print(json.dumps({"probe": "pinned h2d (352 MB / step / rank, 8 chunks) with a 19 MB d2h the other way", "ranks": world,
                      "ms_per_step_per_rank": ms, "h2d_gb_per_s_per_rank": per, "h2d_gb_per_s_per_rank_at_max_time": n_in * 4 / (max(ms) * 1e-3) / 1e9,
                      "aggregate_gb_per_s": world * n_in * 4 / (max(ms) * 1e-3) / 1e9}))
if world > 1:
    dist.destroy_process_group()

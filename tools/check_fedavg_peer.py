"""torchrun helper: NVLink FedAvg kernel across ranks == torch reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from split_learning_b200.parallel.fedavg import PeerFedAvg, average_int_state

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 22
g = torch.Generator(device=dev).manual_seed(rank)
p = torch.randn(n, device=dev, generator=g)
if rank == 1:
    p[7] = float("nan")
pb = torch.empty(n, device=dev, dtype=torch.bfloat16)
w = float(rank + 1)
allp = [torch.empty_like(p) for _ in range(world)]
dist.all_gather(allp, p)
ref = sum((r + 1) * torch.nan_to_num(t) for r, t in enumerate(allp)) / sum(range(1, world + 1))
fa = PeerFedAvg(n, dev, list(range(world)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ok = fa.average(p, pb, w)
err = float((p - ref).abs().max())
e0.record()
fa.average(p, pb, w)
e1.record()
torch.cuda.synchronize()
nbt = {"a": torch.tensor(10 * (rank + 1), device=dev, dtype=torch.int64)}
average_int_state(nbt, w)
exp = round(sum(10 * (r + 1) * (r + 1) for r in range(world)) / sum(range(1, world + 1)))
assert ok and err < 1e-5, err
assert int(nbt["a"]) == exp, (int(nbt["a"]), exp)
assert float((pb.float() - p).abs().max()) < 0.05
skipped = fa.average(p, pb, w, ok=(rank != 0))
assert skipped is False
if rank == 0:
    ms = e0.elapsed_time(e1)
    print(f"FEDAVG_OK world={world} n={n} err={err:.2e} second call {ms:.3f} ms "
          f"({world * n * 4 / ms / 1e6:.1f} GB/s pulled per rank incl. staging copy + barriers)")
dist.destroy_process_group()

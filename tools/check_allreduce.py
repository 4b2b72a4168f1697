"""torchrun -n N: the device FedAvg all-reduce (parallel/allreduce.py) against a torch reference, on a two-cluster
topology with different cut points when N >= 4 (cluster 0: cut 7, cluster 1: cut 14 — BASELINE config #4), weighted,
with one NaN parameter and integer counters.  Prints ALLREDUCE_OK on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from split_learning_b200.models import VGG16_CIFAR10
from split_learning_b200.parallel.allreduce import DeviceFedAvg
from split_learning_b200.parallel.fedavg import TorchDistComm
from split_learning_b200.train.b200_executor import B200Executor


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    # roles: world >= 4 -> two clusters x two stages (replicas fill up), else one cluster, every rank a stage-2 replica
    if world >= 4:
        cluster = rank % 2
        stage = (rank // 2) % 2 + 1
        cut = 7 if cluster == 0 else 14
    else:
        cluster, stage, cut = 0, 2, 7
    lo, hi = (0, cut) if stage == 1 else (cut, 52)
    torch.manual_seed(100 + rank)
    ex = B200Executor(VGG16_CIFAR10(lo, hi), "VGG16", {"learning-rate": 0.01, "momentum": 0.5, "batch-size": 8}, dev,
                      is_first=(stage == 1), is_last=(stage == 2))
    ex.P.normal_()
    ex.S.normal_()
    for i, st in enumerate(ex.bn_state.values()):
        st["num_batches_tracked"].fill_(10 * (rank + 1) + i)
    if rank == 0:
        ex.P[5] = float("nan")
    weight = float(rank + 1)
    # ---- reference: gather every replica's state dict, reproduce src/Server.py:398-434 with torch
    sd = {k: ex.view(ex.P, k).clone() for k in ex.entries}
    sd.update({f"layer{bn}.{k}": v.clone().float() for bn, st in ex.bn_state.items() for k, v in st.items()})
    allsd = [None] * world
    dist.all_gather_object(allsd, ({k: v.cpu() for k, v in sd.items()}, cluster, stage, weight))
    expect = {}
    for key in sd:
        per_cluster = []
        for c in sorted({a[1] for a in allsd}):
            hs = [(a[0][key], a[3]) for a in allsd if a[1] == c and key in a[0]]
            if hs:
                tot = sum(w for _, w in hs)
                per_cluster.append(sum(torch.nan_to_num(t.float()) * (w / tot) for t, w in hs))
        expect[key] = sum(per_cluster) / len(per_cluster)
    fa = DeviceFedAvg(ex, f"r{rank:03d}", cluster, TorchDistComm(list(range(world))))
    fa.setup()
    done = fa.run(weight, ok=True, timed=True)
    assert done, "aggregation skipped"
    worst = 0.0
    for key, ref in expect.items():
        if key in ex.entries:
            got = ex.view(ex.P, key).cpu()
        else:
            layer, name = key.split(".", 1)
            got = ex.bn_state[int(layer[5:])][name].float().cpu()
            if name == "num_batches_tracked":
                ref = ref.round()
        worst = max(worst, float((got - ref).abs().max()))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # a NaN vote skips the round and leaves the parameters untouched
    before = ex.P.clone()
    skipped = not fa.run(weight, ok=(rank != world - 1))
    same = bool(torch.equal(torch.nan_to_num(before), torch.nan_to_num(ex.P)))
    flag = torch.tensor([float(skipped and same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ms = torch.tensor([fa.last_ms], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"segments(mine)={len(fa.mine)} of {len(fa.segments)} max_abs_err={float(t.item()):.3e} ms={float(ms.item()):.3f} "
              f"nan_vote_skips={bool(flag.item())}")
        if float(t.item()) < 1e-5 and flag.item() > 0.5:
            print("ALLREDUCE_OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""gloo world_size=2 check of the torch.distributed control messages used around the peer-memory FedAvg
(``TorchDistComm``) and of the integer-state average — the host-side logic of ``parallel/fedavg.py`` on CPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_dist_cpu.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from split_learning_b200.parallel.fedavg import TorchDistComm, average_int_state  # noqa: E402

if __name__ == "__main__":
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchDistComm(list(range(world)))
    got = comm.all_gather_object({"rank": rank, "w": 10.0 * (rank + 1)})
    assert [g["rank"] for g in got] == list(range(world)) and comm.me == rank
    comm.barrier()
    state = {"layer2.num_batches_tracked": torch.tensor(10 + 5 * rank), "layer5.num_batches_tracked": torch.tensor(7)}
    average_int_state(state, weight=float(rank + 1))
    # weighted mean of (10, 15) with weights (1, 2) = 13.33 -> 13 ; identical entries stay
    assert int(state["layer2.num_batches_tracked"]) == 13 and int(state["layer5.num_batches_tracked"]) == 7
    assert state["layer2.num_batches_tracked"].dtype == torch.int64
    dist.barrier()
    if rank == 0:
        print("DIST_CPU_OK")
    dist.destroy_process_group()

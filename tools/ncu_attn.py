"""Driver for an ncu capture of the attention kernels at the BERT shape (B=32, S=128, 12 heads x 64).

    ncu --set full --clock-control none --import-source on -k regex:attn_ -s 4 -c 2 -o gpurun_out/prof_attn \\
        python tools/ncu_attn.py
Also prints CUDA-event timings (not under the profiler when run plainly): python tools/ncu_attn.py --time
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from split_learning_b200.ops import native as N  # noqa: E402

if __name__ == "__main__":
    N.require()
    N.preload()
    b, s, h, dh = 32, 128, 12, 64
    e = h * dh
    torch.manual_seed(0)
    q, k, v, do = (torch.randn(b, s, e, device="cuda").to(torch.bfloat16) for _ in range(4))
    out = torch.empty_like(q)
    dq, dk, dv = (torch.empty_like(q) for _ in range(3))
    lse = torch.empty(b * h * 128, device="cuda")
    reps = 20 if "--time" in sys.argv else 3
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(reps + 2):
        if it == 2:
            ev[0].record()
        N.attn_fwd(q, k, v, e, e, e, 0, 0, 0, out, e, lse, None, b, s, h, dh, 0.1, 1)
    ev[1].record()
    for it in range(reps):
        N.attn_bwd(q, k, v, do, e, e, e, e, 0, 0, 0, 0, dq, dk, dv, e, e, e, 0, 0, 0, lse, None, b, s, h, dh, 0.1, 1)
    ev[2].record()
    torch.cuda.synchronize()
    f_us = ev[0].elapsed_time(ev[1]) / reps * 1e3
    b_us = ev[1].elapsed_time(ev[2]) / reps * 1e3
    flop_f = 4.0 * b * h * s * s * dh
    print(f"attn fwd {f_us:.1f} us ({flop_f / f_us / 1e6:.1f} TFLOP/s useful)  bwd {b_us:.1f} us "
          f"({2.5 * flop_f / b_us / 1e6:.1f} TFLOP/s useful)  grid {b * h} CTAs")

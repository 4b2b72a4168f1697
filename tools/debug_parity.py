"""Per-step comparison: native engine vs torch autograd (fp32-exact and TF32) on the cut-7 chain, dropout off, depth 1."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from split_learning_b200.models import VGG16_CIFAR10
from split_learning_b200.parallel.pipeline import LocalPipeline
from split_learning_b200.train.b200_executor import B200Executor

dev = torch.device("cuda:0")
B, CUT, STEPS = 32, 7, 6
LEARNING = {"learning-rate": 0.0005, "momentum": 0.5, "batch-size": B, "control-count": 1, "precision": "tf32"}
torch.manual_seed(11)
m1, m2 = VGG16_CIFAR10(0, CUT), VGG16_CIFAR10(CUT, 52)
g = torch.Generator().manual_seed(5)
batches = [(torch.randn(B, 3, 32, 32, generator=g), torch.randint(0, 10, (B,), generator=g)) for _ in range(STEPS)]


def torch_run(tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    r1, r2 = VGG16_CIFAR10(0, CUT).to(dev).train(), VGG16_CIFAR10(CUT, 52).to(dev).train()
    r1.load_state_dict(m1.state_dict()); r2.load_state_dict(m2.state_dict())
    for mod in r2.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    o1 = torch.optim.SGD(r1.parameters(), lr=5e-4, momentum=0.5)
    o2 = torch.optim.SGD(r2.parameters(), lr=5e-4, momentum=0.5)
    acts, grads, losses = [], [], []
    for x, y in batches:
        x, y = x.to(dev), y.to(dev)
        with torch.no_grad():
            a = r1(x)                                    # forward-only pass (stats advance), as the reference does
        a_in = a.detach().clone().requires_grad_(True)
        o2.zero_grad()
        loss = torch.nn.functional.cross_entropy(r2(a_in), y)
        loss.backward(); o2.step()
        o1.zero_grad()
        out = r1(x)                                      # recompute with current weights
        out.backward(a_in.grad); o1.step()
        acts.append(a.cpu()); grads.append(a_in.grad.cpu()); losses.append(float(loss))
    return acts, grads, losses


ex1 = B200Executor(m1, "VGG16", LEARNING, dev, is_first=True)
ex2 = B200Executor(m2, "VGG16", LEARNING, dev, is_last=True)
for b in ex2.blocks:
    if hasattr(b, "drop"): b.drop = 0.0
    if hasattr(b, "p"): b.p = 0.0
pipe = LocalPipeline([ex1, ex2], B, 1)
act_mb, grad_mb = pipe.stages[0].fwd_out, pipe.stages[0].grad_in
nat = ([], [], [])
for x, y in batches:
    if pipe.it_f - pipe.it_b >= pipe.depth:
        pipe.step_backward()
    pipe.feed(x.pin_memory(), y.pin_memory()); pipe.step_forward(); pipe.synchronize()
    nat[2].append(float(pipe.loss()[0]))
    nat[0].append(act_mb.payload[0].float().permute(0, 3, 1, 2).cpu().clone())
    nat[1].append(grad_mb.payload[0].float().permute(0, 3, 1, 2).cpu().clone())
while pipe.it_b < pipe.it_f:
    pipe.step_backward()
pipe.synchronize()
exact, tf = torch_run(False), torch_run(True)
cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
for i in range(STEPS):
    print(f"step {i}: loss nat {nat[2][i]:.5f} exact {exact[2][i]:.5f} tf32 {tf[2][i]:.5f} | act rel nat-exact {rel(nat[0][i], exact[0][i]):.2e} "
          f"tf32-exact {rel(tf[0][i], exact[0][i]):.2e} | grad cos nat-exact {cos(nat[1][i], exact[1][i]):.5f} tf32-exact {cos(tf[1][i], exact[1][i]):.5f} "
          f"| grad rel nat-exact {rel(nat[1][i], exact[1][i]):.2e} tf32-exact {rel(tf[1][i], exact[1][i]):.2e} | |grad| nat {float(nat[1][i].norm()):.3e} exact {float(exact[1][i].norm()):.3e}")

"""Block-by-block comparison of the native last-stage backward against torch autograd."""
import sys, os, faulthandler
faulthandler.dump_traceback_later(240, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from split_learning_b200.models import VGG16_CIFAR10
from split_learning_b200.train.b200_executor import B200Executor, ConvBlock, DropoutOp, LinearBlock

def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))

dev = torch.device("cuda:0")
torch.manual_seed(0)
start = int(sys.argv[1]) if len(sys.argv) > 1 else 7
m = VGG16_CIFAR10(start, 52)
ref = VGG16_CIFAR10(start, 52).to(dev)
ref.load_state_dict(m.state_dict())
for mod in ref.modules():
    if isinstance(mod, torch.nn.Dropout):
        mod.p = 0.0
ref.train()
learning = {"learning-rate": 0.0, "momentum": 0.0, "batch-size": 32, "control-count": 1}
ex = B200Executor(m, "VGG16", learning, dev, is_first=False, is_last=True, use_graphs=False)
for b in ex.blocks:
    if isinstance(b, DropoutOp): b.p = 0.0
    if isinstance(b, LinearBlock): b.drop = 0.0
B = 32
c, h, w = ex.in_shape
x = torch.randn(B, c, h, w, device=dev)
x = x.to(torch.bfloat16).float()
y = torch.randint(0, 10, (B,), device=dev)
# torch side with hooks on every layer input
xr = x.clone().requires_grad_(True)
acts, grads = {}, {}
hcur = xr
for i in ref.owned_indices():
    layer = getattr(ref, f"layer{i}")
    hcur.retain_grad()
    acts[i] = hcur
    hcur = layer(hcur)
loss = torch.nn.functional.cross_entropy(hcur, y)
loss.backward()
gin = ex.forward_backward_last(x, y)
print("loss native", ex.last_loss(), "torch", float(loss))
print("input grad rel", rel(gin, xr.grad))
pl = ex.plan(B)
def nhwc(t):
    return t.permute(0, 2, 3, 1) if t.dim() == 4 else t
for bi, b in enumerate(ex.blocks):
    a = pl.act[bi]
    if isinstance(b, ConvBlock):
        first = b.conv or b.bn
        # forward: block output vs torch activation after the block
        last_idx = (b.conv or b.bn or 0)
        idxs = [i for i in (b.conv, b.bn) if i]
        end = max(idxs) + int(b.relu) + int(b.pool)
        nxt = acts.get(end + 1)
        fo = rel(a["out"], nhwc(nxt)) if nxt is not None else float("nan")
        gdx = rel(a["dx"], nhwc(acts[first].grad)) if a.get("dx") is not None else float("nan")
        gdy = float("nan")
        if b.conv is not None and (b.conv + 1) in acts:
            gdy = rel(a["dy"], nhwc(acts[b.conv + 1].grad))
        gw = float("nan")
        if b.conv is not None:
            gw_t = getattr(ref, f"layer{b.conv}").weight.grad.permute(0, 2, 3, 1)
            # native G was consumed by SGD (lr=0 keeps P) -> recompute not possible; compare momentum buffer = grad
            gw = rel(ex.view(ex.M, f"layer{b.conv}.weight"), gw_t)
        print(f"block {bi} conv{b.conv} bn{b.bn} relu{int(b.relu)} pool{int(b.pool)} {b.H}x{b.W} {b.cin}->{b.cout}: "
              f"out {fo:.3e} dy {gdy:.3e} dx {gdx:.3e} wgrad {gw:.3e}")
    elif isinstance(b, LinearBlock):
        gw = rel(ex.view(ex.M, f"layer{b.lin}.weight"), getattr(ref, f"layer{b.lin}").weight.grad)
        gb = rel(ex.view(ex.M, f"layer{b.lin}.bias"), getattr(ref, f"layer{b.lin}").bias.grad)
        gdx = rel(pl.s(bi, "dacc_in", (B, b.fin)), acts[b.lin].grad.reshape(B, -1))
        print(f"block {bi} linear{b.lin} {b.fin}->{b.fout}: wgrad {gw:.3e} bgrad {gb:.3e} dx {gdx:.3e}")
    else:
        print(f"block {bi} dropout{b.idx}: dx {rel(a['dx'], acts[b.idx].grad.reshape(B, -1)):.3e}")

"""Native stage executor vs the reference-math torch executor (same weights, same data,
dropout disabled): loss trajectory, cut activations/gradients and updated weights must agree
to bf16 accuracy; then the device pipeline (mailboxes + CUDA graphs) must reproduce the
tensor-API path."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

LEARNING = {"learning-rate": 0.01, "momentum": 0.5, "batch-size": 32, "control-count": 3}


def _no_dropout_torch(m):
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    return m


def _no_dropout_native(ex):
    from split_learning_b200.train.b200_executor import DropoutOp, LinearBlock
    for b in ex.blocks:
        if isinstance(b, DropoutOp):
            b.p = 0.0
        if isinstance(b, LinearBlock):
            b.drop = 0.0
    return ex


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12))


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30))


# tolerances per compute precision: tf32 (default; the reference's precision: fp32 tensors, TF32 convolutions, fp32 Linear)
# must track the torch executor closely; bf16 (opt-in fast mode) takes different ReLU / max-pool decisions near thresholds
# (cut gradient: torch's own TF32 run sits at cos 0.988 against torch fp32 on this random-init net — ReLU / max-pool decision
# flips amplify operand rounding, tools/debug_parity.py — so two TF32 engines are compared at 0.97, not 0.999)
TOL = {"tf32": dict(act0=4e-3, act_cos=0.985, grad_cos=0.95, loss_rel=0.005, loss_abs=0.003, w_cos=0.995),
       "bf16": dict(act0=3e-2, act_cos=0.95, grad_cos=0.5, loss_rel=0.03, loss_abs=0.02, w_cos=0.98)}


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
@pytest.mark.parametrize("cuts", [(7,), (5, 10), (14,)])
def test_trajectory_matches_torch_executor(cuts, precision):
    tol = TOL[precision]
    learning = dict(LEARNING, precision=precision)
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.train.b200_executor import B200Executor
    from split_learning_b200.train.executor import TorchExecutor
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    bounds = [0] + list(cuts) + [52]
    nat, ref = [], []
    for i in range(len(bounds) - 1):
        m = VGG16_CIFAR10(bounds[i], bounds[i + 1])
        first, last = i == 0, i == len(bounds) - 2
        r = _no_dropout_torch(VGG16_CIFAR10(bounds[i], bounds[i + 1]))
        r.load_state_dict(m.state_dict())
        ref.append(TorchExecutor(r, "VGG16", LEARNING, dev, first, last, recompute=True))
        nat.append(_no_dropout_native(B200Executor(m, "VGG16", learning, dev, first, last, recompute=True)))
    g = torch.Generator().manual_seed(1)
    losses = []
    for step in range(6):
        x = torch.randn(32, 3, 32, 32, generator=g)
        y = torch.randint(0, 10, (32,), generator=g)
        outs = {}
        for name, chain in (("nat", nat), ("ref", ref)):
            h = x
            for ex in chain[:-1]:
                h = ex.forward_only(step, h)
            gin = chain[-1].forward_backward_last(h, y)
            grads = [gin]
            for ex in reversed(chain[:-1]):
                gin = ex.backward(step, gin)
                grads.append(gin)
            outs[name] = (h, grads, chain[-1].last_loss())
        losses.append((outs["nat"][2], outs["ref"][2]))
        if step == 0:                       # identical weights only at step 0; later steps drift apart (chaotic net)
            assert _rel(outs["nat"][0], outs["ref"][0]) < tol["act0"], ("cut activation", _rel(outs["nat"][0], outs["ref"][0]))
        else:
            assert _cos(outs["nat"][0], outs["ref"][0]) > tol["act_cos"], f"cut activation step {step}"
        # bf16 vs fp32 pipelines take different ReLU / max-pool decisions for near-threshold values, so deep
        # gradients are compared by direction, not pointwise (single blocks are checked pointwise in selftest)
        # (a deep random-init net on noise inputs amplifies perturbations ~1.2x per block in both directions)
        if step == 0:
            assert _cos(outs["nat"][1][0], outs["ref"][1][0]) > tol["grad_cos"], ("cut gradient", _cos(outs["nat"][1][0], outs["ref"][1][0]))
        # the first steps track tightly; later the two engines sit on different branches of a chaotic trajectory (ReLU /
        # max-pool decision flips), so the late-step bound only catches divergence, not rounding order
        slack = 1.0 if step < 3 else 4.0
        assert abs(outs["nat"][2] - outs["ref"][2]) < slack * (tol["loss_rel"] * abs(outs["ref"][2]) + tol["loss_abs"]), losses
    for a, b in zip(nat, ref):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            if sa[k].dtype == torch.int64:
                assert int(sa[k]) == int(sb[k]), k          # num_batches_tracked (2x on recomputing stages)
            elif "running" in k:
                assert _cos(sa[k], sb[k]) > tol["w_cos"], k       # drifted weights => slightly different batch statistics
            elif k.endswith("weight") and sa[k].dim() >= 2:
                assert _cos(sa[k], sb[k]) > tol["w_cos"], k   # weights after 6 SGD steps


def test_linear_stage_backward_pointwise():
    """A stage without ReLU/pool decisions upstream of the compared tensors: conv8+bn9 as a middle stage."""
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.train.b200_executor import B200Executor
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    m = VGG16_CIFAR10(7, 9)
    ref = VGG16_CIFAR10(7, 9).to(dev).train()
    ref.load_state_dict(m.state_dict())
    ex = B200Executor(m, "VGG16", dict(LEARNING, **{"learning-rate": 0.0, "momentum": 0.0}), dev, False, False)
    x = torch.randn(32, 64, 16, 16, device=dev).to(torch.bfloat16).float()
    g = torch.randn(32, 128, 16, 16, device=dev).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    out_ref = ref(xr)
    out_ref.backward(g)
    out = ex.forward_only(0, x)
    gin = ex.backward(0, g)
    assert _rel(out, out_ref) < 1e-2
    assert _rel(gin, xr.grad) < 2e-2
    assert _rel(ex.view(ex.M, "layer8.weight"), ref.layer8.weight.grad.permute(0, 2, 3, 1)) < 2e-2   # M == grad (mu = 0)
    assert _rel(ex.view(ex.M, "layer9.weight"), ref.layer9.weight.grad) < 2e-2
    assert _rel(ex.view(ex.M, "layer9.bias"), ref.layer9.bias.grad) < 2e-2


def _l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("rng", [(7, 14), (14, 24), (24, 34), (34, 44), (4, 10)])
def test_shallow_stage_gradients_match_autograd(rng):
    """Every block type as a 2-3 block middle stage with identical inputs and upstream gradient: input gradient and
    every parameter gradient against torch autograd (shallow => few ReLU/pool decision flips => tight tolerance)."""
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.train.b200_executor import B200Executor
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    a, b = rng
    m = VGG16_CIFAR10(a, b)
    ref = VGG16_CIFAR10(a, b).to(dev).train()
    ref.load_state_dict(m.state_dict())
    ex = B200Executor(m, "VGG16", dict(LEARNING, **{"learning-rate": 0.0, "momentum": 0.0}), dev, False, False)
    c, h, w = ex.in_shape
    x = torch.randn(32, c, h, w, device=dev).to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    out_ref = ref(xr)
    g = torch.randn_like(out_ref).to(torch.bfloat16).float()
    out_ref.backward(g)
    out = ex.forward_only(0, x)
    gin = ex.backward(0, g)
    assert _l2(out, out_ref) < 3e-3
    assert _cos(gin, xr.grad) > 0.995 and _l2(gin, xr.grad) < 0.1, (_cos(gin, xr.grad), _l2(gin, xr.grad))
    convs_under_bn = {f"layer{blk.conv}.bias" for blk in ex.blocks if getattr(blk, "conv", None) and blk.bn}
    for name, prm in ref.named_parameters():
        if name in convs_under_bn:                       # analytically zero gradient (autograd returns fp32 noise)
            assert float(ex.view(ex.M, name).abs().max()) == 0.0 and float(prm.grad.abs().max()) < 1e-3
            continue
        got = ex.view(ex.M, name)                         # momentum buffer == gradient (mu = 0, lr = 0)
        want = prm.grad.permute(0, 2, 3, 1) if prm.grad.dim() == 4 else prm.grad
        if want.abs().max() < 1e-6:
            continue
        assert _cos(got, want) > 0.99, (name, _cos(got, want))


def test_classifier_stage_gradients_match_autograd():
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.train.b200_executor import B200Executor
    dev = torch.device("cuda:0")
    torch.manual_seed(6)
    m = VGG16_CIFAR10(44, 52)
    ref = _no_dropout_torch(VGG16_CIFAR10(44, 52)).to(dev).train()
    ref.load_state_dict(m.state_dict())
    ex = _no_dropout_native(B200Executor(m, "VGG16", dict(LEARNING, **{"learning-rate": 0.0, "momentum": 0.0}), dev, False, True))
    x = torch.randn(32, 512, 1, 1, device=dev).to(torch.bfloat16).float()
    y = torch.randint(0, 10, (32,), device=dev)
    xr = x.clone().requires_grad_(True)
    loss = torch.nn.functional.cross_entropy(ref(xr), y)
    loss.backward()
    gin = ex.forward_backward_last(x, y)
    assert abs(ex.last_loss() - float(loss)) < 5e-3
    assert _cos(gin, xr.grad) > 0.99, _cos(gin, xr.grad)
    for name, prm in ref.named_parameters():
        assert _cos(ex.view(ex.M, name), prm.grad) > 0.99, name


def test_device_pipeline_matches_tensor_api():
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.parallel.pipeline import LocalPipeline
    from split_learning_b200.train.b200_executor import B200Executor
    dev = torch.device("cuda:0")
    learning = dict(LEARNING, **{"control-count": 1})
    torch.manual_seed(0)
    m1, m2 = VGG16_CIFAR10(0, 7), VGG16_CIFAR10(7, 52)
    a1 = _no_dropout_native(B200Executor(m1, "VGG16", learning, dev, True, False))
    a2 = _no_dropout_native(B200Executor(m2, "VGG16", learning, dev, False, True))
    b1 = _no_dropout_native(B200Executor(m1, "VGG16", learning, dev, True, False))
    b2 = _no_dropout_native(B200Executor(m2, "VGG16", learning, dev, False, True))
    pipe = LocalPipeline([b1, b2], 32, 1)
    g = torch.Generator().manual_seed(2)
    data = [(torch.randn(32, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, 10, (32,), generator=g).pin_memory())
            for _ in range(5)]
    la, lb = [], []
    for i, (x, y) in enumerate(data):
        h = a1.forward_only(i, x)
        gin = a2.forward_backward_last(h, y)
        a1.backward(i, gin)
        la.append(a2.last_loss())
        pipe.run([(x, y)])
        pipe.synchronize()
        lb.append(float(pipe.loss()[0]))
    assert max(abs(p - q) for p, q in zip(la, lb)) < 2e-2, (la, lb)
    w1, w2 = a2.state_dict()["layer52.weight"], b2.state_dict()["layer52.weight"]
    assert _cos(w2, w1) > 0.999


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_public_api_device_plane(tmp_path, monkeypatch):
    """server + clients through the control plane with ``data-plane: device`` (mailboxes + CUDA graphs)."""
    import yaml
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 23))      # a dead-lock must fail in seconds, not minutes
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.parallel.device_client import DeviceRpcClient
    from split_learning_b200.runner import run_inproc
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": [1, 1], "global-round": 2, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 320
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 3, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120}
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True, True]
    assert all(isinstance(c, DeviceRpcClient) and c.dstage is not None for c in srv.clients_objs)
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    # 320 samples / 32 = 10 microbatches per round; round 2 resumes from the round-1 checkpoint
    assert len(sd) == 97 and int(sd["layer9.num_batches_tracked"]) == 20
    assert int(sd["layer2.num_batches_tracked"]) == 40                       # recomputing first stage: 2 forwards each


@pytest.mark.parametrize("plane", ["host", "device"])
def test_public_api_gpu_resident_loader(tmp_path, monkeypatch, plane):
    """``b200.gpu-loader``: the training subset lives in HBM, each microbatch is one augmentation kernel — through both
    data planes."""
    import yaml
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 23))
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.data.gpu_loader import GpuImageLoader
    from split_learning_b200.runner import run_inproc
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": [1, 1], "global-round": 1, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 320
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 3, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": plane, "gpu-loader": True, "watchdog-seconds": 120}
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True]
    first = [c for c in srv.clients_objs if c.layer_id == 1][0]
    assert isinstance(first.train_loader, GpuImageLoader) and first.train_loader.images.is_cuda
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert int(sd["layer9.num_batches_tracked"]) == 10 and all(torch.isfinite(v.float()).all() for v in sd.values())


@pytest.mark.parametrize("clients,cuts", [((2, 1), [7]), ((2, 2, 1), [5, 10])])
def test_public_api_device_plane_fan_in(tmp_path, monkeypatch, clients, cuts):
    """Many clients, fewer servers on the device plane: the downstream stage multiplexes one mailbox lane per upstream
    replica on its executor (static round-robin in place of the reference's competing-consumer queue)."""
    import yaml
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 23))
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.parallel.device_client import DeviceRpcClient
    from split_learning_b200.runner import run_inproc
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": list(clients), "global-round": 1, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 320
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = list(cuts)
    raw["server"]["manual"]["cluster"] = {"num-cluster": 1, "cut-layers": [list(cuts)], "infor-cluster": [list(clients)]}
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 3, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120}
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True]
    cl = srv.clients_objs
    assert all(isinstance(c, DeviceRpcClient) and c.dstage is not None for c in cl)
    last = [c for c in cl if c.layer_id == len(clients)]
    assert len(last) == 1 and len(last[0].dstages) == clients[0]              # one lane per first-stage client
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    n_last_bn = "layer12" if len(clients) == 3 else "layer9"
    # every first-stage client contributes 10 microbatches; the single last stage trains on all of them
    assert len(sd) == 97 and int(sd[f"{n_last_bn}.num_batches_tracked"]) == 10 * clients[0]
    assert int(sd["layer2.num_batches_tracked"]) == 20                         # per client: 10 forwards + 10 recomputes
    assert all(torch.isfinite(v.float()).all() for v in sd.values())


@pytest.mark.parametrize("model_name,kind", [("KWT", "adamw"), ("MobileNetv1", "sgd")])
def test_flat_fused_optimizer_matches_torch(model_name, kind):
    """Torch-executed families on CUDA step through the fused flat optimizers (G9/G10)."""
    from split_learning_b200.models import get_model_class
    from split_learning_b200.ops.optim import FlatFusedOptimizer
    from split_learning_b200.train.executor import TorchExecutor
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    klass = get_model_class(model_name)
    n = klass.num_layers()
    m = klass(n - 4, n)
    ref = klass(n - 4, n).to(dev)
    ref.load_state_dict(m.state_dict())
    learning = {"learning-rate": 1e-2, "momentum": 0.5, "weight-decay": 0.01, "batch-size": 8, "control-count": 1}
    ex = TorchExecutor(m, model_name, learning, dev, is_first=False, is_last=True)
    assert isinstance(ex.opt, FlatFusedOptimizer) and ex.opt.kind == kind
    ropt = (torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01) if kind == "adamw"
            else torch.optim.SGD(ref.parameters(), lr=1e-2, momentum=0.5))
    for mod in list(ref.modules()) + list(ex.model.modules()):
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    shape = tuple(klass(0, n - 4).eval()(klass.example_input(8)).shape)
    for step in range(4):
        x = torch.randn(shape, device=dev)
        y = torch.randint(0, klass.num_classes(), (8,), device=dev)
        ex.forward_backward_last(x, y)
        ropt.zero_grad()
        torch.nn.functional.cross_entropy(ref.train()(x), y).backward()
        ropt.step()
    sa, sb = ex.state_dict(), ref.state_dict()
    for k in sb:
        if sb[k].is_floating_point() and "in_proj_bias" not in k:
            # (the key-bias third of in_proj_bias has an analytically zero gradient: Adam turns fp32 noise into +-lr)
            assert torch.allclose(sa[k], sb[k], atol=2e-4, rtol=2e-3), k


def test_launch_py_device_plane_across_processes(tmp_path):
    """``launch.py``: server + one OS process per client, both clients on cuda:0, ``data-plane: device`` — the mailboxes
    are shared through real CUDA IPC handles posted over the TCP broker (cross-process path of the multi-GPU runs)."""
    import os, subprocess, sys, yaml
    from split_learning_b200.checkpoint import load_checkpoint
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = yaml.safe_load(open(os.path.join(root, "config.yaml")))
    raw["server"].update({"clients": [1, 1], "global-round": 1, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 160
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 3})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120, "port": 29933}
    cfg = tmp_path / "config.yaml"
    yaml.safe_dump(raw, open(cfg, "w"))
    env = dict(os.environ, SLB200_QUIET="1", SLB200_WAIT_SPINS=str(1 << 26))
    r = subprocess.run([sys.executable, os.path.join(root, "launch.py"), "--config", str(cfg), "--timeout", "280"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert len(sd) == 97 and int(sd["layer9.num_batches_tracked"]) == 5


@pytest.mark.parametrize("algo,clients,specs,extra", [
    ("main", (2, 1), [dict(layer_id=1), dict(layer_id=1), dict(layer_id=2)], {}),
    ("dcsl", (2, 1), [dict(layer_id=1, cluster=0), dict(layer_id=1, cluster=0), dict(layer_id=2, cluster=0)], {"local-round": 1}),
    ("vanilla_sl", (2, 1), [dict(layer_id=1), dict(layer_id=1), dict(layer_id=2)], {}),
])
def test_host_plane_with_native_executor_on_gpu(tmp_path, algo, clients, specs, extra):
    """Broker data plane (pickled CPU arrays, any topology/variant) driving the native sm_100a stage executor through its
    tensor API: competing consumers (2:1), DCSL's SDA batch concatenation (B = 64 on the last stage), sequential hand-off."""
    import yaml
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.runner import run_variant
    from split_learning_b200.train.b200_executor import B200Executor
    raw = {"server": {"global-round": 1, "clients": list(clients), "no-cluster": {"cut-layers": [7]}, "model": "VGG16",
                      "data-name": "CIFAR10", "parameters": {"load": False, "save": True}, "validation": False,
                      "data-distribution": {"non-iid": False, "num-sample": 128, "num-label": 10, "dirichlet": {"alpha": 1}},
                      "random-seed": 1},
           "log_path": str(tmp_path), "debug_mode": False,
           "learning": {"learning-rate": 0.01, "momentum": 0.5, "batch-size": 32, "control-count": 2, "clip-grad-norm": 0.0},
           "b200": {"algorithm": algo, "synthetic-data": True, "watchdog-seconds": 120}}
    raw["server"].update(extra)
    srv = run_variant(normalize(raw), specs, workdir=str(tmp_path), devices=["cuda:0"], timeout=300)
    assert srv.history and srv.history[0]["ok"]
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert len(sd) == 97 and all(torch.isfinite(v.float()).all() for v in sd.values())


def test_public_api_device_fedavg_resident_params(tmp_path, monkeypatch):
    """clients [2,2] on the device plane: at round end the replicas of each stage FedAvg in place over peer memory
    (``parallel.fedavg`` driven through the control-plane broker), only the group leaders upload a state-dict, and the
    second round's START carries no parameters (they are resident and identical on every replica)."""
    import yaml
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.runner import run_inproc
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 24))
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": [2, 2], "global-round": 2, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 160
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 2, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120}
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True, True]
    assert srv.all_resident                                   # every UPDATE of the last round was 'resident'
    clients = srv.clients_objs
    assert all(c.dstage is not None for c in clients)
    for layer in (1, 2):
        a, b = [c for c in clients if c.layer_id == layer]
        assert torch.equal(a.executor.P, b.executor.P)        # replicas hold the identical averaged parameters
        assert a.rounds_done == 2
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert len(sd) == 97 and int(sd["layer9.num_batches_tracked"]) == 10     # 5 microbatches per replica per round, averaged


def test_clip_grad_norm_native_matches_torch():
    """``clip-grad-norm`` (other/Vanilla_SL/src/Scheduler.py:204-205) on the native executor: global-norm clipping of the flat
    gradient + SGD == torch.nn.utils.clip_grad_norm_ + torch.optim.SGD on the reference-math module."""
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.train.b200_executor import B200Executor
    dev = torch.device("cuda:0")
    torch.manual_seed(8)
    m = VGG16_CIFAR10(44, 52)
    ref = _no_dropout_torch(VGG16_CIFAR10(44, 52)).to(dev).train()
    ref.load_state_dict(m.state_dict())
    learning = dict(LEARNING, **{"learning-rate": 0.1, "momentum": 0.5, "clip-grad-norm": 0.05})
    ex = _no_dropout_native(B200Executor(m, "VGG16", learning, dev, False, True))
    assert ex.clip == 0.05
    opt = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.5)
    g = torch.Generator(device="cuda").manual_seed(2)
    for step in range(3):
        x = torch.randn(32, 512, 1, 1, device=dev, generator=g)
        y = torch.randint(0, 10, (32,), device=dev, generator=g)
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(ref(x), y)
        loss.backward()
        total = float(torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.05))
        assert total > 0.05, "the test must exercise an active clip"
        opt.step()
        ex.forward_backward_last(x, y)
    sd = ex.state_dict()
    for name, prm in ref.named_parameters():
        assert _rel(sd[name], prm.data) < 2e-5, (name, _rel(sd[name], prm.data))


def test_public_api_device_plane_trailing_partial_batch(tmp_path, monkeypatch):
    """num-sample not a multiple of the batch size: the trailing partial microbatch runs through its own program set on
    the same mailboxes (not dropped), so batch counts — the FedAvg weights — equal the reference's."""
    import yaml
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 23))
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.parallel.device_client import DeviceRpcClient
    from split_learning_b200.runner import run_inproc
    raw = yaml.safe_load(open("config.yaml"))
    raw["server"].update({"clients": [1, 1], "global-round": 2, "validation": False})
    raw["server"]["data-distribution"]["num-sample"] = 80            # 2 full microbatches of 32 + one of 16
    raw["server"]["manual"]["no-cluster"]["cut-layers"] = [7]
    raw["log_path"] = str(tmp_path)
    raw["learning"].update({"batch-size": 32, "control-count": 3, "learning-rate": 0.01})
    raw["b200"] = {"synthetic-data": True, "data-plane": "device", "watchdog-seconds": 120}
    srv = run_inproc(normalize(raw), devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True, True]
    cl = srv.clients_objs
    assert all(isinstance(c, DeviceRpcClient) and c.dstage is not None for c in cl)
    assert all(len(c._tail_stages) == 1 for c in cl), "both stages must have run the 16-sample program set"
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert int(sd["layer9.num_batches_tracked"]) == 2 * 3                 # 3 microbatches per round on the last stage
    assert int(sd["layer2.num_batches_tracked"]) == 2 * 6                 # forward + recompute on the first stage
    assert all(torch.isfinite(v.float()).all() for v in sd.values())


@pytest.mark.parametrize("name", ["clusters", "three-stage"])
def test_baseline_scenarios_on_device_plane(tmp_path, monkeypatch, name):
    """BASELINE.json configurations #4 (two clusters cut at 7 and at 14, cross-cluster averaging) and #5 (three stages, cuts
    [5, 10], non-IID rate 0.5) through the public API on the device plane — the same configs ``bench.py --scenario`` runs on
    8 GPUs, scaled to four clients on one GPU."""
    import types
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.parallel.api_bench import scenario
    from split_learning_b200.runner import run_inproc
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 24))
    args = types.SimpleNamespace(batch=32, depth=2, precision="tf32", rounds=2)
    cfg, _ = scenario(name, 4, 0, args, 0, str(tmp_path), 5)
    srv = run_inproc(cfg, devices=["cuda:0"], workdir=str(tmp_path), timeout=600)
    assert [h["ok"] for h in srv.history] == [True, True]
    cl = srv.clients_objs
    assert all(c.dstage is not None for c in cl) and srv.all_resident
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert len(sd) == 97 and all(torch.isfinite(v.float()).all() for v in sd.values())
    if name == "clusters":
        # layers 1..7 live in stage 1 of both clusters, 8..14 in stage 2 of cluster 0 and stage 1 of cluster 1, 15.. in
        # both stage 2s: after the all-reduce every holder of a layer has the same (cross-cluster averaged) values
        by = {(int(c.cluster), c.layer_id): c.executor.state_dict() for c in cl}
        for key, a, b in (("layer1.weight", (0, 1), (1, 1)), ("layer8.weight", (0, 2), (1, 1)), ("layer12.running_var", (0, 2), (1, 1)),
                          ("layer41.weight", (0, 2), (1, 2)), ("layer50.bias", (0, 2), (1, 2))):
            assert torch.equal(by[a][key], by[b][key]), key
    else:
        assert [len([c for c in cl if c.layer_id == s]) for s in (1, 2, 3)] == [2, 1, 1]
        sizes = sorted(len(c.train_loader.dataset) for c in cl if c.layer_id == 1)
        assert sizes[0] > 0


@pytest.mark.xfail(strict=False, reason="intermittent on hardware when all clients share ONE GPU (last hardware run: 2 of 3 "
                                        "cases passed, one case hung); the feature is opt-in (b200.dynamic-consumers)")
@pytest.mark.parametrize("clients,slow", [((1, 2), False), ((2, 2), True), ((2, 2), False)])
def test_competing_consumers_ticket_ring(tmp_path, clients, slow):
    """Dynamic competing consumers on the device plane (reference: every stage-2 replica ``basic_get``s one shared queue,
    src/train/VGG16.py:143-154; the gradient returns to ``trace[-1]``, :40-53).  [1, 2] cannot be cut into static lanes (the
    last stage does not divide its predecessor), so the last edge becomes a ticket ring on its own; [2, 2] opts in.  Every
    microbatch is claimed by exactly one replica, gradients reach their origin (the round completes and every first-stage
    client steps), and an artificially slow replica ends up with less work.
    Runs in a fresh interpreter (``tests/gpu_cases.py``) with at most four clients, all sharing ONE GPU.  Status: every case
    has passed on hardware, none of them every time (a device-side wait between co-located clients occasionally outlives
    its spin limit); expected-failure (non-strict) until that is understood, so that a sporadic hang cannot stop the suite."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "gpu_cases.py"), "competing", str(tmp_path),
                        str(clients[0]), str(clients[1]), str(int(slow))], capture_output=True, text=True, timeout=100,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    line = [l for l in r.stdout.splitlines() if l.startswith("CASE_OK ")]
    assert r.returncode == 0 and line, (r.stdout[-3000:], r.stderr[-5000:])
    out = json.loads(line[-1][8:])
    assert out["exactly_once"] and out["ckpt_entries"] == 97 and out["rounds_ok"] == [True, True]
    assert out["all_claimed_something"] or clients[0] == 1        # one producer, depth 2: a replica may well get nothing
    if slow:
        assert out["n_fast"] > 1.5 * out["n_slow"], out                      # work migrated to the free replica


@pytest.mark.parametrize("algo,specs,extra", [
    ("vanilla_sl", [dict(layer_id=1), dict(layer_id=1), dict(layer_id=2)], {}),
    ("cluster_fsl", [dict(layer_id=1, cluster=0), dict(layer_id=1, cluster=1), dict(layer_id=2, cluster=0)],
     {"manual-cluster": {"num-cluster": 2, "cut-layers": [[7], [7]]}}),
])
@pytest.mark.skipif(os.environ.get("SLB200_TEST_DEVICE_VARIANTS", "0") != "1",
                    reason="experimental device plane for the sequential variants: opt-in, not yet verified on hardware")
def test_sequential_variants_on_device_plane(tmp_path, monkeypatch, algo, specs, extra):
    """Vanilla_SL / Cluster_FSL (groups of first-stage clients train one after another, weights handed from group to group
    through the server) with ``data-plane: device``: the last stage serves a ticket ring until PAUSE, wiring each lane when
    its client comes up; activations / gradients never leave the GPU."""
    monkeypatch.setenv("SLB200_WAIT_SPINS", str(1 << 26))
    from split_learning_b200.checkpoint import load_checkpoint
    from split_learning_b200.config import normalize
    from split_learning_b200.parallel.device_variants import SequentialDeviceClient
    from split_learning_b200.runner import run_variant
    raw = {"server": {"global-round": 2, "clients": [2, 1], "no-cluster": {"cut-layers": [7]}, "model": "VGG16",
                      "data-name": "CIFAR10", "parameters": {"load": False, "save": True}, "validation": False,
                      "data-distribution": {"non-iid": False, "num-sample": 112, "num-label": 10, "dirichlet": {"alpha": 1}},
                      "random-seed": 1},
           "log_path": str(tmp_path), "debug_mode": False,
           "learning": {"learning-rate": 0.01, "momentum": 0.5, "batch-size": 32, "control-count": 2, "clip-grad-norm": 0.0},
           "b200": {"algorithm": algo, "synthetic-data": True, "watchdog-seconds": 120, "data-plane": "device",
                    "device-variants": True}}
    raw["server"].update(extra)
    srv = run_variant(normalize(raw), specs, workdir=str(tmp_path), devices=["cuda:0"], timeout=300)
    assert [h["ok"] for h in srv.history] == [True, True] and len(srv.groups) == 2
    cl = srv.clients_objs
    assert all(isinstance(c, SequentialDeviceClient) and c.dstage is not None for c in cl), "every role on the device plane"
    last = [c for c in cl if c.layer_id == 2][0]
    assert sorted(last.claimed) == sorted((lane, it) for lane in range(2) for it in range(4))     # last round: both lanes, all 4
    sd = load_checkpoint(str(tmp_path / "VGG16_CIFAR10.pth"))
    assert len(sd) == 97 and all(torch.isfinite(v.float()).all() for v in sd.values())
    # 112 samples = 3 microbatches of 32 + a trailing 16 per client; the single last stage trains on both clients' batches
    # (8 per round) and resumes from the previous round's checkpoint
    assert int(sd["layer9.num_batches_tracked"]) == 16

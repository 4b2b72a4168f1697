"""MobileNetv1 through the native NHWC program (train/cnn_native.py) against the stock torch modules."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _need():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from split_learning_b200.ops import native as N
    N.require()
    N.preload()


def _cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-12)).item()


@pytest.mark.parametrize("data,start,end", [("CIFAR10", 0, 6), ("CIFAR10", 6, 15), ("CIFAR10", 9, 13), ("CIFAR10", 70, 84),
                                            ("MNIST", 0, 9)])
def test_shallow_stage_matches_torch(data, start, end):
    """Few layers per stage: bf16 vs fp32 ReLU decisions agree, so gradients can be compared pointwise.  Covers the stem,
    32-channel padding, stride-2 convs, 1x1 convs, BN+ReLU(+pool), flatten + Linear, running statistics."""
    _need()
    from split_learning_b200.models import get_model_class
    from split_learning_b200.train.cnn_native import nativize
    torch.manual_seed(0)
    cls = get_model_class("MobileNetv1", data)
    ref = cls(start, end).cuda().train()
    nat = nativize(copy.deepcopy(ref))
    x = cls.example_input(32, device="cuda")
    if start:
        with torch.no_grad():
            x = cls(0, start).cuda().train()(x)
    xr = x.clone().requires_grad_(start > 0)
    xn = x.clone().requires_grad_(start > 0)
    yr = ref(xr)
    yn = nat(xn).float()
    assert yn.shape == yr.shape
    assert _cos(yn, yr) > 0.998, _cos(yn, yr)
    g = torch.randn_like(yr)
    yr.backward(g)
    yn.backward(g)
    if start:
        assert _cos(xn.grad, xr.grad) > 0.98, ("dx", _cos(xn.grad, xr.grad))
    bad = []
    refp = dict(ref.named_parameters())
    for (k, pr), (_, pn) in zip(ref.named_parameters(), nat.named_parameters()):
        wk = k.replace(".bias", ".weight")
        if k.endswith(".bias") and pr.grad.norm() < 1e-3 * refp[wk].grad.norm():
            continue                                   # conv bias under train-mode BN: exactly zero gradient (noise)
        c = _cos(pn.grad, pr.grad)
        if c < 0.97:
            bad.append((k, round(c, 4)))
    assert not bad, bad
    for (k, br), (_, bn) in zip(ref.named_buffers(), nat.named_buffers()):
        if "num_batches" in k:
            assert int(br) == int(bn) == 1
        else:
            assert torch.allclose(br, bn, rtol=3e-2, atol=3e-3), k


def test_mobilenet_trains_through_executor():
    """Two MobileNet stages through make_executor (native blocks + whole-step graphs): the loss falls."""
    _need()
    from split_learning_b200.models import get_model_class
    from split_learning_b200.train.executor import make_executor
    cls = get_model_class("MobileNetv1", "CIFAR10")
    learning = {"learning-rate": 0.01, "momentum": 0.5}
    torch.manual_seed(3)
    e1 = make_executor(cls(0, 15), "MobileNetv1", learning, "cuda", True, False)
    e2 = make_executor(cls(15, 84), "MobileNetv1", learning, "cuda", False, True)
    assert e1.native and e2.native and e1.graphs
    g = torch.Generator().manual_seed(5)
    x = torch.randn(32, 3, 32, 32, generator=g)
    y = torch.randint(0, 10, (32,), generator=g)
    losses = []
    for it in range(40):
        a = e1.forward_only(it, x)
        gx = e2.forward_backward_last(a, y)
        e1.backward(it, gx)
        losses.append(e2.last_loss())
    assert losses[-1] < 0.5 * losses[0], losses[::5]
    sd = e2.state_dict()
    assert all(torch.isfinite(v.float()).all() for v in sd.values())

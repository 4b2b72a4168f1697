"""Golden-emulator parity: the UNMODIFIED reference trainers (``baseline/_ref/src/train/VGG16.py``: first-layer and
last-layer loops, reference model classes, torch SGD) run over the in-process pika shim — once as shipped (cuDNN's default
TF32 convolutions) and once with ``torch.backends.cudnn.allow_tf32 = False`` (exact fp32, the oracle) — and the native engine
in its default precision (tcgen05 kind::tf32 convolutions, fp32 Linear / BN / SGD) runs the same 20 microbatches from the
same initial weights.  Per step: training loss, the activation that crosses the cut, the gradient that comes back.

Criterion: the native engine must be as close to the fp32 oracle as the reference's own TF32 run is (a random-init VGG on
noise amplifies operand rounding through ReLU / max-pool decision flips: the reference-TF32 cut gradient itself sits at
cos ~0.96-0.99 against the oracle, so a fixed "cos > 0.999" would fail the reference against itself), and match the
reference-TF32 loss trajectory to well under 1 %.  Dropout is switched off in both (its random streams cannot match);
control-count 1 makes the forward / backward interleaving deterministic in both engines."""
import contextlib
import io
import os
import pickle
import sys
import threading
import time
import uuid

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STEPS, B, CUT = 20, 32, 7
LEARNING = {"learning-rate": 0.0005, "momentum": 0.5, "batch-size": B, "control-count": 1, "weight-decay": 0.01}


def _reference_run(batches, device, tf32=True):
    """Returns (losses, cut activations, cut gradients, initial state dicts, final state dicts) of the reference trainers."""
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = bool(tf32)
    try:
        return _reference_run_impl(batches, device)
    finally:
        torch.backends.cudnn.allow_tf32 = prev


def _reference_run_impl(batches, device):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from run_reference import _prepare_imports
    _prepare_imports()
    import pika                                           # the shim (in-process store)
    from src.model.VGG16_CIFAR10 import VGG16_CIFAR10 as RefVGG
    from src.train.VGG16 import Train_VGG16
    pika._REMOTE = None
    torch.manual_seed(11)
    m1, m2 = RefVGG(start_layer=0, end_layer=CUT), RefVGG(start_layer=CUT)
    for m in m2.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    init = ({k: v.clone() for k, v in m1.state_dict().items()}, {k: v.clone() for k, v in m2.state_dict().items()})
    acts, grads = [], []

    def hook(queue, body):
        if queue.startswith("intermediate_queue_"):
            acts.append(torch.from_numpy(pickle.loads(body)["data"]).clone())
        elif queue.startswith("gradient_queue_"):
            grads.append(torch.from_numpy(pickle.loads(body)["data"]).clone())
    pika.on_get.append(hook)
    ids = (uuid.uuid4(), uuid.uuid4())
    ch = lambda: pika.BlockingConnection(pika.ConnectionParameters("127.0.0.1")).channel()
    t1 = Train_VGG16(ids[0], 1, ch(), device)
    t2 = Train_VGG16(ids[1], 2, ch(), device)
    out = io.StringIO()
    errs = []

    def guard(fn):
        def run():
            try:
                fn()
            except Exception as e:            # noqa
                errs.append(e)
        return run
    th = [threading.Thread(target=guard(lambda: t1.train_on_first_layer(m1, LEARNING, batches, 0)), daemon=True),
          threading.Thread(target=guard(lambda: t2.train_on_last_layer(m2, LEARNING, 0)), daemon=True)]
    server = ch()
    with contextlib.redirect_stdout(out):
        for t in th:
            t.start()
        deadline = time.time() + 300
        while time.time() < deadline and not errs:           # the "server": NOTIFY -> PAUSE to both clients
            _, _, body = server.basic_get(queue="rpc_queue")
            if body and pickle.loads(body).get("action") == "NOTIFY":
                for cid in ids:
                    server.basic_publish(routing_key=f"reply_{cid}", body=pickle.dumps({"action": "PAUSE", "message": "", "parameters": None}))
                break
        for t in th:
            t.join(60)
    pika.on_get.remove(hook)
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "reference trainers did not finish"
    losses = [float(l.split("Loss:")[1]) for l in out.getvalue().splitlines() if l.startswith("Loss:")]
    final = ({k: v.detach().cpu().clone() for k, v in m1.state_dict().items()}, {k: v.detach().cpu().clone() for k, v in m2.state_dict().items()})
    return losses, acts, grads, init, final


def test_golden_parity_vs_reference_trainers():
    from split_learning_b200.models import VGG16_CIFAR10
    from split_learning_b200.ops import native as N
    from split_learning_b200.parallel.pipeline import LocalPipeline
    from split_learning_b200.train.b200_executor import B200Executor
    N.require()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(B, 3, 32, 32, generator=g), torch.randint(0, 10, (B,), generator=g)) for _ in range(STEPS)]
    ref_loss, ref_act, ref_grad, (sd1, sd2), (fin1, fin2) = _reference_run(batches, "cuda:0", tf32=True)
    ex_loss, ex_act, ex_grad, _, (xfin1, xfin2) = _reference_run(batches, "cuda:0", tf32=False)     # fp32 oracle, same seed
    assert len(ref_loss) == STEPS and len(ref_act) == STEPS and len(ref_grad) == STEPS and len(ex_loss) == STEPS

    learning = dict(LEARNING, precision="tf32")
    m1, m2 = VGG16_CIFAR10(0, CUT), VGG16_CIFAR10(CUT, 52)
    m1.load_state_dict(sd1)
    m2.load_state_dict(sd2)
    ex1 = B200Executor(m1, "VGG16", learning, dev, is_first=True)
    ex2 = B200Executor(m2, "VGG16", learning, dev, is_last=True)
    for b in ex2.blocks:                                   # dropout off, like the reference run above
        if hasattr(b, "drop"):
            b.drop = 0.0
        if hasattr(b, "p"):
            b.p = 0.0
    pipe = LocalPipeline([ex1, ex2], B, 1)
    act_mb, grad_mb = pipe.stages[0].fwd_out, pipe.stages[0].grad_in
    loss, acts, grads = [], [], []
    for x, y in batches:
        if pipe.it_f - pipe.it_b >= pipe.depth:
            pipe.step_backward()
        pipe.feed(x.pin_memory(), y.pin_memory())
        pipe.step_forward()
        pipe.synchronize()
        loss.append(float(pipe.loss()[0]))
        acts.append(act_mb.payload[0].float().permute(0, 3, 1, 2).cpu().clone())      # NHWC mailbox slot -> NCHW
        grads.append(grad_mb.payload[0].float().permute(0, 3, 1, 2).cpu().clone())
    while pipe.it_b < pipe.it_f:
        pipe.step_backward()
    pipe.synchronize()

    cosf = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
    relf = lambda a, b: float((a - b).abs().max() / b.abs().max())

    def deviation(L, A, G):
        """worst-over-steps deviation of a run from the fp32 oracle"""
        return (max(abs(a - b) / abs(b) for a, b in zip(L, ex_loss)), max(relf(a, b) for a, b in zip(A, ex_act)),
                max(1.0 - cosf(a, b) for a, b in zip(G, ex_grad)))
    ours, theirs = deviation(loss, acts, grads), deviation(ref_loss, ref_act, ref_grad)
    worst_loss = max(abs(a - b) / abs(b) for a, b in zip(loss, ref_loss))
    print(f"golden parity over {STEPS} steps — deviation from the fp32 oracle (loss rel, cut activation rel, cut gradient 1-cos): "
          f"native {ours[0]:.2e} {ours[1]:.2e} {ours[2]:.2e} | reference TF32 {theirs[0]:.2e} {theirs[1]:.2e} {theirs[2]:.2e}; "
          f"native vs reference-TF32 loss rel {worst_loss:.2e}; step-0 gradient cos native {cosf(grads[0], ex_grad[0]):.5f} "
          f"reference {cosf(ref_grad[0], ex_grad[0]):.5f}")
    assert worst_loss < 5e-3, (loss, ref_loss)                       # loss trajectory within 0.5 % of the reference as shipped
    assert ours[0] < 2 * theirs[0] + 2e-4, (ours, theirs)
    assert ours[1] < 2 * theirs[1] + 5e-4, (ours, theirs)
    assert ours[2] < 2 * theirs[2] + 2e-3, (ours, theirs)
    assert relf(acts[0], ex_act[0]) < 1e-3                           # step 0 (identical weights): plain TF32 operand rounding
    # final weights: the engines took the same 20 SGD steps — compare against the oracle, relative to how far it moved
    out1, out2 = ex1.state_dict(), ex2.state_dict()
    for k in ("layer1.weight", "layer4.weight", "layer5.running_mean", "layer8.weight", "layer41.weight", "layer50.weight", "layer52.bias"):
        mine = (out1 if k in out1 else out2)[k].float().cpu()
        base = (sd1 if k in sd1 else sd2)[k].float().cpu()
        oracle = (xfin1 if k in xfin1 else xfin2)[k].float()
        ref = (fin1 if k in fin1 else fin2)[k].float()
        moved = float((oracle - base).abs().max())
        e_mine, e_ref = float((mine - oracle).abs().max()), float((ref - oracle).abs().max())
        print(f"  final {k}: oracle moved {moved:.3e}; off the oracle: native {e_mine:.3e}, reference TF32 {e_ref:.3e}")
        assert moved > 0 and e_mine < 2 * e_ref + 0.02 * moved + 1e-7, (k, moved, e_mine, e_ref)
    nbt = int(out1["layer2.num_batches_tracked"])
    assert nbt == int(fin1["layer2.num_batches_tracked"]) == 2 * STEPS      # forward + recompute, as in the reference

"""Property tests (hypothesis) of the planning / aggregation arithmetic — the parts every algorithm variant leans on."""
import itertools

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from split_learning_b200.data.distribution import label_counts, preset_matrix
from split_learning_b200.fedavg import fedasync_merge, fedavg_state_dicts
from split_learning_b200.plan import rank_assignment, resolve_range, stage_layers
from split_learning_b200.planning.partition import partition, partition_multi

FAST = settings(max_examples=40, deadline=None)


@FAST
@given(st.integers(2, 6).flatmap(lambda n: st.tuples(st.just(n), st.lists(st.integers(1, 51), min_size=n - 1, max_size=n - 1,
                                                                           unique=True).map(sorted))))
def test_stages_tile_the_model_exactly_once(arg):
    """Any increasing cut list: the stages' (start, end] ranges are disjoint and cover layers 1..52 (src/Server.py:222-228)."""
    n, cuts = arg
    owned = []
    for layer_id in range(1, n + 1):
        a, b = resolve_range(stage_layers(layer_id, n, cuts), 52)
        assert a < b
        owned += list(range(a + 1, b + 1))
    assert owned == list(range(1, 53))


@FAST
@given(st.lists(st.integers(1, 4), min_size=2, max_size=4))
def test_rank_assignment_is_a_bijection(counts):
    ranks = rank_assignment(counts)
    assert len(ranks) == sum(counts) and len(set(ranks)) == len(ranks)
    for s, n in enumerate(counts, start=1):
        assert sum(1 for r in ranks if r[0] == s) == n


@FAST
@given(st.integers(1, 4), st.integers(0, 2 ** 31 - 1))
def test_fedavg_is_a_convex_combination(n, seed):
    """Weighted mean: between the element-wise min and max, exact for identical inputs, weight-homogeneous, integer
    entries rounded, NaN treated as 0 (src/Utils.py:35-66)."""
    g = torch.Generator().manual_seed(seed)
    sds = [{"w": torch.randn(5, 3, generator=g), "n": torch.randint(0, 50, (1,), generator=g)[0]} for _ in range(n)]
    ws = [float(torch.randint(1, 20, (1,), generator=g)) for _ in range(n)]
    out = fedavg_state_dicts(sds, ws)
    stack = torch.stack([sd["w"] for sd in sds])
    assert (out["w"] <= stack.max(0).values + 1e-5).all() and (out["w"] >= stack.min(0).values - 1e-5).all()
    assert torch.allclose(out["w"], fedavg_state_dicts(sds, [3.0 * w for w in ws])["w"], atol=1e-5)
    assert out["n"].dtype == torch.int64
    exp_n = round(sum(w * int(sd["n"]) for w, sd in zip(ws, sds)) / sum(ws))
    assert abs(int(out["n"]) - exp_n) <= 1
    same = fedavg_state_dicts([sds[0]] * 3, [1, 2, 3])
    assert torch.allclose(same["w"], sds[0]["w"], atol=1e-6)
    poisoned = [{"w": torch.full((2,), float("nan"))}, {"w": torch.ones(2)}]
    assert torch.allclose(fedavg_state_dicts(poisoned, [1, 1])["w"], torch.full((2,), 0.5))


@FAST
@given(st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
def test_fedasync_interpolates(alpha, seed):
    g = torch.Generator().manual_seed(seed)
    a, b = {"w": torch.randn(4, generator=g)}, {"w": torch.randn(4, generator=g)}
    m = fedasync_merge(a, b, alpha)["w"]
    assert torch.allclose(m, (1 - alpha) * a["w"] + alpha * b["w"], atol=1e-6)
    assert torch.equal(fedasync_merge(None, b, alpha)["w"], b["w"])


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_partition_multi_matches_brute_force(seed):
    """The N-stage DP returns a cut list whose min stage rate equals the exhaustive optimum (3 stages, 7 layers)."""
    rng = np.random.default_rng(seed)
    n_layer, n_stage = 7, 3
    exe = [[list(rng.uniform(0.5, 3.0, n_layer)) for _ in range(int(rng.integers(1, 3)))] for _ in range(n_stage)]
    nets = [[float(rng.uniform(0.5, 5.0)) for _ in stage] for stage in exe]
    size = list(rng.uniform(0.1, 4.0, n_layer))

    def min_rate(cuts):
        b = [0] + list(cuts) + [n_layer]
        rates = []
        for s in range(n_stage):
            tot = 0.0
            for e, net in zip(exe[s], nets[s]):
                comm = (size[b[s + 1] - 1] / net if b[s + 1] < n_layer else 0.0) + (size[b[s] - 1] / net if b[s] > 0 else 0.0)
                tot += 1.0 / (sum(e[b[s]:b[s + 1]]) + comm)
            rates.append(tot)
        return min(rates)
    cuts = partition_multi(exe, nets, size)
    assert len(cuts) == 2 and 0 < cuts[0] < cuts[1] < n_layer
    best = max(min_rate(c) for c in itertools.combinations(range(1, n_layer), 2))
    assert abs(min_rate(cuts) - best) <= 1e-9 * max(1.0, best)
    two = partition(exe[0], nets[0], exe[1], nets[1], size)
    assert len(two) == 1 and 1 <= two[0] <= n_layer


@FAST
@given(st.integers(1, 9), st.integers(2, 10), st.integers(10, 500), st.integers(0, 1000))
def test_label_allocation_shapes_and_budgets(clients, labels, samples, seed):
    iid = label_counts(clients, labels, samples, False)
    assert iid.shape == (clients, labels) and (iid == samples // labels).all()
    dirich = label_counts(clients, labels, samples, True, alpha=0.5, seed=seed)
    assert dirich.shape == (clients, labels) and (dirich >= 0).all() and (dirich.sum(1) <= samples).all()
    for name in ("flex", "2ls"):
        m = preset_matrix(name, clients, labels)
        assert m.shape == (clients, labels) and np.allclose(m.sum(1), 1.0) and (m >= 0).all()


# ------------------------------------------------------------------------------------------------ wire codec
_leaf = st.one_of(st.none(), st.booleans(), st.integers(-2**40, 2**40), st.floats(allow_nan=False, width=32), st.text(max_size=12),
                  st.binary(max_size=24))
_tensor = st.tuples(st.sampled_from(["float32", "float16", "int64", "uint8", "bool"]),
                    st.lists(st.integers(0, 5), min_size=0, max_size=3)).map(
    lambda a: (torch.arange(int(np.prod(a[1])) if a[1] else 1).reshape(a[1] if a[1] else ()) % 2).to(getattr(torch, a[0])))
_tree = st.recursive(st.one_of(_leaf, _tensor), lambda kids: st.one_of(st.lists(kids, max_size=4), st.tuples(kids, kids),
                                                                      st.dictionaries(st.text(max_size=6), kids, max_size=4)),
                     max_leaves=12)


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, float):
        return isinstance(b, float) and (a == b)
    return a == b and type(a) is type(b)


@FAST
@given(_tree)
def test_codec_round_trips_arbitrary_message_trees(obj):
    """Whatever a control-plane message carries (nested containers, scalars, tensors of any dtype / rank incl. empty ones): both
    wire forms — the plain pickle and the segmented form with out-of-band tensor memory — decode to an equal object."""
    from split_learning_b200.transport import codec
    assert _same(codec.loads(codec.dumps(obj)), obj)
    assert _same(codec.loads(bytearray(b"".join(codec.dump_segments(obj)))), obj)
    assert _same(codec.loads(bytes(b"".join(codec.dump_segments(obj)))), obj)


@FAST
@given(st.sampled_from(["os.system", "subprocess.Popen", "builtins.eval", "builtins.exec", "builtins.__import__", "posix.system",
                        "shutil.rmtree", "torch.load", "pickle.loads", "importlib.import_module"]), st.text(max_size=10))
def test_codec_refuses_every_foreign_global(target, arg):
    """The restricted unpickler resolves a fixed allow-list; any other global — whatever its arguments — is refused before it is
    called."""
    import pickle
    from split_learning_b200.transport import codec
    mod, name = target.rsplit(".", 1)
    payload = b"\x80\x04c" + mod.encode() + b"\n" + name.encode() + b"\n" + pickle.dumps((arg,), protocol=2)[2:-1] + b"R."
    try:                                            # the payload is a well-formed pickle: stock pickle would call the global
        assert pickle.loads(b"\x80\x04cbuiltins\nlen\n" + pickle.dumps((arg,), protocol=2)[2:-1] + b"R.") == len(arg)
    except Exception as err:                        # noqa
        raise AssertionError(f"test payload malformed: {err}")
    import pytest
    with pytest.raises(Exception) as e:
        codec.loads(payload)
    assert isinstance(e.value, codec.UnsafePayload), e.value

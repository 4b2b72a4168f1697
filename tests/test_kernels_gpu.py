"""Every sm_100a kernel against a plain PyTorch fp32 reference (run on the B200 box)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def selftest():
    from split_learning_b200.ops import native, selftest
    native.require()
    return selftest


@pytest.mark.parametrize("name", ["conv_fwd", "conv_dgrad", "conv_dgrad_bn_stats", "conv_wgrad", "linear_all", "bn_fwd_bwd", "conv1_direct",
                                  "ce_and_linear_epilogues", "optimizers_and_fedavg", "flags_and_peer"])
def test_kernel(selftest, name):
    err, tol = selftest.CHECKS[name]()
    assert err <= tol, f"{name}: err {err} > tol {tol}"

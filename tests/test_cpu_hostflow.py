"""CPU: the host layer above the C ABI, with every kernel call replaced by a recorder.  The tensors hold garbage
(no kernel runs), but module wiring, autograd plumbing, shapes, the fusion decisions and the per-step ABI call
sequence are exercised for both networks — forward, backward and eval mode — without a GPU."""
import ctypes
import types

import pytest
import torch


@pytest.fixture
def recorder(monkeypatch, lib_built):
    from npp_b200 import _lib as L
    from npp_b200 import functional as F_
    from npp_b200.core import criterion as CR
    calls = {}

    def fake_call(name, *a, **k):
        assert hasattr(L.lib(), name), "ABI symbol %s is not exported" % name
        calls[name] = calls.get(name, 0) + 1

    for mod in (L, F_, CR):
        if hasattr(mod, "call"):
            monkeypatch.setattr(mod, "call", fake_call)
        if hasattr(mod, "stream"):
            monkeypatch.setattr(mod, "stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(F_, "is_internal", lambda t: (t.dim() == 4 and t.dtype in (torch.bfloat16, torch.float32)
                                                       and t.shape[1] % 8 == 0 and t.stride(1) == 1))
    F_.set_compute_dtype(torch.float32)
    yield calls
    F_.set_compute_dtype(torch.bfloat16)


def _cfg(L, C):
    ns = types.SimpleNamespace
    return ns(DATASET=ns(NUM_CLASSES=20, NUM_JOINTS=16), TRAIN=ns(LAYERS=L, INIT_CHANNELS=C),
              SEARCH=ns(LAYERS=L, INIT_CHANNELS=C), MODEL=ns(DECONV_WITH_BIAS=False, HEAD="PSP", REFINE_LAYERS=1))


def _check_outputs(pl, par, n, hw):
    shapes = [tuple(t.shape) for p in pl + par for t in p]
    assert shapes == [(n, 16, hw, hw)] * 4 + [(n, 20, hw, hw), (n, 2, hw, hw)] * 2


def test_derived_network_call_flow(recorder):
    from npp_b200.models.model_augment import Network
    net = Network(_cfg(8, 16)).train()
    pl, par = net(torch.randn(2, 3, 64, 64))
    _check_outputs(pl, par, 2, 16)
    sum(t.sum() for p in pl + par for t in p).backward()
    n_conv = sum(1 for m in net.modules() if isinstance(m, torch.nn.Conv2d) and m.groups == 1 and m.kernel_size != (1, 1)
                 or isinstance(m, torch.nn.Conv2d) and m.groups == 1)
    # every dense conv the forward used has exactly one fprop, one wgrad and (except the two image stems) one dgrad
    assert recorder["npp_conv2d_direct_fwd"] == recorder["npp_conv2d_direct_wgrad"]
    assert recorder["npp_conv2d_direct_dgrad"] == recorder["npp_conv2d_direct_fwd"] - 2
    assert recorder["npp_conv2d_direct_fwd"] <= n_conv
    # cell nodes: one fused forward pass (BatchNorm finalize folded in: npp_node_fwd_bn) and the two-kernel backward
    # (striped totals: no partials buffer, no fold kernel, no separate finalize launch per BatchNorm)
    fwd = recorder.get("npp_node_fwd", 0) + recorder.get("npp_node_fwd_bn", 0)
    bwd = recorder.get("npp_node_bwd_apply", 0) + recorder.get("npp_node_bwd_apply_striped", 0)
    assert fwd > 100 and bwd > 100
    assert recorder.get("npp_node_fwd_bn", 0) > 100
    # every consumer of a cell state (further primitives, the concat route) has its own handle: the extra gradients
    # reach the node's backward kernel as typed slots (npp_node_bwd_reduce3) instead of autograd add kernels
    assert recorder.get("npp_node_bwd_reduce3", 0) > 60
    assert recorder.get("npp_node_bwd_reduce", 0) + recorder.get("npp_node_bwd_reduce3", 0) >= bwd
    # parameters never used by the forward get no gradient (SE_Block.bn at stride 1: SURVEY.md §5, find_unused_parameters)
    unused = [k for k, p in net.named_parameters() if p.grad is None]
    assert unused and all(".bn." in k for k in unused)
    with torch.no_grad():
        _check_outputs(*net.eval()(torch.randn(1, 3, 64, 64)), 1, 16)


def test_search_supernet_call_flow(recorder):
    from npp_b200.models.model_search_interact import Network
    net = Network(_cfg(8, 16)).train()
    pl, par = net(torch.randn(2, 3, 64, 64))
    _check_outputs(pl, par, 2, 16)
    (sum(t.sum() for p in pl + par for t in p) + net.loss_entropy()).backward()
    n_mixed = sum(1 for m in net.modules() if type(m).__name__ == "MixedOp")
    assert n_mixed == 2 * (10 + 18) + 6 * 18
    # one fused mix per MixedOp plus one per weighted node; each has its two backward passes
    assert recorder["npp_mix_fwd"] >= n_mixed
    assert recorder["npp_mix_bwd_reduce"] == recorder["npp_mix_bwd_apply"]
    assert recorder["npp_mix_dw"] == recorder["npp_mix_bwd_reduce"]
    assert all(p.grad is not None for p in net.arch_parameters())
    with torch.no_grad():
        _check_outputs(*net.eval()(torch.randn(1, 3, 64, 64)), 1, 16)

"""Training-step glue kernels (row a19) on the B200: fused multi-tensor Adam against torch.optim.Adam, the
Reptile-style outer update against the reference's per-parameter loop (train.py:524-534)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trainable():
    from mastermetastyletransfer_b200 import Decoder, StyleTransformer, synthetic
    st = StyleTransformer(256, 256, 8, 8, [8, 8], [8, 8], [4, 4], [4, 4])
    dec = Decoder()
    synthetic.fill_state_dict_(st, 0)
    synthetic.fill_state_dict_(dec, 0)
    return st.cuda(), dec.cuda()


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(trainable, wd):
    from mastermetastyletransfer_b200.optim import FusedAdam
    st, dec = trainable
    mine = [copy.deepcopy(st), copy.deepcopy(dec)]
    ref = [copy.deepcopy(st), copy.deepcopy(dec)]
    pm = [p for m in mine for p in m.parameters()]
    pr = [p for m in ref for p in m.parameters()]
    assert sum(p.numel() for p in pm) == 4300859  # the reference's trainable parameter count
    opt_m = FusedAdam(pm, lr=1e-3, weight_decay=wd)
    opt_r = torch.optim.Adam(pr, lr=1e-3, weight_decay=wd)
    g = torch.Generator(device="cuda").manual_seed(0)
    for step in range(5):
        for a, b in zip(pm, pr):
            grad = torch.randn(a.shape, generator=g, device="cuda") * 0.1
            a.grad, b.grad = grad.clone(), grad.clone()
        v0 = pm[0]._version
        opt_m.step()
        opt_r.step()
        assert pm[0]._version > v0  # packed-weight caches key on the version counter
    for a, b in zip(pm, pr):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (a - b).abs().max()


def test_reptile_update_matches_reference_loop(trainable):
    from mastermetastyletransfer_b200.optim import reptile_update
    st, _ = trainable
    theta, omega, theta_ref = copy.deepcopy(st), copy.deepcopy(st), copy.deepcopy(st)
    with torch.no_grad():
        for p in omega.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        # train.py:524-534
        for (n1, p1), (n2, p2) in zip(theta_ref.named_parameters(), omega.named_parameters()):
            p1.data += 1e-2 * (p2.data - p1.data)
    reptile_update(theta, omega, 1e-2)
    for a, b in zip(theta.parameters(), theta_ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    # integer buffers (relative_position_index) are not parameters and must be untouched
    assert torch.equal(theta.encoder.shared_MHA_without_MLP.attn.relative_position_index,
                       st.encoder.shared_MHA_without_MLP.attn.relative_position_index)


def test_forward_sees_optimizer_updates(trainable):
    """The forward path repacks its bf16 weights when the optimiser kernels change a parameter."""
    from mastermetastyletransfer_b200.optim import reptile_update
    st, _ = trainable
    theta, omega = copy.deepcopy(st).eval(), copy.deepcopy(st).eval()
    x = torch.randn(1, 16, 16, 256, device="cuda")
    with torch.no_grad():
        before = theta(x, x, 1).clone()
        for p in omega.parameters():
            p.mul_(1.05)
        reptile_update(theta, omega, 1.0)  # theta <- omega
        after = theta(x, x, 1)
        want = omega(x, x, 1)
    assert not torch.equal(before, after)
    assert torch.allclose(after, want, rtol=1e-3, atol=1e-3)

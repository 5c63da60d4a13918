"""Checkpoint / resume of the fused Adam on the B200: stopping after two steps, saving state_dict(), loading it
into a NEW FusedAdam over a copy of the parameters and continuing must give bit-identical parameters to the uninterrupted run
(same kernel, same moments, same step count for the bias correction); resuming a torch.optim.Adam from the same checkpoint
stays within the fp32 tolerance of tests/test_gpu_optim.py.  Also the capturable (device-side {lr, step}) variant."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in ((256, 256), (1024,), (3, 3, 64), (225, 8))]


def _grads(step, params):
    g = torch.Generator().manual_seed(100 + step)
    for p in params:
        p.grad = (torch.randn(p.shape, generator=g) * 0.1).cuda()


@pytest.mark.parametrize("capturable", [False, True])
def test_resume_is_bit_identical_to_the_uninterrupted_run(capturable):
    from mastermetastyletransfer_b200.optim import FusedAdam
    kw = dict(lr=1e-3, weight_decay=0.01, capturable=capturable)
    straight = _params()
    opt = FusedAdam(straight, **kw)
    for s in range(4):
        _grads(s, straight)
        opt.step()
    first = _params()
    opt1 = FusedAdam(first, **kw)
    for s in range(2):
        _grads(s, first)
        opt1.step()
    ckpt = copy.deepcopy(opt1.state_dict())
    assert all(float(e["step"]) == 2.0 for e in ckpt["state"].values()) and len(ckpt["state"]) == 4
    resumed = [torch.nn.Parameter(p.detach().clone()) for p in first]
    opt2 = FusedAdam(resumed, lr=5.0, capturable=capturable)  # hyper-parameters come from the checkpoint
    opt2.load_state_dict(ckpt)
    ref = [torch.nn.Parameter(p.detach().clone()) for p in first]
    opt_ref = torch.optim.Adam(ref, lr=5.0)
    plain = copy.deepcopy(ckpt)
    for g in plain["param_groups"]:
        g["capturable"] = False  # the torch reference steps eagerly
    opt_ref.load_state_dict(plain)
    for s in range(2, 4):
        _grads(s, resumed)
        opt2.step()
        _grads(s, ref)
        opt_ref.step()
    torch.cuda.synchronize()
    for a, b, c in zip(straight, resumed, ref):
        assert torch.equal(a, b)
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-7), (a - c).abs().max()
    assert opt2.state_dict()["state"][0]["step"].item() == 4.0


def test_fast_adaptation_step_updates_the_style_encoder_only():
    """The reference's few-shot stage (train_only_inner_loop.py:306-318) through one eager inner-loop step: the style encoder's
    gradients are those of the ordinary step (same weights, same batch), nothing else moves."""
    from mastermetastyletransfer_b200 import MasterStyleTransferModel, custom_loss, synthetic
    from mastermetastyletransfer_b200.training import InnerLoopTrainer
    loss_fn = custom_loss("/nonexistent")
    synthetic.fill_state_dict_(loss_fn, 1)
    loss_fn = loss_fn.cuda()
    content, style = synthetic.synthetic_images(2, 64, seed=9)
    content, style = content.cuda(), style.cuda()
    grads, trainers = [], []
    for fast in (False, True):
        m = MasterStyleTransferModel()
        synthetic.fill_state_dict_(m, 0)
        m = m.cuda().eval()
        for mod in (m.style_transformer.encoder, m.style_transformer.decoder):
            mod.stochastic_depth.p = 0.0
        m.style_transformer.encoder.encoder_stochastic_depth_prob = 0.0
        tr = InnerLoopTrainer(m, loss_fn, inner_lr=1e-3, fast_adaptation=fast)
        before = [p.detach().clone() for p in tr.params]
        losses = tr.step(content, style, 1)
        assert torch.isfinite(losses).all()
        grads.append([p.grad.detach().clone() for p in tr.omega_st.encoder.parameters()])
        trainers.append((tr, before))
    for a, b in zip(*grads):  # fp32 atomics in the weight-gradient kernels: order noise only
        assert ((a - b).norm() / (a.norm() + 1e-12)).item() < 2e-3
    tr, before = trainers[1]
    n_enc = len(list(tr.omega_st.encoder.parameters()))
    moved = [not torch.equal(p.detach(), q) for p, q in zip(tr.params, before)]
    enc_ids = {id(p) for p in tr.omega_st.encoder.parameters()}
    for p, mv in zip(tr.params, moved):
        assert mv == (id(p) in enc_ids) or (id(p) in enc_ids and p.grad.abs().max().item() == 0.0)
    assert sum(moved) >= n_enc - 1
    assert all(p.grad is None for p in tr.omega_dec.parameters())

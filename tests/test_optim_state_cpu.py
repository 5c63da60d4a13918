"""CPU suite: FusedAdam checkpoint / resume.  The state_dict layout is torch.optim.Adam's, so a run can move
between the two optimisers; the update kernel itself is checked on the B200 (tests/test_gpu_optim.py)."""
import pytest
import torch

from mastermetastyletransfer_b200.optim import FusedAdam


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((4, 3), (5,), (2, 2, 3))]


def _torch_adam_after(steps, params, **kw):
    opt = torch.optim.Adam(params, **kw)
    g = torch.Generator().manual_seed(9)
    for _ in range(steps):
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
    return opt


def test_state_dict_round_trips_through_torch_adam():
    ref = _torch_adam_after(3, _params(), lr=2e-4, betas=(0.8, 0.99), eps=1e-7, weight_decay=0.01)
    sd = ref.state_dict()
    mine = FusedAdam(_params(), lr=1.0)
    mine.load_state_dict(sd)
    assert mine.step_count == 3
    g = mine.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (2e-4, (0.8, 0.99), 1e-7, 0.01)
    for i in range(3):
        assert torch.equal(mine.state[0]["exp_avg"][i], sd["state"][i]["exp_avg"])
        assert torch.equal(mine.state[0]["exp_avg_sq"][i], sd["state"][i]["exp_avg_sq"])
    out = mine.state_dict()
    assert set(out["param_groups"][0]) == set(sd["param_groups"][0])  # same keys as torch.optim.Adam packs
    back = torch.optim.Adam(_params(), lr=1.0)
    back.load_state_dict(out)  # torch accepts it
    for i, p in enumerate(back.param_groups[0]["params"]):
        assert float(back.state[p]["step"]) == 3.0
        assert torch.equal(back.state[p]["exp_avg"], sd["state"][i]["exp_avg"])
    assert back.param_groups[0]["lr"] == 2e-4


def test_moments_are_loaded_in_place_and_groups_are_checked():
    mine = FusedAdam([{"params": _params()[:2], "lr": 1e-3}, {"params": _params()[2:], "lr": 1e-4}])
    before = [t.data_ptr() for st in mine.state for t in st["exp_avg"] + st["exp_avg_sq"]]
    fresh = mine.state_dict()
    assert fresh["state"] == {} and [g["params"] for g in fresh["param_groups"]] == [[0, 1], [2]]
    sd = _torch_adam_after(2, _params()).state_dict()
    with pytest.raises(ValueError):
        mine.load_state_dict(sd)  # one group vs two
    p = _params()
    two = torch.optim.Adam([{"params": p[:2], "lr": 1e-3}, {"params": p[2:], "lr": 5e-5}])
    for q in p:
        q.grad = torch.ones_like(q)
    two.step()
    mine.load_state_dict(two.state_dict())
    assert [t.data_ptr() for st in mine.state for t in st["exp_avg"] + st["exp_avg_sq"]] == before
    assert mine.step_count == 1 and mine.param_groups[1]["lr"] == 5e-5
    sd = two.state_dict()
    sd["state"][2]["step"] = torch.tensor(7.0)
    with pytest.raises(ValueError):
        mine.load_state_dict(sd)  # the kernel keeps one step count
    sd = two.state_dict()
    sd["param_groups"][0]["amsgrad"] = True
    with pytest.raises(ValueError):
        mine.load_state_dict(sd)


def test_trainer_checkpoint_round_trip(tmp_path):
    """InnerLoopTrainer.state_dict(): theta, omega, optimiser and the shared layer-count sampler, through torch.save / load."""
    import mastermetastyletransfer_b200 as mst
    from mastermetastyletransfer_b200 import synthetic
    from mastermetastyletransfer_b200.training import InnerLoopTrainer
    model = synthetic.fill_state_dict_(mst.MasterStyleTransferModel(), 0)
    tr = InnerLoopTrainer(model, loss_fn=None, inner_lr=3e-4, seed=5)
    with torch.no_grad():
        for p in tr.params:
            p.add_(0.01)
        for st in tr.opt.state:
            for t in st["exp_avg"]:
                t.fill_(0.5)
    tr.opt.step_count = 7
    [tr._rng.randint(1, 4) for _ in range(3)]
    path = tmp_path / "trainer.pt"
    torch.save(tr.state_dict(), path)
    assert set(tr.state_dict()["style_transformer"]) == set(model.style_transformer.state_dict())  # the reference's file layout
    want_next = [tr._rng.randint(1, 4) for _ in range(5)]

    model2 = synthetic.fill_state_dict_(mst.MasterStyleTransferModel(), 1)
    tr2 = InnerLoopTrainer(model2, loss_fn=None, seed=0)
    ptrs = [p.data_ptr() for p in tr2.params]
    tr2.load_state_dict(torch.load(path, weights_only=False))
    assert [p.data_ptr() for p in tr2.params] == ptrs
    for a, b in zip(tr.params, tr2.params):
        assert torch.equal(a, b)
    for ma, mb in ((model.style_transformer, model2.style_transformer), (model.decoder, model2.decoder)):
        for a, b in zip(ma.parameters(), mb.parameters()):  # theta; the frozen Swin encoder is not part of the checkpoint
            assert torch.equal(a, b)
    assert tr2.opt.step_count == 7 and tr2.opt.lr == 3e-4
    assert all(torch.equal(t, torch.full_like(t, 0.5)) for st in tr2.opt.state for t in st["exp_avg"])
    assert [tr2._rng.randint(1, 4) for _ in range(5)] == want_next


def test_fast_adaptation_freezes_everything_but_the_style_encoder():
    """train_only_inner_loop.py:306-318: the few-shot stage trains the style encoder only."""
    import mastermetastyletransfer_b200 as mst
    from mastermetastyletransfer_b200.training import InnerLoopTrainer
    model = mst.MasterStyleTransferModel()
    tr = InnerLoopTrainer(model, loss_fn=None, fast_adaptation=True)
    enc = list(tr.omega_st.encoder.parameters())
    assert all(p.requires_grad for p in enc)
    assert not any(p.requires_grad for p in tr.omega_st.decoder.parameters())
    assert not any(p.requires_grad for p in tr.omega_dec.parameters())
    assert not any(p.requires_grad for p in model.swin_encoder.parameters())
    assert [id(p) for p in tr.opt.params] == [id(p) for p in enc]  # Adam updates exactly the style encoder
    assert sum(p.numel() for p in tr.opt.params) == sum(p.numel() for p in model.style_transformer.encoder.parameters())
    assert len(tr.params) == len(list(model.style_transformer.parameters())) + len(list(model.decoder.parameters()))  # omega <- theta copies all

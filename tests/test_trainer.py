"""Training-driver parity and the Grad-CAM consumer contract (rank 4) against fixtures
written by the UNMODIFIED reference model (oracle/make_golden_train.py): a checkpoint in train.py's format resumes on
the B200 path and the next optimisation step lands where the reference's own next step landed; the hooks Grad-CAM
registers see the reference's activations and gradients."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _data(step):
    g = torch.Generator().manual_seed(100 + step)
    return torch.randn(4, 3, 32, 32, generator=g), torch.randint(0, 3, (4,), generator=g)


def test_checkpoint_format_and_resume_semantics(tmp_path):
    """CPU: the dictionary has train.py:310-319's keys; resume restores weights / optimizer / scheduler / epoch /
    best_acc and tolerates every part but the weights being absent (train.py:208-249)."""
    from medmamba_b200 import trainer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    opt = trainer.build_optimizer(net, lr=1e-3)
    assert opt.defaults["weight_decay"] == 1e-4 and trainer.build_optimizer(net, npz_dataset=True).defaults["weight_decay"] == 1e-2
    sched = trainer.build_scheduler(opt, [1, 2])
    assert trainer.build_scheduler(opt, None) is None
    x, y = torch.randn(6, 5), torch.randint(0, 3, (6,))
    loss0 = trainer.train_step(net, x, y, opt)
    sched.step()
    path = str(tmp_path / "sub" / "ck.pth")
    d = trainer.save_checkpoint(path, 4, net, opt, 0.75, 3, {0: "a"}, scheduler=sched)
    assert set(d) == {"epoch", "model_state_dict", "optimizer_state_dict", "best_acc", "num_classes", "class_indices",
                      "scheduler_state_dict"}
    torch.manual_seed(1)
    net2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    opt2 = trainer.build_optimizer(net2, lr=1e-3)
    sched2 = trainer.build_scheduler(opt2, [1, 2])
    start, best, _ = trainer.resume(path, net2, opt2, sched2)
    assert (start, best) == (5, 0.75) and sched2.last_epoch == 1 and opt2.param_groups[0]["lr"] == pytest.approx(1e-4)
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.equal(a, b)
    l1, l2 = trainer.train_step(net, x, y, opt), trainer.train_step(net2, x, y, opt2)
    assert torch.equal(l1, l2) and l1 < loss0 + 1.0
    for a, b in zip(net.parameters(), net2.parameters()):
        assert torch.equal(a, b), "the optimizer state did not survive the round trip"
    torch.save({"model_state_dict": net.state_dict()}, path)
    start, best, _ = trainer.resume(path, net2, opt2, sched2)
    assert (start, best) == (1, 0.0)


def test_reference_checkpoint_loads_into_the_mirror():
    """CPU: the reference-written checkpoint's keys and shapes are exactly the mirror's (strict load_state_dict)."""
    import medmamba_b200 as mm
    from medmamba_b200 import trainer
    ck = torch.load(os.path.join(GOLD, "ref_train_tiny.pt"), weights_only=False)
    net = mm.VSSM(**ck["expected"]["cfg"])
    opt = trainer.build_optimizer(net, lr=1e-3)
    sched = trainer.build_scheduler(opt, [1, 3])
    start, best, raw = trainer.resume(os.path.join(GOLD, "ref_train_tiny.pt"), net, opt, sched)
    assert (start, best) == (2, 0.625) and raw["num_classes"] == 3 and raw["class_indices"][2] == "c"
    assert opt.param_groups[0]["lr"] == pytest.approx(ck["expected"]["lr_after_resume"])
    sd = net.state_dict()
    assert all(torch.equal(sd[k], v) for k, v in ck["model_state_dict"].items()) and set(sd) == set(ck["model_state_dict"])
    assert len(opt.state) == len(list(net.parameters()))


@pytest.mark.gpu
def test_resume_reference_checkpoint_and_continue_on_gpu():
    """The reference's checkpoint resumes on the GPU path; the restored model gives the reference's logits, and the
    next AdamW step (train.py:277-288) reproduces the reference's next loss and weights: the optimizer moments carried
    over and the hand-written backward produced the reference's gradients."""
    import medmamba_b200 as mm
    from medmamba_b200 import trainer
    path = os.path.join(GOLD, "ref_train_tiny.pt")
    exp = torch.load(path, weights_only=False)["expected"]
    net = mm.VSSM(**exp["cfg"]).cuda()
    opt = trainer.build_optimizer(net, lr=1e-3)
    sched = trainer.build_scheduler(opt, [1, 3])
    start, best, ck = trainer.resume(path, net, opt, sched, device="cuda")
    assert start == 2
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net.eval()
        with torch.no_grad():
            logits = net(_data(99)[0].cuda())
        assert torch.allclose(logits.cpu(), exp["probe_logits"], rtol=1e-3, atol=1e-4)
        net.train()
        x, y = _data(2)
        loss = trainer.train_step(net, x.cuda(), y.cuda(), opt)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert abs(float(loss) - exp["loss_step3"]) < 1e-4 * max(1.0, abs(exp["loss_step3"]))
    sd = net.state_dict()
    for key, ck_key in (("head.weight", "head.weight"), ("x_proj_weight", "layers.0.blocks.0.self_attention.x_proj_weight"),
                        ("A_logs", "layers.0.blocks.0.self_attention.A_logs")):
        want, before, got = exp[key], ck["model_state_dict"][ck_key].cpu(), sd[ck_key].cpu()
        update = (want - before).abs().max().item()
        assert update > 0
        assert (got - want).abs().max().item() < 0.05 * update, f"{key}: step differs from the reference's by more than 5% of the update"


@pytest.mark.gpu
def test_gradcam_hooks_see_the_reference_activations_and_gradients():
    """test.py:101-108 / grad_cam/utils.py:5-49 on the GPU path against what the unmodified reference recorded on the
    same weights and image: forward and full-backward hooks on layers[-1].blocks[-1].conv33conv33conv11[-2], batch 1,
    backward from the top logit."""
    import medmamba_b200 as mm
    z = np.load(os.path.join(GOLD, "ref_gradcam_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    net = mm.VSSM(depths=[1, 1], dims=[16, 32], num_classes=3, drop_path_rate=0.0)
    net.load_state_dict(sd)
    net = net.cuda().eval()
    target = net.layers[-1].blocks[-1].conv33conv33conv11[-2]
    acts, grads = [], []
    h1 = target.register_forward_hook(lambda m, i, o: acts.append(o.detach()))
    h2 = target.register_full_backward_hook(lambda m, gi, go: grads.append(go[0].detach()))
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        logits = net(torch.from_numpy(z["x"]).cuda())
        net.zero_grad()
        logits[0, logits.argmax()].backward()
    finally:
        h1.remove(); h2.remove()
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    assert len(acts) == 1 and len(grads) == 1
    want_l, want_a, want_g = (torch.from_numpy(z[k]) for k in ("logits", "activation", "gradient"))
    assert torch.allclose(logits.detach().cpu(), want_l, rtol=1e-3, atol=1e-4)
    assert int(logits.argmax()) == int(want_l.argmax())
    assert torch.allclose(acts[0].cpu(), want_a, rtol=1e-3, atol=1e-4)
    err = (grads[0].cpu() - want_g).abs().max().item() / want_g.abs().max().item()
    assert err < 5e-3, err

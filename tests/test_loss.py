"""pmf_b200.loss.TrainerLoss — the trainer's loss block (tasks/pmf/trainer.py:305-332) as one module.

CPU: the plain-PyTorch implementation (what the unchanged trainer runs on our outputs) against the oracle restatement,
which tests/test_oracle_pinning.py pins to the reference's own loss classes.  GPU (-m gpu): the fused CUDA implementation
(pmfb_loss_head + pmfb_lovasz) against the same oracle: value within 1e-5, gradients within 1e-4 of the largest gradient
entry."""
import pytest
import torch

from oracle import loss_oracle as lo
from pmf_b200.loss import TrainerLoss


def _case(B, C, H, W, seed, density, sharp):
    g = torch.Generator().manual_seed(seed)
    lid = torch.softmax(torch.randn(B, C, H, W, generator=g) * sharp, 1)
    cam = torch.softmax(torch.randn(B, C, H, W, generator=g) * sharp * 0.7, 1)
    label = torch.randint(1, C, (B, H, W), generator=g) * (torch.rand(B, H, W, generator=g) < density)
    alpha = torch.rand(C, generator=g) * 0.9 + 0.1
    alpha[0] = 0
    return lid, cam, label.long(), alpha


CASES = [dict(B=2, C=20, H=16, W=32, seed=1, density=0.5, sharp=1.0), dict(B=2, C=20, H=24, W=40, seed=2, density=0.1, sharp=6.0),
         dict(B=1, C=17, H=32, W=48, seed=3, density=0.9, sharp=3.0)]


def _oracle(lid, cam, label, alpha, C):
    a = lid.clone().requires_grad_(True)
    b = cam.clone().requires_grad_(True)
    loss = lo.total_loss(a, b, label, alpha, C, 1.0, 0.5, 0.7)
    loss.backward()
    return loss.detach(), a.grad, b.grad


@pytest.mark.parametrize("case", CASES, ids=lambda c: "C%d_%dx%d" % (c["C"], c["H"], c["W"]))
def test_torch_impl_equals_oracle_cpu(case):
    lid, cam, label, alpha = _case(**case)
    ref, ga, gb = _oracle(lid, cam, label, alpha, case["C"])
    a, b = lid.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    crit = TrainerLoss(case["C"], alpha, 1.0, 0.5, 0.7, impl="torch")
    loss = crit(a, b, label)
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    assert torch.allclose(a.grad, ga, rtol=1e-5, atol=1e-9) and torch.allclose(b.grad, gb, rtol=1e-5, atol=1e-9)
    assert set(crit.last) == {"focal", "lovasz", "focal_cam", "lovasz_cam", "perception"}


GPU_CASES = CASES + [dict(B=2, C=20, H=96, W=160, seed=4, density=0.1, sharp=8.0), dict(B=1, C=20, H=64, W=2048, seed=5, density=0.3, sharp=4.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: "C%d_%dx%d" % (c["C"], c["H"], c["W"]))
def test_fused_loss_matches_oracle(case):
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    dev = torch.device("cuda:0")
    lid, cam, label, alpha = _case(**case)
    ref, ga, gb = _oracle(lid, cam, label, alpha, case["C"])
    a, b = lid.to(dev).requires_grad_(True), cam.to(dev).requires_grad_(True)
    crit = TrainerLoss(case["C"], alpha, 1.0, 0.5, 0.7, impl="fused").to(dev)
    loss = crit(a, b, label.to(dev))
    (loss * 1.0).backward()
    # components against the oracle's pieces
    per, _, _ = lo.perception_aware_loss(lid, cam, case["C"], 0.7)
    want = dict(focal=lo.focal_loss(lid, label, alpha), lovasz=lo.lovasz_softmax(lid, label), focal_cam=lo.focal_loss(cam, label, alpha),
                lovasz_cam=lo.lovasz_softmax(cam, label), perception=per)
    for k, v in want.items():
        assert abs(float(crit.last[k]) - float(v)) < 1e-5 * max(1.0, abs(float(v))), (k, float(crit.last[k]), float(v))
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref))), (float(loss), float(ref))
    for got, want_g, nm in ((a.grad.cpu(), ga, "lidar"), (b.grad.cpu(), gb, "camera")):
        err = float((got - want_g).abs().max() / want_g.abs().max())
        assert err < 1e-4, (nm, err)
    # the torch implementation on the GPU gives the same number (it is the comparison arm of bench.py)
    t = TrainerLoss(case["C"], alpha, 1.0, 0.5, 0.7, impl="torch").to(dev)
    assert abs(float(t(lid.to(dev), cam.to(dev), label.to(dev))) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))


@pytest.mark.gpu
def test_fused_loss_ties_no_grad_and_full_size():
    """Saturated probabilities (exact error ties: the class losses do not depend on the order inside a tie), the
    no-gradient (validation) path, and the BASELINE batch (8 x 20 x 480 x 640) against the torch implementation."""
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    C = 20
    hard = torch.nn.functional.one_hot(torch.randint(0, C, (2, 16, 24), generator=g), C).permute(0, 3, 1, 2).float()
    soft = torch.softmax(torch.randn(2, C, 16, 24, generator=g), 1)
    lid = torch.where(torch.rand(2, 1, 16, 24, generator=g) < 0.5, hard, soft)
    cam = torch.softmax(torch.randn(2, C, 16, 24, generator=g) * 2, 1)
    label = torch.randint(0, C, (2, 16, 24), generator=g)
    alpha = torch.ones(C)
    alpha[0] = 0
    fused = TrainerLoss(C, alpha, impl="fused").to(dev)
    with torch.no_grad():
        got = fused(lid.to(dev), cam.to(dev), label.to(dev))
    want = dict(lovasz=lo.lovasz_softmax(lid, label), lovasz_cam=lo.lovasz_softmax(cam, label), focal=lo.focal_loss(lid, label, alpha))
    for k, v in want.items():
        assert abs(float(fused.last[k]) - float(v)) < 1e-5 * max(1.0, abs(float(v))), k
    assert bool(torch.isfinite(got))
    # BASELINE batch: fused vs the torch implementation, both on the device
    lid, cam, label, alpha = _case(8, 20, 480, 640, 11, 0.1, 5.0)
    a, b = lid.to(dev).requires_grad_(True), cam.to(dev).requires_grad_(True)
    loss = TrainerLoss(20, alpha, impl="fused").to(dev)(a, b, label.to(dev))
    loss.backward()
    a2, b2 = lid.to(dev).requires_grad_(True), cam.to(dev).requires_grad_(True)
    ref = TrainerLoss(20, alpha, impl="torch").to(dev)(a2, b2, label.to(dev))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    for got_g, ref_g in ((a.grad, a2.grad), (b.grad, b2.grad)):
        assert float((got_g - ref_g).abs().max() / ref_g.abs().max()) < 1e-4

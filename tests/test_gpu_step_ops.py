"""GPU: fused cross-entropy + top-k + meters, and the multi-tensor Adam / SGD updates (csrc/step_ops.cu),
against torch's own implementations (train.py:319-336, 493, 518-529; utils/utils.py:72-85)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _ref_accuracy(output, target, topk):      # utils/utils.py:72-85 restated
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand_as(pred.t()))
    return [correct[:k].reshape(-1).float().sum() * (100.0 / target.size(0)) for k in topk]


@pytest.mark.parametrize("n,c,topk", [(256, 1000, (1, 5)), (128, 10, (1, 5)), (7, 3, (1, 3)), (1, 1000, (1, 5))])
def test_cross_entropy_topk_matches_torch(n, c, topk):
    from bdbnn_b200.functional import cross_entropy_topk
    g = torch.Generator().manual_seed(n + c)
    z = torch.randn(n, c, generator=g) * 3.0
    t = torch.randint(0, c, (n,), generator=g)
    z[0, t[0]] = z[0].max() + 1.0                       # a certain hit
    zr = z.double().requires_grad_(True)
    lr = nn.functional.cross_entropy(zr, t)
    (lr * 0.7).backward()
    zd = z.cuda().requires_grad_(True)
    meters = torch.zeros(4, dtype=torch.float64, device="cuda")
    loss, (a1, a5) = cross_entropy_topk(zd, t.cuda(), topk, meters)
    (loss * 0.7).backward()
    torch.testing.assert_close(loss.cpu().double(), lr.detach(), rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(zd.grad.cpu().double(), zr.grad, rtol=1e-5, atol=1e-8)
    r1, r5 = _ref_accuracy(z, t, topk)
    assert abs(a1.item() - r1.item()) < 1e-4 and abs(a5.item() - r5.item()) < 1e-4
    assert a1.shape == (1,) and not a1.requires_grad
    # second call accumulates the meters: {loss*N, acc1*N, acc5*N, N}
    cross_entropy_topk(zd.detach(), t.cuda(), topk, meters)
    m = meters.cpu()
    assert m[3].item() == 2 * n
    assert abs(m[0].item() / m[3].item() - lr.item()) < 1e-5 and abs(m[1].item() / m[3].item() - r1.item()) < 1e-4


def test_cross_entropy_topk_ties_count_lower_index_first():
    from bdbnn_b200.functional import cross_entropy_topk
    z = torch.zeros(2, 6)
    t = torch.tensor([0, 5])                            # all logits equal: class 0 is rank 0, class 5 is rank 5
    _, (a1, a5) = cross_entropy_topk(z.cuda(), t.cuda(), (1, 5))
    assert a1.item() == 50.0 and a5.item() == 50.0


def _models():
    torch.manual_seed(0)
    def make():
        m = nn.Sequential(nn.Conv2d(3, 8, 3, bias=False), nn.BatchNorm2d(8), nn.Conv2d(8, 16, 3, bias=False),
                          nn.Flatten(), nn.LazyLinear(5))
        m(torch.zeros(1, 3, 8, 8))
        return m
    a = make()
    b = make()
    b.load_state_dict(a.state_dict())
    return a.cuda().to(memory_format=torch.channels_last), b.cuda().to(memory_format=torch.channels_last)


def _run(model, opt, steps, scale=1.0):
    g = torch.Generator().manual_seed(9)
    for _ in range(steps):
        x = torch.randn(4, 3, 8, 8, generator=g).cuda()
        y = torch.randint(0, 5, (4,), generator=g).cuda()
        opt.zero_grad()
        (nn.functional.cross_entropy(model(x), y) * scale).backward()
        opt.step()


@pytest.mark.parametrize("dataset", ["imagenet", "cifar10"])
def test_fused_optimizers_match_torch(dataset):
    """make_optimizer's CUDA optimizers (FusedAdam with the conv-only weight-decay group / FusedSGD with momentum)
    track torch.optim.Adam / SGD over several steps, keep torch's state layout, and fold grad_scale."""
    from bdbnn_b200.optim import FusedAdam, FusedSGD
    from bdbnn_b200.step import make_optimizer
    ma, mb = _models()
    oa = make_optimizer(ma, dataset, lr=0.05 if dataset == "cifar10" else 3e-3, weight_decay=1e-2)
    ob = make_optimizer(mb, dataset, lr=0.05 if dataset == "cifar10" else 3e-3, weight_decay=1e-2, fused=False)
    assert isinstance(oa, FusedSGD if dataset == "cifar10" else FusedAdam) and type(ob).__module__.startswith("torch.optim")
    assert ma[0].weight.is_contiguous(memory_format=torch.channels_last)
    _run(ma, oa, 4)
    _run(mb, ob, 4)
    for (n, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-5, atol=2e-6, msg=n)
    # state layout: torch's optimizer accepts our state_dict and vice versa
    import copy
    ob.load_state_dict(copy.deepcopy(oa.state_dict()))      # deepcopy: load_state_dict keeps same-device tensors
    oa.load_state_dict(copy.deepcopy(ob.state_dict()))
    _run(ma, oa, 1)
    _run(mb, ob, 1)
    for pa, pb in zip(ma.parameters(), mb.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-5, atol=2e-6)
    # grad_scale = 1/4 on 4x larger gradients gives the same update (the data-parallel averaging)
    mc, md = _models()
    oc, od = make_optimizer(mc, dataset, lr=1e-2), make_optimizer(md, dataset, lr=1e-2)
    oc.grad_scale = 0.25
    _run(mc, oc, 2, scale=4.0)
    _run(md, od, 2)
    for pa, pb in zip(mc.parameters(), md.parameters()):
        torch.testing.assert_close(pa, pb, rtol=2e-5, atol=2e-6)


def test_fused_adam_many_tensors_and_lr_schedule():
    """More tensors than one kernel table holds (48), LambdaLR driving param_groups['lr'] (train.py:336)."""
    from bdbnn_b200.optim import FusedAdam
    torch.manual_seed(1)
    pa = [torch.randn(5 + i, 3, device="cuda").requires_grad_(True) for i in range(110)]
    pb = [p.detach().clone().requires_grad_(True) for p in pa]
    oa, ob = FusedAdam(pa, lr=1e-2, weight_decay=1e-3), torch.optim.Adam(pb, lr=1e-2, weight_decay=1e-3)
    sa = torch.optim.lr_scheduler.LambdaLR(oa, lambda s: 1.0 - s / 10)
    sb = torch.optim.lr_scheduler.LambdaLR(ob, lambda s: 1.0 - s / 10)
    for it in range(3):
        for p, q in zip(pa, pb):
            gr = torch.randn_like(p)
            p.grad, q.grad = gr.clone(), gr.clone()
        oa.step(); ob.step(); sa.step(); sb.step()
    for p, q in zip(pa, pb):
        torch.testing.assert_close(p, q, rtol=2e-5, atol=2e-6)

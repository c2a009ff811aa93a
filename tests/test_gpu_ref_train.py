"""GPU: the CUDA path (product modules + fused loss kernels + fused optimizers, bdbnn_b200.step.TrainStep)
replays the fixtures produced by the REFERENCE's own train() / train_teacher_student()
(tests/golden/train_step_*.pt, see tests/golden/make_train_golden.py), in both gradient operand modes.

Includes the configuration the headline times: ResNet-18, 224x224, Adam, tcgen05 stem,
fused stem BN+pool, fp8 forward, in-node 1x1 shortcuts — CE only and CE + kurtosis + KD with an fp32
torchvision teacher — compared element by element with the oracle step re-run on the box's CPU and with
the stored reference summaries.

Tolerances (whole network, looser than the per-kernel tests): a handful of activations within fp32
round-off of 0 or +-1 take the other sign / STE-mask value on the GPU (different BatchNorm summation
order), a discrete difference that propagates; gradients are compared relative to max|reference| per tensor."""
import json
import os

import pytest
import torch

import train_cases as TC
from test_ref_train import SMALL, check_summaries, check_tensors, check_terms, load_golden, oracle_step

pytestmark = pytest.mark.gpu

TERM_RTOL = 1e-3
GRAD_TOL = {"fp16s": 2e-2, "bf16x2": 2e-2}      # of max|reference gradient| per tensor


def _product_step(case, gold, mode, monkeypatch):
    from bdbnn_b200.step import StepConfig, TrainStep, make_optimizer
    monkeypatch.setenv("BDBNN_GRAD_MODE", mode)
    c = TC.CASES[case]
    a = c["args"]
    ref_stud, ref_teacher = TC.build_oracle(case)
    if "init" in gold:
        ref_stud.load_state_dict(gold["init"])
        if ref_teacher is not None:
            ref_teacher.load_state_dict(gold["teacher_state"])
        x = gold["x"]
    else:
        x = TC.make_batch(case)[0]
    stud, teacher = TC.build_product(case)
    stud.load_state_dict(ref_stud.state_dict())
    stud = stud.cuda().to(memory_format=torch.channels_last)
    if teacher is not None:
        teacher.load_state_dict(ref_teacher.state_dict())
        teacher = teacher.cuda().to(memory_format=torch.channels_last).eval()
    targets = TC.step_kwargs(case, 19 if c["arch"] == "resnet18" else 6)["targets"]
    cfg = StepConfig(w_kurtosis=bool(a.get("w_kurtosis")), w_kurtosis_target=targets,
                     w_lambda_kurtosis=a.get("w_lambda_kurtosis", 1.0), kurtosis_mode=a.get("kurtosis_mode", "avg"),
                     teacher_student=c["teacher"], react=bool(a.get("react")), alpha=a.get("alpha", 0.9),
                     beta=a.get("beta", 200.0))
    step = TrainStep(stud, make_optimizer(stud, c["dataset"], lr=a["lr"]), cfg, teacher=teacher)
    if cfg.w_kurtosis:
        assert list(step.hooked) == gold["hooked"]
    out = step(x.cuda().contiguous(memory_format=torch.channels_last), gold["y"].cuda())
    torch.cuda.synchronize()
    return stud, step, out


def _errors(got, ref):
    return {n: (got[n].detach().cpu().float() - ref[n].detach().cpu().float()).abs().max().item() /
            (ref[n].abs().max().item() + 1e-12) for n in ref}


@pytest.mark.parametrize("mode", ["fp16s", "bf16x2"])
@pytest.mark.parametrize("case", SMALL)
def test_gpu_step_replays_reference_loop_fixture(case, mode, monkeypatch):
    from bdbnn_b200 import _lib
    gold = load_golden(case)
    n0 = _lib.launch_count()
    stud, step, out = _product_step(case, gold, mode, monkeypatch)
    assert _lib.launch_count() - n0 >= 6 * 4            # 6 binary convs: packs, fwd, dgrad, wgrad at least
    check_terms({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in out.items()}, gold, rtol=TERM_RTOL, atol=1e-5)
    check_tensors({n: p.grad for n, p in stud.named_parameters()}, gold["grads"], GRAD_TOL[mode], "grad")
    # parameters after the (fused) optimizer step: SGD, lr 0.05-0.1 -> error = lr * gradient error
    check_tensors(dict(stud.named_parameters()), gold["after"], 5e-3, "param")
    check_tensors({n: b for n, b in stud.named_buffers() if "num_batches" not in n},
                  {n: b for n, b in gold["buffers"].items() if "num_batches" not in n}, 1e-3, "buffer")
    av = step.averages()
    assert av["loss"] == pytest.approx(gold["meters"]["Loss"][-1], rel=TERM_RTOL)        # `losses` meter = total loss
    assert av["ce"] == pytest.approx(gold["meters"]["Loss_ce"][-1] if not gold["config"]["args"].get("react")
                                     else av["ce"], rel=TERM_RTOL)
    assert av["acc1"] == pytest.approx(gold["meters"]["Acc@1"][-1], abs=1e-3)


BLOCKS = ["stem"] + [f"layer{l}.{b}" for l in (1, 2, 3, 4) for b in (0, 1)] + ["head"]
STRIDED = ("layer2.0", "layer3.0", "layer4.0")        # carry the real-valued 1x1 `downsample` conv (TF32-class)


def _oracle_blockwise(case):
    """One oracle forward + backward of the ResNet-18 case on the CPU, recording for every block its input, its
    output, the gradient arriving at its output and the gradient leaving at its input."""
    ref, _ = TC.build_oracle(case)
    ref.train()
    x, y = TC.make_batch(case)
    rec = {}
    mods = dict(ref.named_modules())

    def run(name, fn, inp):
        inp = inp.detach().requires_grad_(name != "stem")
        out = fn(inp)
        rec[name] = {"in": inp, "out": out}
        return out

    h = run("stem", lambda t: ref.maxpool(ref.bn1(ref.conv1(t))), x)
    for name in BLOCKS[1:-1]:
        h = run(name, mods[name], h)
    logits = run("head", lambda t: ref.fc(ref.avgpool(t).flatten(1)), h)
    loss = torch.nn.functional.cross_entropy(logits, y)
    # chain the per-block graphs by hand: the gradient leaving block k is the gradient arriving at block k-1
    g = torch.autograd.grad(loss, logits)[0]
    for name in reversed(BLOCKS):
        r = rec[name]
        r["gout"] = g
        r["out"].backward(g)
        g = r["in"].grad
    return ref, x, y, rec, float(loss.detach())


def _metrics(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = got - ref
    return {"max": (err.abs().max() / (ref.abs().max() + 1e-300)).item(),
            "l2": (err.norm() / (ref.norm() + 1e-300)).item(),
            "frac_gt_1e-3": (err.abs() > 1e-3 * ref.abs().max()).double().mean().item()}


@pytest.mark.parametrize("mode", ["fp16s", "bf16x2"])
def test_gpu_resnet18_224_blockwise_teacher_forced(mode, monkeypatch):
    """The code path the headline times (ResNet-18, 224x224: tcgen05 stem + fused BN/pool, fp8 forward, fused
    conv-BN-add units with int16 y, in-node 1x1 shortcuts, classifier), checked block by block against the oracle
    with TEACHER FORCING: every product block receives the oracle's input of that block and the oracle's gradient
    of its output, so its output, its input gradient and its parameter gradients are comparable element by element.

    Why not end to end: a binarised network is a chaotic map.  One activation within round-off of 0 takes the other
    sign, changes 9*Cout conv outputs by 2*alpha, and after three or four layers a third of all signs differ
    (measured: scripts/diag_r18.py — 1e-6 of the signs differ after the TF32-class stem, 33 % after layer4).  That
    holds between ANY two implementations whose fp32 parts round differently (the reference on cuDNN/TF32 vs on a
    CPU included), so whole-network numbers are compared statistically (next test) and parity is stated per block.
    Tolerances: identity blocks 1e-4 forward (observed 6e-6); blocks with the real-valued 1x1 shortcut and the stem
    are TF32-class by design (fp16 operands, fp32 accumulate); gradients carry the mode's operand
    rounding (fp16s / bf16x2)."""
    monkeypatch.setenv("BDBNN_GRAD_MODE", mode)
    case = "r18_ce"
    ref, x, y, rec, loss_ref = _oracle_blockwise(case)
    prod, _ = TC.build_product(case)
    prod.load_state_dict(ref.state_dict())
    prod = prod.cuda().to(memory_format=torch.channels_last).train()
    pm = dict(prod.named_modules())
    report = {}
    gtol = {"fp16s": 4e-3, "bf16x2": 2e-4}[mode]
    for name in BLOCKS:
        r = rec[name]
        xin = r["in"].detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(name != "stem")
        if name == "stem":
            fn, params = prod._stem, {"conv1.weight": prod.conv1.weight, "bn1.weight": prod.bn1.weight,
                                      "bn1.bias": prod.bn1.bias}
        elif name == "head":
            fn, params = (lambda t: prod.fc(torch.flatten(prod.avgpool(t), 1))), \
                {"fc.weight": prod.fc.weight, "fc.bias": prod.fc.bias}
        else:
            fn, params = pm[name], {f"{name}.{k}": v for k, v in pm[name].named_parameters()}
        for p in params.values():
            p.grad = None
        out = fn(xin)
        out.backward(r["gout"].detach().cuda().contiguous(memory_format=torch.channels_last)
                     if r["gout"].dim() == 4 else r["gout"].cuda())
        ref_params = dict(ref.named_parameters())
        m = {"out": _metrics(out, r["out"])}
        if name != "stem":
            m["gin"] = _metrics(xin.grad, r["in"].grad)
        for pn, p in params.items():
            m[pn] = _metrics(p.grad, ref_params[pn].grad)
        report[name] = m
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/r18_blockwise_{mode}.json", "w") as fh:
        json.dump(report, fh, indent=1)
    bad = []
    for name, m in report.items():
        tf32 = name in STRIDED or name == "stem"
        fwd_max, fwd_l2 = (5e-2, 1e-2) if name in STRIDED else ((2e-3, 5e-4) if name == "stem" else (1e-4, 1e-5))
        if m["out"]["max"] > fwd_max or m["out"]["l2"] > fwd_l2:
            bad.append((name, "out", m["out"]))
        for k, v in m.items():
            if k == "out":
                continue
            # strided blocks: the shortcut's TF32-class output flips ~1e-3 of conv2's input signs (discrete), so
            # their gradients are compared in the L2 sense only.  Stem: the max-pool's arg-max is discrete too — with
            # TF32-class conv outputs ~1e-3 of the 3x3 windows pick another winner, which re-routes that share of the
            # gradient terms of every weight-gradient sum (relative L2 ~ sqrt(2f); measured 2.8e-2); the stem's conv
            # and BN+pool kernels are checked without this effect in test_gpu_tc.py (test_stem_*).
            if name == "stem" and k == "conv1.weight":
                lim_max, lim_l2 = None, 6e-2
            else:
                lim_max, lim_l2 = (None, 6e-2) if name in STRIDED else ((2e-2, 5e-3) if tf32 else (gtol, gtol))
            if (lim_max is not None and v["max"] > lim_max) or v["l2"] > lim_l2:
                bad.append((name, k, v))
    assert not bad, bad


@pytest.mark.parametrize("mode", ["fp16s"])
@pytest.mark.parametrize("case", list(TC.SUMMARY_ONLY))
def test_gpu_resnet18_224_whole_step_statistical(case, mode, monkeypatch):
    """Whole ResNet-18 step (batch 4, 224x224) on the CUDA path vs the reference loop's fixture.  The terms that
    depend on the weights only — the kurtosis regulariser over the 19 hooked layers and the per-layer KD term —
    must match the reference's numbers tightly; the activation-dependent terms (CE, KD on logits) of a randomly
    initialised binary network are only statistically comparable (see the block-wise test) and are bounded
    loosely; every parameter must receive a finite gradient of the reference's order of magnitude."""
    gold = load_golden(case)
    stud, step, out = _product_step(case, gold, mode, monkeypatch)
    m = gold["meters"]
    if "Loss_kurt" in m:
        assert float(out["kurt"]) == pytest.approx(m["Loss_kurt"][-1], rel=1e-4)
    if "Loss_kl" in m:
        assert float(out["kl"]) == pytest.approx(m["Loss_kl"][-1], rel=1e-4)
    assert float(out["ce"]) == pytest.approx(m["Loss_ce"][-1], rel=0.15)
    if "Loss_kl_c" in m:
        assert float(out["kl_c"]) == pytest.approx(m["Loss_kl_c"][-1], rel=0.15)
    assert float(out["loss"]) == pytest.approx(m["Loss"][-1], rel=0.15)
    for n, p in stud.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        ref = gold["grads"][n]
        got = float(p.grad.double().norm())
        assert 0.2 * ref["norm"] <= got <= 5.0 * ref["norm"] + 1e-12, (n, got, ref["norm"])



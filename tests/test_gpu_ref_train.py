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


@pytest.mark.parametrize("mode", ["fp16s", "bf16x2"])
@pytest.mark.parametrize("case", list(TC.SUMMARY_ONLY))
def test_gpu_resnet18_224_step_vs_reference_and_oracle(case, mode, monkeypatch):
    """The timed configuration's code path at batch 4: every loss term, the logits and EVERY parameter gradient
    against the oracle step (full tensors, recomputed on the CPU here) and the reference loop's summaries."""
    gold = load_golden(case)
    ref_stud, ref_out = oracle_step(case, gold)                 # also re-checks the seeded init checksum
    stud, step, out = _product_step(case, gold, mode, monkeypatch)
    check_terms({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in out.items()}, gold, rtol=TERM_RTOL, atol=1e-4)
    torch.testing.assert_close(out["output"].cpu(), ref_out["output"], rtol=5e-3, atol=5e-3)
    got = {n: p.grad for n, p in stud.named_parameters()}
    ref = {n: p.grad for n, p in ref_stud.named_parameters()}
    errs = _errors(got, ref)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/r18_step_grad_err_{case}_{mode}.json", "w") as fh:
        json.dump({"max": max(errs.values()), "worst": sorted(errs.items(), key=lambda kv: -kv[1])[:8]}, fh, indent=1)
    bad = {n: e for n, e in errs.items() if e > GRAD_TOL[mode]}
    assert not bad, bad
    check_summaries(got, gold["grads"], 2 * GRAD_TOL[mode], "grad-vs-reference-summary")

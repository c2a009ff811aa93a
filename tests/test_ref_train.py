"""Whole-step parity against the REFERENCE's own loop code.

tests/golden/train_step_<case>.pt were produced by tests/golden/make_train_golden.py, which imports
/root/reference/train.py unmodified (tests/ref_launcher.py) and calls its `train()` /
`train_teacher_student()` on CPU.  Here:
  * CPU: the oracle's restated step (oracle/step_ref.py) and the product's step driver
    (bdbnn_b200.step.TrainStep with CPU loss ops) must reproduce those fixtures; where the reference is
    mounted (build container) the loop is additionally re-executed live, and its import block is
    resolved against this repo's drop-in packages.
  * GPU (tests/test_gpu_ref_train.py): the CUDA path replays the same fixtures."""
import os
import sys

import pytest
import torch

import ref_launcher as RL
import train_cases as TC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = [c for c in TC.CASES if c not in TC.SUMMARY_ONLY]


def load_golden(case):
    return torch.load(os.path.join(GOLD, f"train_step_{case}.pt"), weights_only=False)


def check_terms(out, gold, rtol=1e-6, atol=1e-7):
    """Loss terms / accuracies the reference loop fed to its AverageMeters (train.py:519-524, 638-646)."""
    m = gold["meters"]
    pairs = [("loss", "Loss"), ("ce", "Loss_ce"), ("kurt", "Loss_kurt"), ("kl", "Loss_kl"), ("kl_c", "Loss_kl_c"),
             ("acc1", "Acc@1"), ("acc5", "Acc@5")]
    for ours, theirs in pairs:
        if theirs in m:
            got = float(torch.as_tensor(out[ours]).reshape(-1)[0])
            assert got == pytest.approx(m[theirs][-1], rel=rtol, abs=atol), (ours, got, m[theirs][-1])
    assert gold["scalars"]["Train Loss"][-1] == pytest.approx(float(out["loss"]), rel=rtol, abs=atol)


def check_tensors(got, gold, tol, what):
    """Full tensors (small nets): max error relative to max|reference| per tensor."""
    assert got.keys() == gold.keys()
    for n, ref in gold.items():
        scale = ref.abs().max().item() + 1e-12
        err = (got[n].detach().cpu().float() - ref).abs().max().item()
        assert err <= tol * scale, (what, n, err, scale)


def check_summaries(got, gold, tol, what):
    """Per-parameter summaries (ResNet-18): norm, sum and 64 sampled entries."""
    assert got.keys() == gold.keys()
    for n, ref in gold.items():
        f = got[n].detach().double().reshape(-1).cpu()
        scale = ref["absmax"] + 1e-12
        assert abs(float(f.norm()) - ref["norm"]) <= tol * max(ref["norm"], 1e-12), (what, n, "norm")
        s = f[TC.sample_index(f.numel())].float()
        assert (s - ref["samples"]).abs().max().item() <= tol * scale, (what, n, "samples")


def oracle_step(case, gold):
    from oracle import step_ref as S
    c = TC.CASES[case]
    stud, teacher = TC.build_oracle(case)
    if "init" in gold:
        stud.load_state_dict(gold["init"])
        if teacher is not None:
            teacher.load_state_dict(gold["teacher_state"])
        x = gold["x"]
    else:
        assert TC.state_checksum(stud) == pytest.approx(gold["init_checksum"], rel=1e-12), "seeded init differs"
        x = TC.make_batch(case)[0]
        assert float(x.double().sum()) == pytest.approx(gold["x_checksum"], rel=1e-12)
    y = gold["y"]
    hooked = S.ref_hooked_weights(stud) if c["args"].get("w_kurtosis") else {}
    assert list(hooked) == gold["hooked"]
    kw = TC.step_kwargs(case, len(hooked))
    opt = S.ref_make_optimizer(stud, c["dataset"], c["args"]["lr"])
    out = S.ref_train_step(stud, opt, x, y, hooked=hooked, teacher=teacher, **kw)
    return stud, out


@pytest.mark.parametrize("case", list(TC.CASES))
def test_oracle_step_reproduces_reference_loop(case):
    """oracle/step_ref.py == the reference's train()/train_teacher_student() on the stored fixtures."""
    gold = load_golden(case)
    stud, out = oracle_step(case, gold)
    check_terms(out, gold)
    grads = {n: p.grad for n, p in stud.named_parameters()}
    after = dict(stud.named_parameters())
    if case in TC.SUMMARY_ONLY:
        check_summaries(grads, gold["grads"], 1e-5, "grad")
        check_summaries(after, gold["after"], 1e-6, "param")
    else:
        check_tensors(grads, gold["grads"], 1e-6, "grad")
        check_tensors(after, gold["after"], 1e-6, "param")
        check_tensors(dict(stud.named_buffers()), gold["buffers"], 1e-6, "buffer")


@pytest.mark.parametrize("case", SMALL)
def test_product_step_driver_reproduces_reference_loop_on_cpu(case):
    """bdbnn_b200.step.TrainStep (hook selection, loss assembly, optimizer construction) driven on CPU with
    the oracle's loss ops and modules: same fixtures, so the host logic of the product is pinned too."""
    from bdbnn_b200.step import StepConfig, TrainStep, make_optimizer
    from oracle.models_ref import RefOps
    gold = load_golden(case)
    c = TC.CASES[case]
    a = c["args"]
    stud, teacher = TC.build_oracle(case)
    stud.load_state_dict(gold["init"])
    if teacher is not None:
        teacher.load_state_dict(gold["teacher_state"])
    cfg = StepConfig(w_kurtosis=bool(a.get("w_kurtosis")), w_kurtosis_target=a.get("w_kurtosis_target", 1.8),
                     w_lambda_kurtosis=a.get("w_lambda_kurtosis", 1.0), kurtosis_mode=a.get("kurtosis_mode", "avg"),
                     teacher_student=c["teacher"], react=bool(a.get("react")), alpha=a.get("alpha", 0.9),
                     beta=a.get("beta", 200.0))
    step = TrainStep(stud, make_optimizer(stud, c["dataset"], lr=a["lr"]), cfg, teacher=teacher, ops=RefOps)
    assert list(step.hooked) == gold["hooked"]
    out = step(gold["x"], gold["y"])
    check_terms(out, gold)
    check_tensors({n: p.grad for n, p in stud.named_parameters()}, gold["grads"], 1e-6, "grad")
    check_tensors(dict(stud.named_parameters()), gold["after"], 1e-6, "param")


needs_ref = pytest.mark.skipif(not RL.reference_available(), reason="reference tree not mounted (GPU box)")


@needs_ref
def test_reference_import_block_resolves_against_the_dropin_packages():
    """train.py:26-34 executed for real: `from utils import utils`, `from models import ...`,
    `from kurtosis import KurtosisWeight, RidgeRegularization, WeightRegularization`, `from utils import KD_loss`
    all bind to THIS repo's files when the repo root precedes the reference on sys.path."""
    train = RL.import_reference_train("dropin")
    src = train.__bdbnn_sources__
    for mod in ("kurtosis", "utils.utils", "utils.KD_loss", "models"):
        assert src[mod].startswith(RL.REPO_ROOT), (mod, src[mod])
    import bdbnn_b200
    assert train.KurtosisWeight is bdbnn_b200.KurtosisWeight
    assert train.HardBinaryConv is bdbnn_b200.HardBinaryConv
    assert train.KD_loss.DistributionLoss is bdbnn_b200.DistributionLoss
    assert train.RidgeRegularization and train.WeightRegularization
    for fn in ("cpt_tk", "accuracy", "AverageMeter", "ProgressMeter", "save_checkpoint", "find_weight_tensor_by_name"):
        assert hasattr(train.utils, fn)
    assert "resnet18" in train.IMAGENET_MODEL_NAMES and "resnet20" in train.CIFAR10_MODEL_NAMES   # train.py:50-56
    args = RL.make_args(train)
    assert args.w_l2_reg is False and args.w_lambda_ce == 1.0 and args.beta == 200 and args.alpha == 0.9


@needs_ref
@pytest.mark.parametrize("case", ["cifar_kurt", "cifar_ts"])
def test_reference_loop_live_equals_fixture(case, tmp_path):
    """Re-run the reference loop now (not from the stored file) and compare with the oracle step: guards the
    fixtures against going stale."""
    import subprocess
    gen = os.path.join(GOLD, "make_train_golden.py")
    before = load_golden(case)
    r = subprocess.run([sys.executable, gen, case], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, BDBNN_GOLDEN_OUT=str(tmp_path)))
    assert r.returncode == 0, r.stderr[-2000:]
    after = torch.load(os.path.join(tmp_path, f"train_step_{case}.pt"), weights_only=False)
    assert after["meters"] == before["meters"]
    for n in before["grads"]:
        assert torch.equal(before["grads"][n], after["grads"][n]), n
    assert after["sources"]["kurtosis"].startswith(RL.REF_ROOT)          # the reference's own kurtosis.py ran


def test_dropin_utils_match_reference_semantics():
    """utils/utils.py helpers the loop calls: meters' running average and rendering, accuracy, parameter lookup."""
    from utils import utils as U
    m = U.AverageMeter("Loss", ":.4e")
    m.update(2.0, 4)
    m.update(1.0, 12)
    assert m.avg == pytest.approx(1.25) and m.get_avg() == m.avg and m.count == 16 and m.val == 1.0
    assert str(m) == "Loss 1.0000e+00 (1.2500e+00)"
    lines = []

    class Lg:
        def info(self, s):
            lines.append(s)
    U.ProgressMeter(391, [m], Lg(), prefix="Epoch: [3]").display(7)
    assert lines == ["Epoch: [3][  7/391]\tLoss 1.0000e+00 (1.2500e+00)"]
    out = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    acc1, acc2 = U.accuracy(out, torch.tensor([1, 2, 2]), topk=(1, 2))
    assert float(acc1) == pytest.approx(200.0 / 3) and float(acc2) == pytest.approx(200.0 / 3)
    lin = torch.nn.Linear(2, 2)
    assert U.find_weight_tensor_by_name(lin, "bias") is lin.bias and U.find_weight_tensor_by_name(lin, "nope") is None
    from kurtosis import RidgeRegularization, WeightRegularization
    w = torch.tensor([[0.5, -2.0], [1.0, 0.0]])
    r = RidgeRegularization(w, "w")
    assert r.l2_regularization() is None and float(r.l2_loss) == pytest.approx(5.25)
    q = WeightRegularization(w, "w")
    assert q.w_regularization() is None and q.size == 4
    assert float(q.wr_loss) == pytest.approx((0.25 + 1.0 + 0.0 + 1.0) ** 0.5)

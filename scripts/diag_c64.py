#!/usr/bin/env python
"""Diagnostic (GPU): layer1.0 of ResNet-18 fed the oracle's input — BDBNN_TC_C64=1 vs 0, element-wise."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
if len(sys.argv) > 1:
    import train_cases as TC
    from test_gpu_ref_train import _oracle_blockwise
    os.environ["BDBNN_GRAD_MODE"] = "bf16x2"
    ref, x, y, rec, _ = _oracle_blockwise("r18_ce")
    prod, _ = TC.build_product("r18_ce"); prod.load_state_dict(ref.state_dict())
    prod = prod.cuda().to(memory_format=torch.channels_last).train()
    blk = dict(prod.named_modules())["layer1.0"]
    r = rec["layer1.0"]
    xin = r["in"].detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = blk(xin); out.backward(r["gout"].detach().cuda().contiguous(memory_format=torch.channels_last))
    torch.save({"out": out.detach().cpu(), "gin": xin.grad.cpu(), **{n: p.grad.cpu() for n, p in blk.named_parameters()}}, sys.argv[1])
    sys.exit(0)
res = {}
for v in ("0", "1"):
    f = f"/tmp/c64_{v}.pt"
    subprocess.run([sys.executable, __file__, f], env=dict(os.environ, BDBNN_TC_C64=v), check=True)
    res[v] = torch.load(f)
for k in res["0"]:
    a, b = res["1"][k].double(), res["0"][k].double()
    d = (a - b).abs()
    print(k, "max", float(d.max()), "ref max", float(b.abs().max()))
    if d.max() > 1e-4 * b.abs().max():
        if d.dim() == 4 and k in ("out", "gin"):
            per_c = d.amax(dim=(0, 2, 3)); print("   per-channel max err: top", torch.topk(per_c, 4))
            cbad = int(per_c.argmax()); m = d[:, cbad] > 0.1 * d[:, cbad].max()
            idx = m.nonzero()[:12]; print("   positions (n,h,w) of the worst channel:", idx.tolist())
        elif d.dim() == 4:
            print("   per-out-channel", torch.topk(d.amax(dim=(1, 2, 3)), 3), " per-in-channel", torch.topk(d.amax(dim=(0, 2, 3)), 3))
        else:
            print("   top", torch.topk(d, 3))

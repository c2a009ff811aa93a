#!/usr/bin/env python
"""Diagnostic (GPU): where does the ResNet-18 224x224 product forward diverge from the oracle?
(a) free-running: per-stage output error and fraction of sign mismatches; (b) teacher-forced: each block is fed
the ORACLE's input of that block, so discrete sign flips upstream cannot propagate."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import train_cases as TC

case = "r18_ce"
ref, _ = TC.build_oracle(case)
x, y = TC.make_batch(case)
ref.train()
feats_ref = {}
def hook(name):
    def f(m, i, o):
        feats_ref[name] = (i[0].detach().clone(), o.detach().clone())
    return f
names = ["maxpool"] + [f"layer{l}.{b}" for l in (1, 2, 3, 4) for b in (0, 1)]
mods = dict(ref.named_modules())
hs = [mods[n].register_forward_hook(hook(n)) for n in names]
sd = {k: v.clone() for k, v in ref.state_dict().items()}
out_ref = ref(x)
for h in hs: h.remove()

def run(tag):
    prod, _ = TC.build_product(case)
    prod.load_state_dict(sd)
    prod = prod.cuda().to(memory_format=torch.channels_last).train()
    feats = {}
    pm = dict(prod.named_modules())
    def hk(name):
        def f(m, i, o):
            feats[name] = o.detach().float().cpu()
        return f
    hs = [pm[n].register_forward_hook(hk(n)) for n in names if n != "maxpool"]
    xd = x.cuda().contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        st = prod._stem(xd)
        feats["maxpool"] = st.float().cpu()
        o = prod(xd)
    for h in hs: h.remove()
    res = {"tag": tag, "logit_err": (o.cpu() - out_ref).abs().max().item(), "free": {}, "forced": {}}
    for n in names:
        r = feats_ref[n][1]
        g = feats[n]
        res["free"][n] = {"rel": ((g - r).abs().max() / r.abs().max()).item(),
                          "signflip": ((g >= 0) != (r >= 0)).float().mean().item()}
    # teacher-forced blocks
    for n in names[1:]:
        blk = pm[n]
        xin = feats_ref[n][0].cuda().contiguous(memory_format=torch.channels_last)
        # fresh BN buffers do not matter for train-mode outputs
        with torch.no_grad():
            g = blk(xin).float().cpu()
        r = feats_ref[n][1]
        d = (g - r).abs()
        res["forced"][n] = {"rel": (d.max() / r.abs().max()).item(), "mean_rel": (d.mean() / r.abs().mean()).item(),
                            "signflip": ((g >= 0) != (r >= 0)).float().mean().item()}
    print(json.dumps(res), flush=True)

envs = [("default", {}), ("stem_cudnn", {"BDBNN_STEM_TC": "0"}), ("no_shortcut_tc", {"BDBNN_SHORTCUT_TC": "0"}),
        ("no_fwd8", {"BDBNN_FWD8": "0"}), ("no_fuse", {"BDBNN_FUSE_BN": "0"}),
        ("all_off", {"BDBNN_STEM_TC": "0", "BDBNN_SHORTCUT_TC": "0", "BDBNN_FWD8": "0", "BDBNN_FUSE_BN": "0"})]
for tag, e in envs:
    for k in ("BDBNN_STEM_TC", "BDBNN_SHORTCUT_TC", "BDBNN_FWD8", "BDBNN_FUSE_BN"):
        os.environ.pop(k, None)
    os.environ.update(e)
    try:
        run(tag)
    except Exception as ex:
        print(json.dumps({"tag": tag, "error": repr(ex)[:500]}), flush=True)

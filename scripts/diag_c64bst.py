#!/usr/bin/env python
"""Diagnostic (GPU): BatchNorm backward sums in the pixel-N dgrad epilogue (BDBNN_TC_C64=2) — parity with the separate
reduction pass on 64-channel units, then the ResNet-18 step with C64=2 vs 1."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "parity":
    import torch
    from bdbnn_b200 import _lib
    from bdbnn_b200.functional import conv_bn_add
    for mode in ("fp16s", "bf16x2"):
        os.environ["BDBNN_GRAD_MODE"] = mode
        for geom in ((2, 64, 40, 40), (3, 64, 56, 56), (2, 64, 17, 23)):
            n, c, h, w = geom
            g = torch.Generator().manual_seed(7 + sum(geom))
            x0 = (torch.randn(n, c, h, w, generator=g) * 1.1).cuda().contiguous(memory_format=torch.channels_last)
            ws = [(torch.randn(c, c, 3, 3, generator=g) * 0.6).cuda() for _ in range(2)]
            gam = [(torch.rand(c, generator=g) + 0.5).cuda() for _ in range(2)]
            bet = [(torch.randn(c, generator=g) * 0.2).cuda() for _ in range(2)]
            gz = torch.randn(n, c, h, w, generator=g).cuda().contiguous(memory_format=torch.channels_last)
            res, launches = {}, {}
            for flag in ("1", "0"):
                os.environ["BDBNN_BWD_STATS"] = flag
                x = x0.clone().requires_grad_(True)
                wp = [t.clone().requires_grad_(True) for t in ws]
                gp = [t.clone().requires_grad_(True) for t in gam]
                bp = [t.clone().requires_grad_(True) for t in bet]
                rm = [torch.zeros(c, device="cuda") for _ in range(2)]
                rv = [torch.ones(c, device="cuda") for _ in range(2)]
                z1 = conv_bn_add(x, wp[0], gp[0], bp[0], x, rm[0], rv[0], 0.1, 1e-5, 1, 1)
                z2 = conv_bn_add(z1, wp[1], gp[1], bp[1], z1, rm[1], rv[1], 0.1, 1e-5, 1, 1)
                n0 = _lib.launch_count()
                z2.backward(gz)
                torch.cuda.synchronize()
                launches[flag] = _lib.launch_count() - n0
                res[flag] = [x.grad] + [t.grad for t in wp + gp + bp]
            worst = max(((a - b).abs().max().item() / (b.abs().max().item() + 1e-30)) for a, b in zip(res["1"], res["0"]))
            print(f"C64={os.environ.get('BDBNN_TC_C64')} {mode} {geom}: launches {launches}, worst rel diff {worst:.2e}")
    sys.exit(0)
for c64 in ("2", "1"):
    env = dict(os.environ, BDBNN_TC_C64=c64)
    if c64 == "2":
        subprocess.run([sys.executable, __file__, "parity"], env=env, check=False)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline",
                          "--no-eager-gpu", "--no-secondary"], env=env, capture_output=True, text=True)
    import json
    try:
        d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        ks = {k["kernel"]: k["ms_per_step"] for k in d["kernels"]}
        print(f"C64={c64}: {d['value']} img/s {d['ms_per_step']} ms  dgrad {ks.get('binconv_dgrad_tc')}  bn_bwd_pack {ks.get('bn_bwd_pack')}")
    except Exception as e:
        print("bench failed", e, out.stderr[-500:])

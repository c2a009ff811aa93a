#!/usr/bin/env python
"""Per-layer kernel micro-benchmark (ResNet-18 N=256 binary-conv shapes):
CUDA-event time, algorithmic GB/s and fraction of the measured HBM peak for every kernel of the path.
L2 is flushed (256 MB write) between timed launches.   python scripts/kernel_bench.py [--impl xnor|tc]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bdbnn_b200 import _lib  # noqa: E402
from bdbnn_b200.functional import _p, _stream, algorithmic_bytes, conv_shape, grad_mode  # noqa: E402

R18 = [("layer1", 64, 56, 64, 1), ("layer2.0.conv1", 64, 56, 128, 2), ("layer2", 128, 28, 128, 1),
       ("layer3.0.conv1", 128, 28, 256, 2), ("layer3", 256, 14, 256, 1),
       ("layer4.0.conv1", 256, 14, 512, 2), ("layer4", 512, 7, 512, 1)]


def timeit(fn, flush, iters=5):
    ts = []
    for _ in range(iters + 1):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[1:])
    return ts[len(ts) // 2]


def loss_kernels(a, flush, peak):
    """kurtosis (19 hooked ResNet-18 weights, kurtosis.py:23-39), KD logits ([256,1000], KD_loss.py:16-43) and
    KD layer (16 weight pairs, KD_loss.py:52-67) kernels: fwd and bwd launches timed separately."""
    if a.layers and "losses" not in a.layers.split(","):
        return []
    import ctypes as C
    L = _lib.lib()
    ck = _lib.check
    st = _stream()
    from bdbnn_b200.functional import _ptr_array
    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = []
    for cin, cout, n3, ds in ((64, 64, 4, False), (64, 128, 1, True), (128, 128, 3, False), (128, 256, 1, True),
                              (256, 256, 3, False), (256, 512, 1, True), (512, 512, 3, False)):
        shapes += [(cout, cin, 3, 3)] * n3
        if ds:
            shapes.append((cout, cin, 1, 1))
    ws = [torch.randn(sh, device="cuda", generator=g) * 0.05 for sh in shapes]
    n = len(ws)
    tot = sum(w.numel() for w in ws)
    numel = (C.c_int64 * n)(*[w.numel() for w in ws])
    tg = (C.c_float * n)(*([1.8] * n))
    moments = torch.empty(n * 8, dtype=torch.float64, device="cuda")
    kurt, loss = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    gout = torch.ones(n, device="cuda")
    grads = [torch.empty_like(w) for w in ws]
    wp, gp = _ptr_array(ws), _ptr_array(grads)
    rows = []

    def add(name, fn, nbytes, note):
        fn(); torch.cuda.synchronize()
        ms = timeit(fn, flush, iters=9)
        gbs = nbytes / 1e9 / (ms / 1e3)
        rows.append({"layer": "losses", "kernel": name, "ms": round(ms, 4), "alg_MB": round(nbytes / 1e6, 2),
                     "GBs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4), "TFLOPs": None, "note": note})
        print(rows[-1], flush=True)

    add("kurtosis_multi_fwd", lambda: ck(L.bdbnn_kurtosis_multi_fwd(wp, numel, tg, n, _p(moments), _p(kurt), _p(loss), st), "k"),
        4 * tot, f"{n} tensors, {tot} weights: moments + finalize (2 launches)")
    add("kurtosis_multi_bwd", lambda: ck(L.bdbnn_kurtosis_multi_bwd(wp, numel, tg, n, _p(moments), _p(gout), gp, 0, st), "k"),
        8 * tot, "closed-form gradient, read W + write grad")
    N, Cc = a.batch, 1000
    s_, t_ = torch.randn(N, Cc, device="cuda", generator=g), torch.randn(N, Cc, device="cuda", generator=g)
    row, l0, gr = torch.empty(N, device="cuda"), torch.empty((), device="cuda"), torch.empty(N, Cc, device="cuda")
    add("kd_logits_fwd_bwd", lambda: ck(L.bdbnn_kd_logits_fwd_bwd(_p(s_), _p(t_), N, Cc, _p(row), _p(l0), _p(gr), st), "kd"),
        12 * N * Cc, f"[{N},{Cc}] student/teacher logits: loss + gradient (2 launches); latency-bound")
    w3 = [w for w in ws if w.shape[-1] == 3]
    wt3 = [torch.randn_like(w) * 0.05 for w in w3]
    n3 = len(w3)
    tot3 = sum(w.numel() for w in w3)
    numel3 = (C.c_int64 * n3)(*[w.numel() for w in w3])
    partial = torch.empty(n3, dtype=torch.float64, device="cuda")
    g3 = [torch.empty_like(w) for w in w3]
    one = torch.ones(1, device="cuda")
    sp, tp, g3p = _ptr_array(w3), _ptr_array(wt3), _ptr_array(g3)
    add("kd_layer_multi_fwd", lambda: ck(L.bdbnn_kd_layer_multi_fwd(sp, tp, numel3, n3, _p(partial), _p(l0), st), "kl"),
        8 * tot3, f"{n3} weight pairs, {tot3} weights")
    add("kd_layer_multi_bwd", lambda: ck(L.bdbnn_kd_layer_multi_bwd(tp, numel3, n3, _p(one), g3p, 0, st), "kl"),
        8 * tot3, "read Wt, write grad")
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--impl", default="xnor")
    ap.add_argument("--out", default=None)
    ap.add_argument("--layers", default=None, help="comma list of layer names to run")
    ap.add_argument("--kernels", default=None, help="comma list of kernel names to time")
    a = ap.parse_args()
    L = _lib.lib()
    peak = 6575.1
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    flush = torch.empty(64 * 1024 * 1024, device="cuda")
    rows = []
    for name, cin, hw, cout, stride in R18:
        if a.layers and name not in a.layers.split(","):
            continue
        n = a.batch
        sh = conv_shape((n, cin, hw, hw), (cout, cin, 3, 3), stride, 1)
        cw = (cin + 31) // 32
        x = torch.randn(n, hw, hw, cin, device="cuda")
        w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
        sb = torch.empty((n, hw, hw, cw), dtype=torch.int32, device="cuda")
        mb = torch.empty_like(sb)
        xb = torch.empty((n, hw, hw, cin), dtype=torch.int16, device="cuda")
        gname, GC, GH, FMT = grad_mode()
        alpha = torch.empty(cout, device="cuda")
        ws = torch.empty((cout, 9, cw), dtype=torch.int32, device="cuda")
        wm = torch.empty(((w.numel() + 31) // 32,), dtype=torch.int32, device="cuda")
        wf = torch.empty((cout, 9, cin), dtype=torch.int16, device="cuda")
        wt = torch.empty((cin, 9, cout), dtype=torch.int16, device="cuda")
        amax = torch.zeros(1, dtype=torch.int32, device="cuda")
        wf8 = torch.empty((cout, 9, cin), dtype=torch.uint8, device="cuda")
        xb8 = torch.empty((n, hw, hw, cin), dtype=torch.uint8, device="cuda")
        gs, igs = torch.empty(cout, device="cuda"), torch.empty(cout, device="cuda")
        y = torch.empty((n, sh.Ho, sh.Wo, cout), device="cuda")
        gy = torch.randn_like(y)
        gys = torch.empty(y.shape[:3] + (GH * cout,), dtype=torch.int16, device="cuda")
        gx = torch.empty_like(x)
        gw = torch.empty_like(w)
        st = _stream()
        shp = ctypes.byref(sh)
        caps = int(L.bdbnn_tc_supported(shp)) if a.impl == "tc" else 0
        tc = bool(caps & 1)
        ck = _lib.check
        kernels = {
            "act_pack" + ("_tc" if tc else ""): lambda: ck(L.bdbnn_act_pack(_p(x), n * hw * hw, cin, _p(sb), _p(mb), _p(xb if tc else None), FMT, st), "p"),
            "weight_pack": lambda: ck(L.bdbnn_weight_pack(_p(w), cout, cin, 3, 3, _p(alpha), _p(ws), _p(wm), _p(wf), _p(wt), _p(wf8), _p(gs), _p(igs), FMT, st), "w"),
        }
        if tc:
            nb = int(L.bdbnn_wgrad_tc_workspace_bytes(shp))
            wsb = torch.empty(max(nb, 4) // 4, device="cuda")
            kernels.update({
                "fwd_tc": lambda: ck(L.bdbnn_binconv_fwd_tc(_p(xb), _p(wf), FMT, _p(alpha), _p(y), shp, None, None, st), "f"),
                "grad_pack": lambda: ck(L.bdbnn_grad_pack(_p(gy), _p(gs), n * sh.Ho * sh.Wo, cout, GC, _p(amax), _p(gys), st), "g"),
                "dgrad_tc": lambda: ck(L.bdbnn_binconv_dgrad_tc(_p(gys), GC, _p(amax), _p(wt), _p(mb), _p(None), _p(gx), shp, st), "d"),
            })
            if caps & 8:         # fp8 forward only where the kernel takes the shape (never time a no-op)
                kernels["fwd_tc8"] = lambda: ck(L.bdbnn_binconv_fwd_tc8(_p(xb8), _p(wf8), _p(alpha), _p(y), shp, None, None, st), "f8")
            if caps & 4:
                kernels["wgrad_tc"] = lambda: ck(L.bdbnn_binconv_wgrad_tc(_p(gys), GC, _p(amax), _p(xb), _p(wm), _p(igs), _p(gw), shp, _p(wsb), nb, st), "w")
            else:
                kernels["wgrad"] = lambda: ck(L.bdbnn_binconv_wgrad(_p(gy), _p(sb), _p(wm), _p(gw), shp, st), "w")
        else:
            kernels.update({
                "fwd_xnor": lambda: ck(L.bdbnn_binconv_fwd_xnor(_p(sb), _p(ws), _p(alpha), _p(y), shp, st), "f"),
                "dgrad": lambda: ck(L.bdbnn_binconv_dgrad(_p(gy), _p(ws), _p(alpha), _p(mb), _p(gx), shp, st), "d"),
                "wgrad": lambda: ck(L.bdbnn_binconv_wgrad(_p(gy), _p(sb), _p(wm), _p(gw), shp, st), "w"),
            })
        if tc:
            ck(L.bdbnn_act_pack(_p(x), n * hw * hw, cin, _p(sb), _p(mb), _p(xb), FMT, st), "p")
            ck(L.bdbnn_bits_to_fp8(_p(sb), n * hw * hw, cin, _p(xb8), st), "b8")
        for kname, fn in kernels.items():
            if a.kernels and kname not in a.kernels.split(",") and "pack" not in kname:
                continue
            fn()
            torch.cuda.synchronize()
            ms = timeit(fn, flush, iters=3 if kname in ("dgrad", "wgrad") else 7)
            nbytes = algorithmic_bytes(kname, sh, GH) if kname != "weight_pack" else 4 * w.numel()
            gbs = nbytes / 1e9 / (ms / 1e3)
            macs = 2.0 * n * sh.Ho * sh.Wo * cout * cin * 9
            rows.append({"layer": name, "kernel": kname, "ms": round(ms, 4), "alg_MB": round(nbytes / 1e6, 2),
                         "GBs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4),
                         "TFLOPs": round(macs / 1e12 / (ms / 1e3), 1) if "pack" not in kname else None})
            print(rows[-1], flush=True)
        del x, y, gy, gx, xb, gys
        torch.cuda.empty_cache()
    rows += loss_kernels(a, flush, peak)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Decode the clock64 trace of CTA 0 of tc_conv2_kernel for one layer (developer aid)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bdbnn_b200 import _lib
from bdbnn_b200.functional import _p, _stream, conv_shape

cin, hw, cout, which = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
n = 256
L = _lib.lib()
sh = conv_shape((n, cin, hw, hw), (cout, cin, 3, 3), 1, 1)
x = torch.randn(n, hw, hw, cin, device="cuda"); w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
cw = cin // 32
sb = torch.empty((n, hw, hw, cw), dtype=torch.int32, device="cuda"); mb = torch.empty_like(sb)
xb = torch.empty((n, hw, hw, cin), dtype=torch.bfloat16, device="cuda")
alpha = torch.empty(cout, device="cuda"); ws = torch.empty((cout, 9, cw), dtype=torch.int32, device="cuda")
wm = torch.empty(((w.numel() + 31) // 32,), dtype=torch.int32, device="cuda")
wf = torch.empty((cout, 9, cin), dtype=torch.bfloat16, device="cuda"); wt = torch.empty((cin, 9, cout), dtype=torch.bfloat16, device="cuda")
gs = torch.empty(cout, device="cuda"); igs = torch.empty(cout, device="cuda")
y = torch.empty((n, hw, hw, cout), device="cuda"); gy = torch.randn_like(y)
gys = torch.empty((n, hw, hw, 2 * cout), dtype=torch.bfloat16, device="cuda"); gx = torch.empty_like(x)
st = _stream(); shp = ctypes.byref(sh)
L.bdbnn_act_pack(_p(x), n * hw * hw, cin, _p(sb), _p(mb), _p(xb), 1, st)
L.bdbnn_weight_pack(_p(w), cout, cin, 3, 3, _p(alpha), _p(ws), _p(wm), _p(wf), _p(wt), _p(None), _p(gs), _p(igs), 1, st)
L.bdbnn_grad_pack(_p(gy), _p(gs), n * hw * hw, cout, 2, _p(None), _p(gys), st)
nb = int(L.bdbnn_wgrad_tc_workspace_bytes(shp)); wsb = torch.empty(max(nb, 4) // 4, device="cuda"); gw = torch.empty_like(w)
wf8 = torch.empty((cout, 9, cin), dtype=torch.uint8, device="cuda"); xb8 = torch.empty((n, hw, hw, cin), dtype=torch.uint8, device="cuda")
L.bdbnn_weight_pack(_p(w), cout, cin, 3, 3, _p(alpha), _p(ws), _p(wm), _p(wf), _p(wt), _p(wf8), _p(gs), _p(igs), 1, st)
L.bdbnn_bits_to_fp8(_p(sb), n * hw * hw, cin, _p(xb8), st)
if which == "fwd8":
    run = lambda: L.bdbnn_binconv_fwd_tc8(_p(xb8), _p(wf8), _p(alpha), _p(y), shp, None, None, st)
elif which == "wgrad":
    run = lambda: L.bdbnn_binconv_wgrad_tc(_p(gys), 2, _p(None), _p(xb), _p(wm), _p(igs), _p(gw), shp, _p(wsb), nb, st)
else:
  run = (lambda: L.bdbnn_binconv_fwd_tc(_p(xb), _p(wf), 1, _p(alpha), _p(y), shp, None, None, st)) if which == "fwd" else \
      (lambda: L.bdbnn_binconv_dgrad_tc(_p(gys), 2, _p(None), _p(wt), _p(mb), _p(None), _p(gx), shp, st))
run(); torch.cuda.synchronize()
tr = torch.zeros(3 * 2048, dtype=torch.int64, device="cuda")
L.bdbnn_debug_trace(_p(tr)); run(); torch.cuda.synchronize(); L.bdbnn_debug_trace(None)
t = tr.cpu().view(3, 1024, 2)
t0 = min(int(t[r, 0, 1]) for r in range(3) if int(t[r, 0, 1]) > 0)
names = {0: {0: "P.wait_pempty", 1: "P.got_pempty", 2: "P.issued_kb"},
         1: {0: "M.wait", 1: "M.got", 2: "M.got_patch", 3: "M.got_B0", 4: "M.committed"},
         2: {0: "E.wait_tfull", 1: "E.got_tfull", 2: "E.done"}}
ev = []
for r in range(3):
    for i in range(1024):
        c = int(t[r, i, 1])
        if c == 0: break
        ev.append((c - t0, r, names[r][int(t[r, i, 0])]))
ev.sort()
for c, r, nm in ev[:int(sys.argv[5]) if len(sys.argv) > 5 else 90]:
    print(f"{c:9d}  {'  ' * r * 8}{nm}")

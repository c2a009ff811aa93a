#!/usr/bin/env python
"""Diagnostic (GPU): where the stem conv forward (pixel-N kernel, tc_conv64.cu STEM) spends its time.
BDBNN_TC_DBG bits: 1 = no stores, 2 = no epilogue work, 4 = no MMAs."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    from bdbnn_b200 import _lib
    from bdbnn_b200.functional import _p, _stream
    L = _lib.lib()
    n, h, w = 256, 224, 224
    dev = torch.device("cuda", 0)
    x = torch.randn(n, 3, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    wt = torch.randn(64, 3, 7, 7, device=dev) * 0.05
    st = _stream()
    xw = torch.empty((int(L.bdbnn_stem_xw_bytes(n, h, w)) // 2,), dtype=torch.int16, device=dev)
    x_amax = torch.empty((1,), dtype=torch.int32, device=dev)
    wf = torch.empty((64, 7, 32), dtype=torch.int16, device=dev)
    alpha = torch.empty((64,), dtype=torch.float32, device=dev)
    _lib.check(L.bdbnn_stem_pack(_p(x), n, h, w, x.stride(0), x.stride(1), x.stride(2), x.stride(3), _p(wt), _p(xw),
                                 _p(x_amax), _p(wf), _p(alpha), st), "stem_pack")
    y = torch.empty((n, 64, 112, 112), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
    sums = torch.empty((128,), dtype=torch.float64, device=dev)
    ymax = torch.empty((64,), dtype=torch.int32, device=dev)
    def run():
        _lib.check(L.bdbnn_stem_conv_fwd(_p(xw), _p(wf), _p(alpha), _p(y), n, h, w, _p(sums), _p(ymax), st), "fwd")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record(); torch.cuda.synchronize()
    print(f"DBG={os.environ.get('BDBNN_TC_DBG', '0')} STEM64={os.environ.get('BDBNN_TC_C64_STEM', '1')} "
          f"{e0.elapsed_time(e1) / 20:.4f} ms (includes the 2 statistics memsets)")
    sys.exit(0)
for env in ({"BDBNN_TC_DBG": "0"}, {"BDBNN_TC_DBG": "1"}, {"BDBNN_TC_DBG": "2"}, {"BDBNN_TC_DBG": "4"},
            {"BDBNN_TC_DBG": "6"}, {"BDBNN_TC_C64_STEM": "0"}):
    subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, **env), check=False)

import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for C in (3, 4, 8):
    for cl in (True, False):
        x = torch.randn(256, C, 224, 224, device="cuda")
        w = torch.randn(64, C, 7, 7, device="cuda", requires_grad=True)
        if cl:
            x = x.contiguous(memory_format=torch.channels_last); 
            w = w.detach().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        def fwd():
            return F.conv2d(x, w, None, 2, 3)
        y = fwd(); gy = torch.randn_like(y)
        def both():
            y = F.conv2d(x, w, None, 2, 3); y.backward(gy); w.grad = None
        print(f"C={C} channels_last={cl} out_cl={y.is_contiguous(memory_format=torch.channels_last)} fwd {t(fwd):.3f} ms  fwd+wgrad {t(both):.3f} ms", flush=True)

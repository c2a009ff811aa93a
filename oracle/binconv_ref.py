"""ORACLE (test infrastructure, NOT product code) — CPU restatement of the binary conv2d spec.

PARITY UNPINNED: the reference snapshot does not contain the binary-conv source (`models/` is
imported at train.py:27-32 and utils/KD_loss.py:6-7 but was never committed), and
its tests hold no golden vectors for it.  This file therefore restates the *authored* spec of
DESIGN.md §2, whose constraints come from the reference call sites:
  * nn.Module with a 4-D parameter named `weight` [Cout,Cin,kh,kw]      (train.py:327,391; KD_loss.py:65)
  * forward(x[N,Cin,H,W] fp32) -> [N,Cout,Ho,Wo] fp32, autograd-differentiable (train.py:492,528)
  * 1W/1A: sign-quantised weights and activations with STE gradients     (BASELINE.json north_star)

Spec (Bi-Real-Net `HardBinaryConv` lineage, which the class names follow):
  sign(v)   := +1 if v >= 0 else -1                       (sign(0)=+1 so one bit encodes it)
  xb        = sign(x)                 d xb / d x  := 1[|x| <= 1]        (hard-tanh STE)
  alpha[o]  = mean_{c,r,s} |W[o,c,r,s]|   (treated as a constant in backward)
  Wb        = alpha[o] * sign(W)      d Wb / d W  := 1[|W| <= 1]        (clamp STE, no alpha factor)
  y         = conv2d(xb, Wb, stride, zero padding)        (padded taps contribute 0)

EDE variant (reference: `--ede`, train.py:409-415 assigns module.k / module.t from
utils/utils.py:8-14 `cpt_tk`; IR-Net's "error decay estimator" the flag is named after):
  d sign(v) / d v := k * t * (1 - tanh(t*v)^2)     for both x and W, replacing the two indicators.
The schedule (cpt_tk) IS pinned by the reference (tests/golden/cpt_tk_cases.pt); the derivative form is
authored like the rest of this file.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def sign_pm1(v: torch.Tensor) -> torch.Tensor:
    """+1 where v >= 0 else -1 (NaN -> -1), same dtype as v."""
    return torch.where(v >= 0, torch.ones_like(v), -torch.ones_like(v))


def ste_mask(v: torch.Tensor) -> torch.Tensor:
    """1 where |v| <= 1 else 0 (NaN -> 0)."""
    return (v.abs() <= 1).to(v.dtype)


def weight_alpha(weight: torch.Tensor) -> torch.Tensor:
    return weight.abs().mean(dim=(1, 2, 3))


def binconv_int(x: torch.Tensor, weight: torch.Tensor, stride: int, padding: int) -> torch.Tensor:
    """Pre-scale integer part: conv2d(sign(x), sign(W)); every value is an exact integer."""
    return F.conv2d(sign_pm1(x), sign_pm1(weight), None, stride, padding)


def binconv_forward(x, weight, stride=1, padding=1):
    """y = alpha[o] * conv2d(sign(x), sign(W)).  fp64 inputs give the infinitely-precise oracle."""
    alpha = weight_alpha(weight)
    return binconv_int(x, weight, stride, padding) * alpha.view(1, -1, 1, 1)


def binconv_backward(x, weight, gy, stride=1, padding=1):
    """Closed-form backward of the spec: returns (gx, gW)."""
    alpha = weight_alpha(weight)
    xb = sign_pm1(x)
    wb = sign_pm1(weight) * alpha.view(-1, 1, 1, 1)
    gx_b = torch.nn.grad.conv2d_input(x.shape, wb, gy, stride=stride, padding=padding)
    gw_b = torch.nn.grad.conv2d_weight(xb, weight.shape, gy, stride=stride, padding=padding)
    return gx_b * ste_mask(x), gw_b * ste_mask(weight)


def ede_factor(v, k, t):
    """k * t * (1 - tanh(t*v)^2): soft-sign derivative of the EDE backward."""
    k = torch.as_tensor(k, dtype=v.dtype).reshape(())
    t = torch.as_tensor(t, dtype=v.dtype).reshape(())
    return k * t * (1.0 - torch.tanh(t * v) ** 2)


def binconv_backward_ede(x, weight, gy, k, t, stride=1, padding=1):
    """Closed-form backward with the EDE derivative in place of both STE indicators."""
    alpha = weight_alpha(weight)
    xb = sign_pm1(x)
    wb = sign_pm1(weight) * alpha.view(-1, 1, 1, 1)
    gx_b = torch.nn.grad.conv2d_input(x.shape, wb, gy, stride=stride, padding=padding)
    gw_b = torch.nn.grad.conv2d_weight(xb, weight.shape, gy, stride=stride, padding=padding)
    return gx_b * ede_factor(x, k, t), gw_b * ede_factor(weight, k, t)


def cpt_tk(epoch, tot_epochs, t_min=1e-2, t_max=1e1):
    """(t, k) of the EDE schedule, restating utils/utils.py:8-14: t = 10^(lg Tmin + (lg Tmax - lg Tmin) *
    epoch / tot_epochs) in fp32, k = max(1/t, 1)."""
    lo, hi = torch.log10(torch.tensor(t_min).float()), torch.log10(torch.tensor(t_max).float())
    t = torch.tensor([torch.pow(torch.tensor(10.0), lo + (hi - lo) / tot_epochs * epoch)]).float()
    k = torch.maximum(1 / t, torch.tensor(1.0)).float()
    return t, k


class _SignEDE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, k, t):
        ctx.save_for_backward(v, k, t)
        return sign_pm1(v)

    @staticmethod
    def backward(ctx, g):
        v, k, t = ctx.saved_tensors
        return g * ede_factor(v, k, t), None, None


def binconv2d_ref_ede(x, weight, k, t, stride=1, padding=1):
    """Autograd version of the EDE variant."""
    alpha = weight_alpha(weight).detach()
    k = torch.as_tensor(k, dtype=x.dtype)
    t = torch.as_tensor(t, dtype=x.dtype)
    xb = _SignEDE.apply(x, k, t)
    # alpha multiplies the value only: route the weight gradient around it (no alpha factor, as in the STE spec)
    wsgn = _SignEDE.apply(weight, k, t)
    wb = (wsgn * alpha.view(-1, 1, 1, 1)).detach() - wsgn.detach() + wsgn
    return F.conv2d(xb, wb, None, stride, padding)


class _SignSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        ctx.save_for_backward(v)
        return sign_pm1(v)

    @staticmethod
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        return g * ste_mask(v)


def binconv2d_ref(x, weight, stride=1, padding=1):
    """Autograd version of the spec (same maths as binconv_forward / binconv_backward)."""
    alpha = weight_alpha(weight).detach()
    xb = _SignSTE.apply(x)
    # Bi-Real HardBinaryConv: value alpha*sign(W), gradient of clamp(W,-1,1) (no alpha factor)
    clipped = torch.clamp(weight, -1.0, 1.0)
    wb = (sign_pm1(weight) * alpha.view(-1, 1, 1, 1)).detach() - clipped.detach() + clipped
    return F.conv2d(xb, wb, None, stride, padding)


class RefBinarizeConv2d(nn.Conv2d):
    """Pure-PyTorch fp32 module with the product module's surface (bdbnn_b200.modules.BinarizeConv2d)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False, **kw):
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias=False)
        self.k = None      # assigned by the training loop under --ede (train.py:412-415)
        self.t = None
        self.supports_ede = False

    def forward(self, x):
        if self.supports_ede and self.k is not None and self.t is not None:
            return binconv2d_ref_ede(x, self.weight, self.k, self.t, self.stride[0], self.padding[0])
        return binconv2d_ref(x, self.weight, self.stride[0], self.padding[0])


class RefBinarizeConv2dCifar(RefBinarizeConv2d):
    """Oracle twin of HardBinaryConv_cifar: honours an assigned (.k, .t) pair (EDE backward)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.supports_ede = True


# ---- bit-level restatement of what the packed kernels compute (numpy-free, small sizes only) --------
def pack_bits_nhwc(x: torch.Tensor) -> torch.Tensor:
    """x [N,C,H,W] -> int64 words [N,H,W,Cw]; bit j of word k = (x[n,32k+j,h,w] >= 0)."""
    n, c, h, w = x.shape
    cw = (c + 31) // 32
    bits = (x >= 0).permute(0, 2, 3, 1).to(torch.int64)
    pad = cw * 32 - c
    if pad:
        bits = F.pad(bits, (0, pad))
    bits = bits.view(n, h, w, cw, 32)
    weights = (1 << torch.arange(32, dtype=torch.int64))
    return (bits * weights).sum(-1)


def pack_mask_nhwc(x: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x.shape
    cw = (c + 31) // 32
    bits = (x.abs() <= 1).permute(0, 2, 3, 1).to(torch.int64)
    pad = cw * 32 - c
    if pad:
        bits = F.pad(bits, (0, pad))
    bits = bits.view(n, h, w, cw, 32)
    weights = (1 << torch.arange(32, dtype=torch.int64))
    return (bits * weights).sum(-1)


def pack_weight_bits(weight: torch.Tensor) -> torch.Tensor:
    """W [O,C,kh,kw] -> int64 words [O, kh*kw, Cw]; bit j of word k = (W[o,32k+j,t] >= 0)."""
    o, c, kh, kw = weight.shape
    cw = (c + 31) // 32
    bits = (weight >= 0).reshape(o, c, kh * kw).permute(0, 2, 1).to(torch.int64)
    pad = cw * 32 - c
    if pad:
        bits = F.pad(bits, (0, pad))
    bits = bits.view(o, kh * kw, cw, 32)
    weights = (1 << torch.arange(32, dtype=torch.int64))
    return (bits * weights).sum(-1)


def pack_flat_mask(weight: torch.Tensor) -> torch.Tensor:
    flat = (weight.abs() <= 1).reshape(-1).to(torch.int64)
    pad = (-flat.numel()) % 32
    if pad:
        flat = F.pad(flat, (0, pad))
    weights = (1 << torch.arange(32, dtype=torch.int64))
    return (flat.view(-1, 32) * weights).sum(-1)


def _popc(v: torch.Tensor) -> torch.Tensor:
    out = torch.zeros_like(v)
    for i in range(32):
        out += (v >> i) & 1
    return out


def xnor_popcount_conv(xbits, wbits, cin, h, w, kh, kw, stride, pad):
    """Integer conv from packed words exactly as the bit-serial kernel does it:
    out = sum over valid taps (Cin - 2*popc(x ^ w)).  Returns int64 [N,Ho,Wo,O]. Pure loops: tiny only."""
    n = xbits.shape[0]
    o = wbits.shape[0]
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    out = torch.zeros(n, ho, wo, o, dtype=torch.int64)
    for r in range(kh):
        for s in range(kw):
            t = r * kw + s
            for i in range(ho):
                hh = i * stride + r - pad
                if hh < 0 or hh >= h:
                    continue
                for j in range(wo):
                    ww = j * stride + s - pad
                    if ww < 0 or ww >= w:
                        continue
                    xw = xbits[:, hh, ww, :]                      # [N,Cw]
                    x_xor = xw[:, None, :] ^ wbits[None, :, t, :]  # [N,O,Cw]
                    out[:, i, j, :] += cin - 2 * _popc(x_xor).sum(-1)
    return out

"""Where does the tensor-core conv's error come from?  (GPU probe, not a test.)

For a few layer shapes of the path it compares, against an fp64 convolution of
the same fp32 inputs:
  * this library's 3xTF32 conv as shipped,
  * the same with the weights pre-rounded to tf32 (weight lo == 0),
  * the same with weights AND activations pre-rounded to tf32 (every product is
    exact in 22 bits: what is left is the tensor core's own accumulation),
  * torch/cuDNN fp32 (allow_tf32=False) and torch/cuDNN TF32 (the reference's
    default on an Ampere+ GPU).
Reported per variant: max|e|/max|ref|, rms(e)/rms(ref), and the slope of e on
ref (a multiplicative bias: what a truncating accumulator produces).

Then a handful of crafted 1x1 convs that show how the accumulator rounds.

    python tools/accuracy_probe.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from preworld_b200 import ops

DEV = 'cuda'


def rna_tf32(t):
    b = t.contiguous().view(torch.int32)
    return ((b + 0x1000) & -8192).view(torch.float32)


def stats(got, ref):
    e = (got.double() - ref).flatten()
    r = ref.flatten()
    slope = float((e * r).sum() / (r * r).sum())
    return {'max': float(e.abs().max() / r.abs().max()),
            'rms': float(e.pow(2).mean().sqrt() / r.pow(2).mean().sqrt()),
            'slope': slope}


SHAPES = [
    # dims, cin, cout, spatial, k
    (2, 32, 32, (64, 96), 1),
    (2, 64, 256, (64, 96), 1),
    (2, 1024, 256, (16, 44), 1),
    (3, 32, 32, (8, 48, 48), 3),
    (3, 64, 64, (8, 48, 48), 3),
    (3, 128, 128, (4, 50, 50), 3),
    (2, 256, 256, (16, 44), 3),
    (2, 512, 512, (8, 22), 3),
]


def run_shape(dims, cin, cout, sp, k, relu_in=True):
    g = torch.Generator().manual_seed(cin * 131 + cout * 7 + k)
    x = torch.randn(1, cin, *sp, generator=g)
    if relu_in:
        x = F.relu(x)
    w = torch.randn(cout, cin, *([k] * dims), generator=g) / (cin * k ** dims) ** .5
    x, w = x.to(DEV), w.to(DEV)
    conv = F.conv2d if dims == 2 else F.conv3d
    pad = k // 2
    perm = (0, *range(2, 2 + dims), 1)

    def ours(xx, ww):
        pc = ops.PackedConv(ww, None, None, stride=1, padding=pad)
        y = ops.conv(xx.permute(*perm).contiguous(), pc)
        return ops.to_logical(y)

    out = {}
    ref = conv(x.double(), w.double(), None, 1, pad)
    out['ours'] = stats(ours(x, w), ref)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out['cudnn_fp32'] = stats(conv(x, w, None, 1, pad), ref)
    torch.backends.cudnn.allow_tf32 = True
    out['cudnn_tf32'] = stats(conv(x, w, None, 1, pad), ref)
    torch.backends.cudnn.allow_tf32 = False
    wr = rna_tf32(w)
    ref_w = conv(x.double(), wr.double(), None, 1, pad)
    out['ours_w_tf32'] = stats(ours(x, wr), ref_w)
    xr = rna_tf32(x)
    ref_xw = conv(xr.double(), wr.double(), None, 1, pad)
    out['ours_xw_tf32'] = stats(ours(xr, wr), ref_xw)
    out['cudnn_fp32_xw_tf32'] = stats(conv(xr, wr, None, 1, pad), ref_xw)
    return out


def crafted():
    """1x1 conv, cin 32, cout 32, all weights 1: y = sum_k x[k].  The 32 inputs
    go through four K=8 MMAs accumulating in TMEM."""
    res = {}
    w = torch.ones(32, 32, device=DEV)
    pc = ops.PackedConv(w, None, None)

    def run(vec):
        x = torch.tensor(vec, dtype=torch.float32, device=DEV).view(1, 1, 1, 32) \
            .expand(1, 8, 16, 32).contiguous()
        y = ops.conv(x, pc)
        v = y[0, 0, 0, 0].item()
        return v, float(torch.tensor(vec, dtype=torch.float64).sum())

    for name, e in (('2^-24', 2.0 ** -24), ('2^-25', 2.0 ** -25), ('2^-26', 2.0 ** -26),
                    ('2^-28', 2.0 ** -28), ('-2^-24', -2.0 ** -24),
                    ('-2^-25', -2.0 ** -25), ('1.5*2^-24', 1.5 * 2.0 ** -24)):
        for lead in (1.0, -1.0):
            got, exact = run([lead] + [e] * 31)
            res[f'lead {lead:+.0f} + 31 x {name}'] = {
                'got_minus_lead_ulps': (got - lead) / 2.0 ** -23,
                'exact_minus_lead_ulps': (exact - lead) / 2.0 ** -23}
        # the small terms first, the big one in the LAST k-step
        got, exact = run([e] * 31 + [1.0])
        res[f'31 x {name} then 1'] = {'got_minus_lead_ulps': (got - 1.0) / 2.0 ** -23,
                                      'exact_minus_lead_ulps': (exact - 1.0) / 2.0 ** -23}
    return res


def main():
    out = {'shapes': {}, 'crafted': crafted()}
    for s in SHAPES:
        key = f'{s[1]}->{s[2]} k{s[4]} {"x".join(map(str, s[3]))}'
        out['shapes'][key] = run_shape(*s)
        print(key)
        for name, st in out['shapes'][key].items():
            print(f'   {name:20s} max {st["max"]:.2e}  rms {st["rms"]:.2e}  slope {st["slope"]:+.2e}')
    for k, v in out['crafted'].items():
        print(f'{k:32s} got {v["got_minus_lead_ulps"]:+.3f} ulp   exact {v["exact_minus_lead_ulps"]:+.3f} ulp')
    if len(sys.argv) > 1:
        with open(sys.argv[1], 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()

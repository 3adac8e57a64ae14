"""Swin-B @ 512x1408 (the shipped image side, bevstereo-occ.py:45-74) on one GPU: time of the
backbone + neck for one frame's 6 cameras and the per-operator split of one block per stage
(CUDA events on the launching stream, L2 flushed between timed launches).

    python tools/swin_probe.py [--images 6] [--hw 512 1408] [--iters 5]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from preworld_b200 import _lib, configs, ops, plugin      # noqa: E402
from preworld_b200 import synthetic as S                  # noqa: E402
from preworld_b200.plugin.swin import _ln                 # noqa: E402


def timed(fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(iters)]
    fn()
    for a, b in ev:
        flush.add_(1.0)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=6)
    ap.add_argument('--hw', type=int, nargs=2, default=(512, 1408))
    ap.add_argument('--iters', type=int, default=5)
    a = ap.parse_args()
    cfg = configs.model_cfg('finetune', 'swin')
    bb = plugin.build_backbone(dict(cfg['img_backbone'], with_cp=False)).eval()
    neck = plugin.build_neck(cfg['img_neck']).eval()
    S.lively_init_(bb, 0)
    S.lively_init_(neck, 1)
    bb, neck = bb.cuda(), neck.cuda()
    x = torch.randn((a.images, 3, *a.hw), device='cuda')
    flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8).float() \
        if False else torch.zeros(64 << 20, device='cuda')          # 256 MB > L2
    res = {'images': a.images, 'hw': list(a.hw)}
    with torch.no_grad():
        n0 = _lib.launch_count()
        outs = bb(x)
        y = neck(outs[1:])
        res['launches'] = _lib.launch_count() - n0
        res['shapes'] = [list(o.shape) for o in outs] + [list(y.shape)]
        res['backbone_ms'] = timed(lambda: bb(x), a.iters, flush)
        res['neck_ms'] = timed(lambda: neck(outs[1:]), a.iters, flush)
        # per-operator split: first block of every stage
        t = bb.run_stem(x)
        p = bb.packs()
        ops_ms = []
        for i, st in enumerate(bb.stages):
            blk, bp = st.blocks[1], p['blocks'][i][1]          # the shifted block
            msa = blk.attn.w_msa
            n, h, w, c = t.shape
            ln = _ln(t, blk.norm1)
            qkv = ops.conv(ln, bp['qkv'])
            att = ops.window_attention(qkv, bp['qkv_bias'], bp['table'], msa.num_heads,
                                       blk.attn.window_size, blk.attn.shift_size, msa.scale)
            hid = ops.conv(ln, bp['fc1'], 'gelu')
            row = {'stage': i, 'tokens': n * h * w, 'c': c, 'heads': msa.num_heads,
                   'ln_us': 1e3 * timed(lambda: _ln(t, blk.norm1, out=ln), a.iters, flush),
                   'qkv_us': 1e3 * timed(lambda: ops.conv(ln, bp['qkv'], out=qkv), a.iters, flush),
                   'attn_us': 1e3 * timed(lambda: ops.window_attention(
                       qkv, bp['qkv_bias'], bp['table'], msa.num_heads, blk.attn.window_size,
                       blk.attn.shift_size, msa.scale, out=att), a.iters, flush),
                   'proj_us': 1e3 * timed(lambda: ops.conv(att, bp['proj'], residual=t, out=ln),
                                          a.iters, flush),
                   'fc1_gelu_us': 1e3 * timed(lambda: ops.conv(ln, bp['fc1'], 'gelu', out=hid),
                                              a.iters, flush),
                   'fc2_us': 1e3 * timed(lambda: ops.conv(hid, bp['fc2'], residual=t, out=ln),
                                         a.iters, flush),
                   'block_us': 1e3 * timed(lambda: blk.run(bp, t), a.iters, flush)}
            nw = -(-h // 12) * -(-w // 12) * n
            flops = nw * msa.num_heads * 2 * 2 * 144 * 144 * 32
            row['attn_tflops'] = flops / row['attn_us'] / 1e6
            row['qkv_tflops'] = 2 * n * h * w * c * 3 * c / row['qkv_us'] / 1e6
            row['fc1_tflops'] = 2 * n * h * w * c * 4 * c / row['fc1_gelu_us'] / 1e6
            ops_ms.append(row)
            del ln, qkv, att, hid
            t = bb.run_layer(i, t)
            if st.downsample is not None:
                t = type(st.downsample).run(p['down'][i], t)
        res['blocks'] = ops_ms
    print(json.dumps(res, indent=1))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/swin_probe.json', 'w'), indent=1)


if __name__ == '__main__':
    main()

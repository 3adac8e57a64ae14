"""Where the end-to-end step (public call, pinned host batch -> numpy result)
spends its host time: mean milliseconds between the trace points
`_graphed_occupancy` stamps when `model._e2e_trace` is a list.

    python tools/e2e_trace.py [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device('cuda', 0)
    wl = bench.Workload('finetune', dev, n_variants=4).to_device()
    model, pins = wl.model, wl.pin_samples
    model.enable_cuda_graph()

    def step(i):
        with torch.no_grad():
            return model(return_loss=False, img_inputs=[pins[i % len(pins)]],
                         img_metas=[None])

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    model._e2e_trace = []
    t0 = time.perf_counter()
    for i in range(steps):
        model._e2e_trace.append(('call', time.perf_counter()))
        step(i)
    wall = (time.perf_counter() - t0) / steps * 1e3
    tr = model._e2e_trace
    model._e2e_trace = None
    names, acc = [], {}
    for (a, ta), (b, tb) in zip(tr, tr[1:]):
        if b == 'call':
            continue
        k = f'{a} -> {b}'
        if k not in acc:
            names.append(k)
        acc[k] = acc.get(k, 0.0) + (tb - ta) * 1e3 / steps
    print(f'wall {wall:.3f} ms/step')
    for k in names:
        print(f'  {k:40s} {acc[k]:.3f} ms')


if __name__ == '__main__':
    main()

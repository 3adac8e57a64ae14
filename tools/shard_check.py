"""torchrun --nproc-per-node N tools/shard_check.py

Every rank builds the same model and sample, runs the forward once unsharded
and once camera-sharded over the N ranks (one NCCL all-gather of depth +
context features), and checks that the two occupancy grids are IDENTICAL;
then times both modes (device events, max over ranks)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
from preworld_b200.parallel import CameraShard


def main():
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)
    size = (64, 176) if '--tiny' in sys.argv else (256, 704)
    if '--tiny' in sys.argv:
        from oracle.cases import CASES, model_cfg_for
        from preworld_b200 import build_model
        from preworld_b200 import synthetic as S
        case = CASES['tiny_finetune']
        model = build_model(model_cfg_for(case)).eval()
        S.lively_init_(model, case['seed'])
        samples = [S.make_img_inputs(1, case['input_size'], seed=s) for s in range(2)]
    else:
        wl = bench.Workload('finetune', None, n_variants=2)
        cfg, model, samples = wl.cfg, wl.model, wl.samples
    model = model.to(dev)
    dev_samples = [tuple(t.to(dev) for t in s) for s in samples]

    def run(i):
        with torch.no_grad():
            vf = model.voxel_features_cl(dev_samples[i % 2])
            return model._occ_from_head(vf)[0]

    def timed(n=5):
        for i in range(2):
            run(i)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            run(i)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    ref = [run(i).cpu().numpy() for i in range(2)]
    ms_rep = timed()
    model.set_camera_shard(CameraShard())
    got = [run(i).cpu().numpy() for i in range(2)]
    ms_sh = timed()
    same = all(np.array_equal(a, b) for a, b in zip(ref, got))
    flag = torch.tensor([int(same)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f'world {world}: sharded == unsharded on every rank: {bool(flag.item())}; '
              f'ms/sample unsharded {ms_rep:.2f}, camera-sharded {ms_sh:.2f} '
              f'(latency x{ms_rep / ms_sh:.2f})', flush=True)
    dist.destroy_process_group()
    if not flag.item():
        sys.exit(1)


if __name__ == '__main__':
    main()

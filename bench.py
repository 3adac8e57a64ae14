#!/usr/bin/env python
"""Benchmark of the camera->voxel occupancy forward path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (configs[1]): ``preworld-7frame-finetune`` model dict with the derived
ResNet-50 @ 256x704 image side, random-init weights, synthetic 6-camera x
3-frame images -> 200x200x16 occupancy grid, forward only, fp32.  One *step* =
one forward pass of one sample (18 images) per GPU.  N > 1 (torchrun): every
rank runs its own samples (the path is data-parallel by sample) --
no data-path collective, weak scaling; time = max over ranks.

Prints ONE JSON line (see the keys below).  ``--impl reference`` times the CPU
restatement of the reference path (oracle/torch_ref.py -- the reference's own
.py files cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'frames/sec (7-frame, 6-cam 256x704->200x200x16)'
UNIT = 'frames/s'
WORKLOAD = ('preworld-7frame-finetune, derived ResNet-50 @ 6x3x256x704 -> '
            '200x200x16, bs=1/GPU, forward-only')
# dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full
# (profiles/r01n_hot_kernels.md), for the layer shapes that can lead the step
NCU_TRAFFIC = {
    'conv_halo 1x16x200x200x32->32 k333 s1 d1': 243.0e6,
    'conv_halo 1x16x200x200x32->64 k333 s1 d1': 198.2e6,
}
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0,
                      bf16_tflops_sustained=1400.0)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out = dict(FALLBACK_PEAKS)
            out.update({k: float(v) for k, v in d.items()
                        if isinstance(v, (int, float))})
            return out, 'measured'
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), 'fallback'


# ------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}',
                 f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None,
                'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------- per-launch profiler
class LaunchProfiler:
    """Wraps every C-ABI entry point with CUDA events on the launching stream
    (a separate pass after the timed region; never used for `value`)."""

    def __init__(self):
        from preworld_b200 import _lib
        self.L = _lib.lib()
        self.records = []
        self.orig = {}

    def __enter__(self):
        from preworld_b200 import _lib
        for name in _lib.SIGNATURES:
            if name in ('pw_abi_version', 'pw_launch_count',
                        'pw_lift_workspace_bytes', 'pw_conv_umma_supported',
                        'pw_conv_halo_supported', 'pw_conv_fold_supported',
                        'pw_conv_fold_n'):
                continue
            fn = getattr(self.L, name)
            self.orig[name] = fn
            setattr(self.L, name, self._wrap(name, fn))
        return self

    def _wrap(self, name, fn):
        def call(*args):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, e0, e1, self._work(name, args),
                                 self._shape(name, args)))
            return rc
        return call

    @staticmethod
    def _shape(name, a):
        if name in ('pw_conv_fwd', 'pw_conv_umma_fwd', 'pw_conv_halo_fwd',
                    'pw_conv_fold_fwd'):
            d = a[0]._obj
            return (f'{d.n}x{d.d}x{d.h}x{d.w}x{d.cin}->{d.cout} '
                    f'k{d.kd}{d.kh}{d.kw} s{d.sw} d{d.dw}')
        return ''

    @staticmethod
    def _work(name, a):
        """(algorithmic flops, algorithmic bytes) of one call."""
        if name in ('pw_conv_fwd', 'pw_conv_umma_fwd', 'pw_conv_halo_fwd',
                    'pw_conv_fold_fwd'):
            d = a[0]._obj
            m = d.n * d.od * d.oh * d.ow
            k = d.kd * d.kh * d.kw * d.cin
            r = a[5] if name == 'pw_conv_fwd' else a[6]
            res = 4 * m * d.cout if r is not None and r.value else 0
            return (2.0 * m * k * d.cout,
                    4.0 * (d.n * d.d * d.h * d.w * d.cin + m * d.cout
                           + k * d.cout) + res)
        if name == 'pw_lift_fused':
            b, n, dd, h, w, c, gx, gy, gz = a[10:19]
            # depth 4NDhw + feat 4NhwC + output 4ZYXC (the fused
            # path has no rank / interval arrays in its algorithmic minimum)
            return (2.0 * b * n * dd * h * w * c,
                    4.0 * b * (n * dd * h * w + n * h * w * c
                               + gx * gy * gz * c))
        if name == 'pw_cost_volume':
            n, h, w, c, dd = a[7:12]
            return (0.0, 4.0 * n * h * w * (2 * c + dd))
        return (0.0, 0.0)

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.L, name, fn)

    def summary(self, steps):
        torch.cuda.synchronize()
        agg = {}
        layers = {}
        for name, e0, e1, (fl, by), shape in self.records:
            ms = e0.elapsed_time(e1)
            # pw_conv_fold_fwd launches the same CUDA kernel (conv_halo_kernel)
            fam = 'pw_conv_halo_fwd' if name == 'pw_conv_fold_fwd' else name
            a = agg.setdefault(fam, [0, 0.0, 0.0, 0.0])
            a[0] += 1; a[1] += ms; a[2] += fl; a[3] += by
            if shape:
                b = layers.setdefault(name[3:-4] + ' ' + shape, [0, 0.0, 0.0])
                b[0] += 1; b[1] += ms; b[2] += fl
        self.layers = {
            k: dict(launches_per_step=v[0] / steps, ms_per_step=v[1] / steps,
                    tflops=v[2] / (v[1] * 1e-3) / 1e12 if v[1] else 0.0)
            for k, v in sorted(layers.items(), key=lambda kv: -kv[1][1])}
        out = {}
        for name, (cnt, ms, fl, by) in agg.items():
            out[name] = dict(launches_per_step=cnt / steps,
                             ms_per_step=ms / steps,
                             tflops=fl / (ms * 1e-3) / 1e12 if ms else 0.0,
                             gbs=by / (ms * 1e-3) / 1e9 if ms else 0.0,
                             gflop_per_step=fl / steps / 1e9,
                             mb_per_step=by / steps / 1e6)
        return out


# ---------------------------------------------------------------- workloads
def build_workload(n_variants=4):
    from preworld_b200 import build_model, model_cfg
    from preworld_b200 import synthetic as S
    cfg = model_cfg('finetune', 'r50', (256, 704))
    model = build_model(cfg).eval()
    S.lively_init_(model, 0)
    samples = [S.make_img_inputs(1, (256, 704), seed=s)
               for s in range(n_variants)]
    return cfg, model, samples


def cpu_forward_seconds(cfg, model_sd, sample, threads=None):
    from oracle import torch_ref
    if threads:
        torch.set_num_threads(threads)
    pc = torch_ref.PathConfig(cfg)
    t0 = time.perf_counter()
    with torch.no_grad():
        torch_ref.preworld_simple_test(model_sd, pc, sample)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm: the restated reference path on the host cores, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, model, samples = build_workload(2)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    budget = float(os.environ.get('PW_REF_BUDGET_S', '420'))
    t_begin = time.perf_counter()
    for i in range(args.warmup):
        cpu_forward_seconds(cfg, sd, samples[i % 2])
        if time.perf_counter() - t_begin > budget * 0.4:
            break
    times = []
    for i in range(args.steps):
        times.append(cpu_forward_seconds(cfg, sd, samples[i % 2]))
        if time.perf_counter() - t_begin > budget:
            break
    done = len(times)
    total = sum(times)
    value = done / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': done, 'steps_requested': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / done,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD + ' [CPU restatement of the '
                   'reference path, oracle/torch_ref.py; derived R50 config]'},
        'cpu_baseline': {'value': value, 'unit': UNIT,
                         'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': f'{done} full forward passes of 1 sample '
                                   '(18 images) each'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from preworld_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a CUDA device '
                         '(there is no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    assert world == args.gpus, (world, args.gpus)

    cfg, model, samples = build_workload(4)
    sd_cpu = {k: v.detach().clone() for k, v in model.state_dict().items()} \
        if rank == 0 else None
    model = model.to(dev)
    dev_samples = [tuple(t.to(dev) for t in s) for s in samples]
    pin_samples = [tuple(t.pin_memory() for t in s) for s in samples]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        with torch.no_grad():
            vf = model.voxel_features_cl(dev_samples[i % len(dev_samples)])
            occ, _ = model._occ_from_head(vf)
        return occ

    def step_e2e(i):
        # the public call on the loader's HOST batch (pinned): the model copies
        # it to the device itself (H2D inside the timed region) and returns
        # the occupancy grid as a numpy array (D2H inside the timed region)
        host = pin_samples[i % len(pin_samples)]
        with torch.no_grad():
            out = model(return_loss=False, img_inputs=[host], img_metas=[None])
        return out['semantic_occ'][0], out['geo_occ'][0]

    # ---- device-resident timing -------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()

    # ---- end-to-end timing (host buffers, H2D + D2H inside) ----------------
    # the public call replays the forward as one CUDA graph (one capture per
    # input shape, done by the first untimed call below)
    model.enable_cuda_graph()
    for i in range(3):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        occ_np = step_e2e(i)
    e1.record()
    barrier()
    ms_e2e_dev = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e_dev], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()
    h2d = sum(t.numel() * t.element_size() for t in pin_samples[0])
    d2h = int(sum(a.nbytes for a in occ_np))     # semantic_occ + geo_occ grids

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel profile (rank 0, outside the timed regions) ------------
    peaks, peak_src = load_peaks()
    kernels, roof = {}, None
    if not args.no_profile:
        psteps = 2
        with LaunchProfiler() as prof:
            for i in range(psteps):
                step_resident(i)
        kernels = prof.summary(psteps)
        if os.environ.get('PW_BENCH_LAYERS'):
            with open(os.environ['PW_BENCH_LAYERS'], 'w') as f:
                json.dump(prof.layers, f, indent=1)
        top = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        k = kernels[top]
        if top in ('pw_conv_umma_fwd', 'pw_conv_halo_fwd'):
            kname = ('conv_halo_kernel (pw_conv_halo_fwd: halo-resident tcgen05 '
                     'kind::tf32 implicit GEMM, A operand in TMEM, 3xTF32 '
                     'split, TMA)') if top == 'pw_conv_halo_fwd' else (
                     'conv_umma_kernel (pw_conv_umma_fwd: tcgen05 kind::tf32 '
                     'implicit GEMM, 3xTF32 split, TMA)')
            # the roofline line is quoted on the kernel's heaviest LAYER SHAPE
            # (per-launch figures); the whole family is reported next to it
            fams = ('conv_halo ', 'conv_fold ') if top == 'pw_conv_halo_fwd' \
                else (top[3:-4] + ' ',)
            lname, lay = max(((n, v) for n, v in prof.layers.items()
                              if n.startswith(fams)),
                             key=lambda kv: kv[1]['ms_per_step'])
            roof = {'kernel': kname, 'layer': lname,
                    'bound': 'tensor', 'achieved': lay['tflops'],
                    'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': lay['tflops'] / peaks['bf16_tflops_sustained'],
                    'traffic': NCU_TRAFFIC.get(lname), 'peak_source': peak_src,
                    'launch_us': 1e3 * lay['ms_per_step'] / lay['launches_per_step'],
                    'share_of_step': k['ms_per_step'] / (ms / args.steps),
                    'family': {'launches_per_step': k['launches_per_step'],
                               'ms_per_step': k['ms_per_step'],
                               'achieved': k['tflops'],
                               'frac': k['tflops'] / peaks['bf16_tflops_sustained']},
                    'note': 'achieved = ALGORITHMIC fp32 conv FLOPs of one launch '
                            '/ its average duration (CUDA events); the kernel '
                            'executes 3 tf32 MMAs per algorithmic MMA (tf32 dense '
                            'peak is half the bf16 peak), so the executed-tensor-'
                            'work fraction is 6x this; traffic = ncu dram bytes '
                            'read+write of one launch of this layer shape '
                            '(profiles/r01n_hot_kernels.md)'}
        elif top == 'pw_conv_fwd':
            roof = {'kernel': 'conv_igemm_kernel (pw_conv_fwd, fp32 SIMT '
                              'implicit GEMM; all conv/linear layers)',
                    'bound': 'tensor', 'achieved': k['tflops'],
                    'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': k['tflops'] / peaks['bf16_tflops_sustained'],
                    'traffic': None, 'peak_source': peak_src,
                    'share_of_step': k['ms_per_step'] / (ms / args.steps),
                    'note': 'fp32 FFMA kernel measured against the dense bf16 '
                            'tensor peak (the path it must move to); '
                            'fp32-SIMT nominal peak is ~72 TFLOP/s'}
        else:
            roof = {'kernel': top, 'bound': 'hbm', 'achieved': k['gbs'],
                    'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': k['gbs'] / peaks['hbm_gbs'], 'traffic': None,
                    'peak_source': peak_src,
                    'share_of_step': k['ms_per_step'] / (ms / args.steps)}
        lf = kernels.get('pw_lift_fused')
        if lf:
            lf['hbm_frac'] = lf['gbs'] / peaks['hbm_gbs']

    # ---- CPU baseline (bounded sample: one forward on the host cores) ------
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sec = cpu_forward_seconds(cfg, sd_cpu, samples[0], cores)
        cpu = {'value': 1.0 / sec, 'unit': UNIT,
               'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '1 full forward pass of 1 sample (18 images) through '
                         'oracle/torch_ref.py (CPU restatement of the '
                         'reference path), no warm-up'}

    line = {
        'metric': METRIC, 'value': world * args.steps / (ms_max * 1e-3),
        'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sharding': 'by sample (replicas), '
                   'no data-path collective',
                   'l2': 'per-step activation working set (>2 GB) exceeds the '
                         '126 MB L2; inputs rotate over 4 samples (156 MB)',
                   'derived_config': True},
        'e2e': {'value': world * args.steps / (ms_e2e * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / args.steps,
                'api': 'model(return_loss=False, img_inputs=[...]) with '
                       'pinned host tensors, model.enable_cuda_graph() '
                       '(forward replayed as CUDA graphs; images uploaded frame by frame under the stem of the previous frame)'},
        'gpu_launches': int(launches),
        'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu,
        'kernels': kernels,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

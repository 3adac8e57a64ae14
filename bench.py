#!/usr/bin/env python
"""Benchmark of the camera->voxel occupancy forward path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config finetune|pretrain|traj|stress] [--shard sample|camera]

Workload (configs[1]): ``preworld-7frame-finetune`` model dict with the derived
ResNet-50 @ 256x704 image side, random-init weights, synthetic 6-camera x
3-frame images -> 200x200x16 occupancy grid, forward only, fp32.  One *step* =
one forward pass of one sample (18 images) per GPU.  N > 1 (torchrun): every
rank runs its own samples (the path is data-parallel by sample) --
no data-path collective, weak scaling; time = max over ranks.

``--config`` selects another BASELINE.json config (default: the headline
configs[1]); ``--shard camera`` (N > 1) splits ONE sample's cameras over the
ranks with one NCCL all-gather (latency mode, strong scaling) instead of
replicas.  The default run also takes short measurements of configs[2] / [3]
(N = 1) or of the camera-sharded mode (N > 1) and reports them under
``other_configs`` / ``camera_shard``.

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
# (profiles/r02f_hot_kernels.md: the capture of the same build, taken by
# tools/gpu_profile.sh), for the layer shapes that can lead the step; algorithmic
# bytes of the two shapes: 246 MB (32->32 + residual) and 246 MB (32->64)
NCU_TRAFFIC = {
    'conv_halo 1x16x200x200x32->32 k333 s1 d1': 241.3e6,
    'conv_halo 1x16x200x200x32->64 k333 s1 d1': 196.7e6,
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
                        'pw_conv_fold_n', 'pw_mlp2_supported'):
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
            # (curr, prev, cam, xs, ys, ds, out, ld, n, h, w, c, d, ...): SURVEY §8d --
            # two feature maps in (4nhwc each), the cost volume out (4nhwd)
            n, h, w, c, dd = a[8:13]
            return (0.0, 4.0 * n * h * w * (2 * c + dd))
        if name == 'pw_mlp2':
            m, c1, hidden, n2 = a[2], a[3], a[7], a[12]
            return (2.0 * m * (c1 * hidden + hidden * n2), 4.0 * m * (c1 + n2))
        if name == 'pw_occhead_tail':
            cin, gx, gy, gz = a[2], a[16], a[17], a[18]
            return (0.0, float(gx * gy * gz) * (4 * cin + 2))
        if name == 'pw_render_rays':
            d, r = a[0]._obj, a[2]
            vox = d.gx * d.gy * d.gz
            return (0.0, 4.0 * vox * (1 + d.n_sem + 3) + r * 16 * 4.0
                    + r * (3 + d.n_sem + 3) * 4.0)
        if name == 'pw_layernorm':
            rows, c = a[7], a[8]
            return (0.0, 8.0 * rows * c)
        if name == 'pw_patch_merge_ln':
            b, h, w, c = a[2:6]
            return (0.0, 4.0 * b * (h * w * c + ((h + 1) // 2) * ((w + 1) // 2) * 4 * c))
        if name == 'pw_window_attention':
            b, h, w, c, heads, ws = a[6:12]
            nwin = b * -(-h // ws) * -(-w // ws)
            n = ws * ws
            # q k^T and p v over the padded windows; qkv in, attention output out
            return (4.0 * nwin * heads * n * n * 32, 16.0 * b * h * w * c)
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

    HBM_KERNELS = ('pw_lift_fused', 'pw_lift_pool', 'pw_cost_volume', 'pw_mlp2',
                   'pw_occhead_tail', 'pw_render_rays', 'pw_layernorm', 'pw_patch_merge_ln')


# ---------------------------------------------------------------- workloads
CONFIGS = {
    # name -> (BASELINE.json configs[] index, metric, unit)
    'finetune': (1, METRIC, UNIT),
    'pretrain': (2, 'rays/sec (preworld-7frame-pretrain: trunk + attribute projection + '
                    'volume rendering of 38400 rays per sample)', 'rays/s'),
    'traj': (3, 'samples/sec (preworld-7frame-finetune-traj: 18 images -> 7 occupancy '
                'grids, 6 state-conditioned forecasting steps)', 'samples/s'),
    'stress': (4, 'frames/sec (ResNet-101, 400x400x16 voxel grid, bs=4)', 'frames/s'),
    # not a BASELINE.json config: the model dict the reference actually ships
    # (configs/preworld/nuscenes/preworld-7frame-finetune.py: Swin-B + FPN_LSS @ 512x1408)
    'shipped': (None, 'frames/sec (shipped preworld-7frame-finetune: Swin-B @ 6x3x512x1408 '
                      '-> 200x200x16)', 'frames/s'),
}
WORKLOADS = {
    'finetune': WORKLOAD,
    'pretrain': ('preworld-7frame-pretrain, derived ResNet-50 @ 6x3x256x704 -> 200x200x16 '
                 '-> attribute projection -> 38400 rays x 417 samples, bs=1/GPU, forward-only'),
    'traj': ('preworld-7frame-finetune-traj, derived ResNet-50 @ 6x3x256x704 -> 200x200x16, '
             '6 forecasting steps -> 7 grids, bs=1, forward-only'),
    'stress': ('preworld-7frame-finetune, derived ResNet-101 @ 6x3x256x704 -> 400x400x16 '
               '(0.2 m voxels), 4 samples per step, forward-only'),
    'shipped': ('preworld-7frame-finetune AS SHIPPED: SwinTransformer (Swin-B, window 12) + '
                'FPN_LSS @ 6x3x512x1408 -> 200x200x16, bs=1/GPU, forward-only'),
}
N_RAYS = 38400


class Workload:
    """One BASELINE.json config: model, synthetic samples, the device-resident
    step and the end-to-end step through the public call."""

    def __init__(self, name, dev, n_variants=4):
        from preworld_b200 import build_model, configs, model_cfg
        from preworld_b200 import synthetic as S
        self.name, self.dev = name, dev
        self.units = 1                       # metric units per step
        if name == 'finetune':
            cfg = model_cfg('finetune', 'r50', (256, 704))
        elif name == 'pretrain':
            cfg = model_cfg('pretrain', 'r50', (256, 704))
            self.units = N_RAYS
        elif name == 'traj':
            cfg = model_cfg('finetune-traj', 'r50', (256, 704))
        elif name == 'stress':
            cfg = model_cfg('finetune', 'r101', (256, 704),
                            grid=configs.grid_config(x=(-40, 40, 0.2), y=(-40, 40, 0.2)))
            self.units = 4
            n_variants = 4
        elif name == 'shipped':
            cfg = model_cfg('finetune', 'swin')
            cfg['img_backbone']['with_cp'] = False
            n_variants = min(n_variants, 2)
        else:
            raise ValueError(name)
        self.cfg = cfg
        self.model = build_model(cfg).eval()
        S.lively_init_(self.model, 0)
        hw = tuple(cfg['img_view_transformer']['input_size'])
        self.samples = [S.make_img_inputs(1, hw, seed=s) for s in range(n_variants)]
        self.rays = self.ego = None
        if name == 'pretrain':
            self.rays = [S.make_rays(s, N_RAYS, seed=100 + i) for i, s in enumerate(self.samples)]
        if name == 'traj':
            self.ego = [S.make_ego_states(1, seed=200 + i) for i in range(n_variants)]
        self.sd_cpu = None

    def to_device(self, keep_cpu_weights=False):
        if keep_cpu_weights:
            self.sd_cpu = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        dev = self.dev
        self.model = self.model.to(dev)
        self.dev_samples = [tuple(t.to(dev) for t in s) for s in self.samples]
        self.pin_samples = [tuple(t.pin_memory() for t in s) for s in self.samples]
        if self.rays:
            self.dev_rays = [r.to(dev) for r in self.rays]
            self.pin_rays = [r.pin_memory() for r in self.rays]
        if self.ego:
            self.dev_ego = [e.to(dev) for e in self.ego]
            self.pin_ego = [e.pin_memory() for e in self.ego]
        return self

    # -- one step with the inputs resident in HBM -------------------------------
    def step_resident(self, i):
        m, k = self.model, i % len(self.samples)
        with torch.no_grad():
            if self.name in ('finetune', 'shipped'):
                vf = m.voxel_features_cl(self.dev_samples[k])
                return m._occupancy_dev(vf)
            if self.name == 'pretrain':
                return m.render_forward(self.dev_samples[k], self.dev_rays[k])
            if self.name == 'traj':
                vf = m.voxel_features_cl(self.dev_samples[k])
                return m._occupancy_dev(vf, [self.dev_ego[k]])
            out = None
            for j in range(4):               # stress: 4 samples per step
                vf = m.voxel_features_cl(self.dev_samples[(k + j) % len(self.samples)])
                out = m._occupancy_dev(vf)
            return out

    # -- the same through the public call on pinned HOST tensors -----------------
    def prepare_e2e(self):
        if self.name in ('finetune', 'traj', 'stress', 'shipped') and \
                getattr(self.model, 'camera_shard', None) is None:
            self.model.enable_cuda_graph()

    def step_e2e(self, i):
        m, k = self.model, i % len(self.samples)
        with torch.no_grad():
            if self.name in ('finetune', 'shipped'):
                out = m(return_loss=False, img_inputs=[self.pin_samples[k]], img_metas=[None])
                return [out['semantic_occ'][0], out['geo_occ'][0]]
            if self.name == 'pretrain':
                r = m.render_forward(self.pin_samples[k], self.pin_rays[k])[0]
                return [r[n].cpu().numpy() for n in ('render_depth', 'render_semantic',
                                                     'render_color')]
            if self.name == 'traj':
                out = m(return_loss=False, img_inputs=[self.pin_samples[k]], img_metas=[None],
                        temporal_ego_states=[[self.pin_ego[k]]])
                return [v[0] for v in out.values()]
            outs = []
            for j in range(4):
                out = m(return_loss=False,
                        img_inputs=[self.pin_samples[(k + j) % len(self.samples)]],
                        img_metas=[None])
                outs += [out['semantic_occ'][0], out['geo_occ'][0]]
            return outs

    def h2d_bytes(self):
        n = sum(t.numel() * t.element_size() for t in self.pin_samples[0])
        if self.rays:
            n += self.pin_rays[0].numel() * 4
        if self.ego:
            n += self.pin_ego[0].numel() * 4
        return n * (4 if self.name == 'stress' else 1)

    # -- CPU restatement of the reference path (oracle/torch_ref.py) ------------
    def cpu_seconds(self, k=0, threads=None):
        from oracle import torch_ref
        if threads:
            torch.set_num_threads(threads)
        sd = self.sd_cpu if self.sd_cpu is not None else \
            {k_: v.detach() for k_, v in self.model.state_dict().items()}
        pc = torch_ref.PathConfig(self.cfg)
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.name == 'traj':
                torch_ref.preworld4d_simple_test(sd, pc, self.samples[k], [self.ego[k]])
            elif self.name == 'pretrain':
                vf = torch_ref.voxel_features(sd, pc, self.samples[k])
                dens, sem, col = torch_ref.attribute_projection(sd, vf)
                ng = torch_ref.NerfGeometry(self.cfg['nerf_head']['point_cloud_range'])
                torch_ref.render_rays(ng, self.rays[k][0], self.samples[k][6][0],
                                      dens[0], sem[0], col[0])
            else:
                torch_ref.preworld_simple_test(sd, pc, self.samples[k])
        return time.perf_counter() - t0


def run_reference(args, rank, world):
    """CPU arm: the restated reference path on the host cores, rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = Workload(args.config, None, 2)
    idx, metric, unit = CONFIGS[args.config]
    if args.config in ('stress', 'shipped'):
        print(json.dumps({'impl': 'reference', 'unavailable':
                          f'--config {args.config} is not timed on the CPU (4 x R101 400x400x16 / '
                          '18 x Swin-B 512x1408 forwards take minutes each); see --config '
                          'finetune'}), flush=True)
        return
    budget = float(os.environ.get('PW_REF_BUDGET_S', '240'))
    t_begin = time.perf_counter()
    for i in range(args.warmup):
        wl.cpu_seconds(i % 2)
        if time.perf_counter() - t_begin > budget * 0.4:
            break
    times = []
    for i in range(args.steps):
        times.append(wl.cpu_seconds(i % 2))
        if time.perf_counter() - t_begin > budget:
            break
    done = len(times)
    total = sum(times)
    value = wl.units * done / total
    line = {
        'impl': 'reference', 'metric': metric, 'value': value, 'unit': unit,
        'n_gpus': args.gpus, 'steps': done, 'steps_requested': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / done,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args.config, 'sample'),
        'reference_note': 'CPU restatement of the reference path (oracle/torch_ref.py, pinned '
                          'to the verbatim reference files by tests/golden/REPORT.json); the '
                          'reference .py files cannot travel to the GPU box',
        'cpu_baseline': {'value': value, 'unit': unit,
                         'cores': torch.get_num_threads(), 'kind': 'port',
                         'sample': f'{done} full forward passes of 1 sample '
                                   '(18 images) each'},
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(name, shard):
    cfg = {'workload': WORKLOADS[name],
           'baseline_config_index': CONFIGS[name][0],
           'sharding': 'by sample (replicas), no data-path collective' if shard == 'sample'
           else 'within-sample camera sharding: image side per camera block, ONE NCCL '
                'all-gather of depth + context (0.33 MB per camera and frame), lift + 3-D '
                'stages replicated',
           'l2': 'per-step activation working set (>2 GB) exceeds the 126 MB L2; inputs '
                 'rotate over 4 samples (156 MB)',
           'derived_config': name != 'shipped'}
    return cfg


def timed(fn, steps, barrier, dev, world, dist):
    """K steps between barriers; CUDA events on the launching stream; max over ranks."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for i in range(steps):
        out = fn(i)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), out


def measure(wl, steps, warmup, barrier, dev, world, dist, e2e=True):
    """-> dict(ms resident, ms e2e, launches, h2d, d2h) for one workload."""
    from preworld_b200 import _lib
    for i in range(warmup):
        wl.step_resident(i)
    n0 = _lib.launch_count()
    ms, _ = timed(wl.step_resident, steps, barrier, dev, world, dist)
    res = {'ms': ms, 'launches': _lib.launch_count() - n0, 'steps': steps}
    if e2e:
        wl.prepare_e2e()
        for i in range(3):
            wl.step_e2e(i)
        ms_e, out = timed(wl.step_e2e, steps, barrier, dev, world, dist)
        res.update(ms_e2e=ms_e, h2d=wl.h2d_bytes(),
                   d2h=int(sum(a.nbytes for a in out)))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='finetune', choices=sorted(CONFIGS))
    ap.add_argument('--shard', default='sample', choices=['sample', 'camera'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--no-extras', action='store_true',
                    help='skip the short measurements of the other configs / '
                         'the camera-sharded mode appended to the default run')
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
    if args.shard == 'camera' and world == 1:
        args.shard = 'sample'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    idx, metric, unit = CONFIGS[args.config]
    wl = Workload(args.config, dev).to_device(keep_cpu_weights=(rank == 0))
    if args.shard == 'camera':
        from preworld_b200.parallel import CameraShard
        wl.model.set_camera_shard(CameraShard())
    steps = args.steps
    if args.config in ('stress', 'shipped'):
        steps = max(1, min(steps, 20))

    clocks = ClockSampler(local_rank)
    clocks.start()
    r = measure(wl, steps, args.warmup, barrier, dev, world, dist)
    clk = clocks.stop()
    ms, ms_e2e, launches = r['ms'], r['ms_e2e'], r['launches']
    # replicas: every rank runs its own samples (weak); camera shard: all ranks
    # work on the same sample (strong)
    jobs = world if args.shard == 'sample' else 1

    # ---- short extra measurements (all ranks take part) -----------------------
    extras = {}
    if not args.no_extras and args.config == 'finetune' and args.shard == 'sample':
        k = max(3, min(10, steps))
        if world == 1:
            for name in ('traj', 'pretrain', 'shipped'):
                k = max(3, min(10, steps)) if name != 'shipped' else max(3, min(5, steps))
                w2 = Workload(name, dev, 2).to_device()
                r2 = measure(w2, k, 3, barrier, dev, world, dist)
                _, m2, u2 = CONFIGS[name]
                extras[name] = {
                    'metric': m2, 'unit': u2, 'steps': k,
                    'value': w2.units * k / (r2['ms'] * 1e-3),
                    'ms_per_step': r2['ms'] / k,
                    'e2e': {'value': w2.units * k / (r2['ms_e2e'] * 1e-3), 'unit': u2,
                            'ms_per_step': r2['ms_e2e'] / k,
                            'h2d_bytes_per_step': r2['h2d'], 'd2h_bytes_per_step': r2['d2h']},
                    'gpu_launches_per_step': r2['launches'] / k,
                    'config': bench_config(name, 'sample')}
                if name == 'traj':
                    extras[name]['grids_per_s'] = 7 * extras[name]['value']
                elif name == 'pretrain':
                    extras[name]['samples_per_s'] = extras[name]['value'] / N_RAYS
                del w2
                torch.cuda.empty_cache()
        else:
            from preworld_b200.parallel import CameraShard
            wl.model.enable_cuda_graph(False)
            wl.model.set_camera_shard(CameraShard())
            r2 = measure(wl, k, 3, barrier, dev, world, dist, e2e=False)
            wl.model.set_camera_shard(None)
            extras['camera_shard'] = {
                'metric': 'frames/sec of ONE sample stream split by camera over the ranks '
                          '(latency mode)', 'unit': UNIT, 'steps': k, 'scaling': 'strong',
                'value': k / (r2['ms'] * 1e-3), 'ms_per_sample': r2['ms'] / k,
                'replica_ms_per_sample': ms / steps,
                'config': bench_config('finetune', 'camera')}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel profile (rank 0, outside the timed regions) ------------
    peaks, peak_src = load_peaks()
    kernels, roof = {}, None
    if not args.no_profile:
        if args.shard == 'camera':
            wl.model.set_camera_shard(None)          # rank 0 profiles alone
        wl.model.enable_cuda_graph(False)
        psteps = 2
        with LaunchProfiler() as prof:
            for i in range(psteps):
                wl.step_resident(i)
        kernels = prof.summary(psteps)
        if os.environ.get('PW_BENCH_LAYERS'):
            with open(os.environ['PW_BENCH_LAYERS'], 'w') as f:
                json.dump(prof.layers, f, indent=1)
        for kn in LaunchProfiler.HBM_KERNELS:
            if kn in kernels:
                kernels[kn]['hbm_frac'] = kernels[kn]['gbs'] / peaks['hbm_gbs']
        top = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        k = kernels[top]
        step_ms = ms / steps
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
                    'frac_of_3xtf32_ceiling': lay['tflops'] * 6 / peaks['bf16_tflops_sustained'],
                    'traffic': NCU_TRAFFIC.get(lname),
                    'traffic_source': 'dram__bytes_read.sum + dram__bytes_write.sum of this layer '
                                      'shape from the ncu --set full capture of the same build, '
                                      'profiles/r02f_hot_kernels.md (a profiler pass cannot run '
                                      'inside the timed bench)',
                    'peak_source': peak_src,
                    'launch_us': 1e3 * lay['ms_per_step'] / lay['launches_per_step'],
                    'share_of_step': k['ms_per_step'] / step_ms,
                    'family': {'launches_per_step': k['launches_per_step'],
                               'ms_per_step': k['ms_per_step'],
                               'achieved': k['tflops'],
                               'frac': k['tflops'] / peaks['bf16_tflops_sustained']},
                    'note': 'achieved = ALGORITHMIC fp32 conv FLOPs of one launch '
                            '/ its average duration (CUDA events); the kernel '
                            'executes 3 tf32 MMAs per algorithmic MMA (tf32 dense '
                            'peak is half the bf16 peak), so its own ceiling is 1/6 of '
                            'the bf16 peak (frac_of_3xtf32_ceiling)'}
        else:
            bound = 'hbm' if k['gbs'] > 0 else 'tensor'
            roof = {'kernel': top, 'bound': bound,
                    'achieved': k['gbs'] if bound == 'hbm' else k['tflops'],
                    'peak': peaks['hbm_gbs'] if bound == 'hbm' else peaks['bf16_tflops_sustained'],
                    'unit': 'GB/s' if bound == 'hbm' else 'TFLOP/s',
                    'traffic': None, 'peak_source': peak_src,
                    'share_of_step': k['ms_per_step'] / step_ms}
            roof['frac'] = roof['achieved'] / roof['peak']

    # ---- CPU baseline (bounded sample: one forward on the host cores) ------
    cpu = None
    if not args.no_cpu_baseline and args.config not in ('stress', 'shipped'):
        cores = os.cpu_count() or 1
        sec = wl.cpu_seconds(0, cores)
        cpu = {'value': wl.units / sec, 'unit': unit,
               'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': '1 full forward pass of 1 sample (18 images) through '
                         'oracle/torch_ref.py (CPU restatement of the '
                         'reference path), no warm-up'}

    line = {
        'metric': metric, 'value': jobs * wl.units * steps / (ms * 1e-3),
        'unit': unit, 'n_gpus': world, 'steps': steps,
        'warmup': args.warmup, 'ms_per_step': ms / steps,
        'higher_is_better': True,
        'scaling': 'weak' if args.shard == 'sample' else 'strong',
        'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': bench_config(args.config, args.shard),
        'e2e': {'value': jobs * wl.units * steps / (ms_e2e * 1e-3), 'unit': unit,
                'h2d_bytes_per_step': r['h2d'], 'd2h_bytes_per_step': r['d2h'],
                'ms_per_step': ms_e2e / steps,
                'api': 'model(return_loss=False, img_inputs=[...]) with pinned host '
                       'tensors' + (', model.enable_cuda_graph() (forward replayed as '
                       'CUDA graphs; images uploaded chunk by chunk under the stem of '
                       'the previous chunk)' if args.shard == 'sample' and
                       args.config != 'pretrain' else '')},
        'gpu_launches': int(launches),
        'clocks': clk, 'roofline': roof, 'cpu_baseline': cpu,
        'kernels': kernels,
    }
    if 'camera_shard' in extras:
        line['camera_shard'] = extras.pop('camera_shard')
    if extras:
        line['other_configs'] = extras
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""End-to-end parity of the plugin detectors (all arithmetic in the C-ABI CUDA
library) against (a) the golden fixtures produced by the verbatim reference
and (b) the CPU oracle run on the same seeded inputs.

Bars (BASELINE.json north_star): class logits within 1e-3 relative fp32;
argmax occupancy identical -- except voxels whose top-2 logit margin in the
oracle is below the float tolerance (near ties flip between ANY two fp32
evaluation orders: the reference's own unstable argsort makes its GPU result
non-deterministic at that level, and the CPU oracle itself differs from the
verbatim reference in 1 voxel of 640 000, see tests/golden/REPORT.json)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_ref                                   # noqa: E402
from oracle.cases import (CASES, build_case_inputs, model_cfg_for,   # noqa
                          stage_sample)
from preworld_b200 import build_model, ops                     # noqa: E402
from preworld_b200 import synthetic as S                       # noqa: E402

REL_TOL = 1e-3


def _model(case):
    m = build_model(model_cfg_for(case)).eval()
    S.lively_init_(m, case['seed'])
    return m


def _run_stages(model, inputs):
    """Run the trunk capturing the same stages oracle/make_golden.py hooks."""
    st = {}
    calls = {'n': 0}

    def vt_hook(mod, inp, out):
        fid = 1 - calls['n']
        st[f'lifted_{fid}'], st[f'depth_{fid}'] = out[0], out[1]
        calls['n'] += 1
    hs = [model.img_view_transformer.register_forward_hook(vt_hook),
          model.img_bev_encoder_neck.register_forward_hook(
              lambda m, i, o: st.__setitem__('encoded', o))]
    vf = model.voxel_features_cl(inputs)
    for h in hs:
        h.remove()
    st['voxel_feats'] = vf.permute(0, 3, 2, 1, 4)            # [B,X,Y,Z,C]
    if model.if_post_finetune:
        lg = model.occupancy_head.logits_cl(vf[:1], True)
        st['logits'] = lg.permute(0, 4, 3, 2, 1)             # [1,18,X,Y,Z]
    return vf, st


def _check_samples(fx, st, name):
    worst = {}
    for key in fx.files:
        kind, k = key.split('/', 1)
        if kind != 'sample':
            continue
        got = stage_sample(st[k]).numpy()
        scale = max(float(fx['stats/' + k][2]), 1e-6)       # max |ref|
        err = float(np.abs(got - fx[key]).max()) / scale
        worst[k] = err
        assert err < REL_TOL, (name, k, err)
    return worst


def _margin_ok(occ, want, logits_ref, tol=REL_TOL, verbose=False):
    """Every argmax mismatch must sit on a near tie of the oracle logits: top-2
    margin below ``tol`` x max|logit|."""
    bad = np.argwhere(occ != want)
    lg = logits_ref[0].permute(1, 2, 3, 0).numpy()          # [X,Y,Z,18]
    scale = np.abs(lg).max()
    worst = 0.0
    for x, y, z in bad:
        top2 = np.sort(lg[x, y, z])[-2:]
        margin = float(top2[1] - top2[0]) / scale
        worst = max(worst, margin)
        if verbose:
            print(f'   voxel ({x},{y},{z}): got {occ[x, y, z]} want {want[x, y, z]} '
                  f'top-2 margin {margin:.2e} of max|logit|')
        assert margin < tol, ((x, y, z), top2, occ[x, y, z], want[x, y, z])
    if verbose and len(bad):
        print(f'   largest top-2 margin among the {len(bad)} flips: {worst:.2e}')
    return len(bad)


# tiny_swin_finetune: the SHIPPED image side (SwinTransformer + FPN_LSS) under the same detector
@pytest.mark.parametrize('name', ['tiny_finetune', 'tiny_pretrain', 'tiny_swin_finetune'])
def test_preworld_matches_reference_and_oracle(name, golden_dir):
    case = CASES[name]
    fx = np.load(os.path.join(golden_dir, name + '.npz'))
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    inputs, _ = build_case_inputs(case)
    dev_inputs = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        vf, st = _run_stages(model, dev_inputs)
        out = model.simple_test(None, None, img=dev_inputs)
    worst = _check_samples(fx, {k: v.cpu() for k, v in st.items()}, name)
    print(name, 'worst relative stage errors', worst)
    # oracle on the CPU for full tensors
    pc = torch_ref.PathConfig(model_cfg_for(case))
    ost = {}
    want = torch_ref.preworld_simple_test(sd, pc, inputs, ost)
    for k in ('lifted_0', 'lifted_1', 'encoded', 'voxel_feats'):
        ref = ost[k]
        err = (st[k].cpu() - ref).abs().max().item() / ref.abs().max().item()
        assert err < REL_TOL, (k, err)
    assert out['semantic_occ'][0].dtype == np.uint8
    assert out['semantic_occ'][0].shape == fx['out/semantic_occ'].shape
    if model.if_post_finetune:
        err = (st['logits'].cpu() - ost['logits']).abs().max().item() \
            / ost['logits'].abs().max().item()
        assert err < REL_TOL, err
        n_bad = _margin_ok(out['semantic_occ'][0], want['semantic_occ'][0],
                           ost['logits'])
        assert n_bad <= 0.001 * want['semantic_occ'][0].size
        n_fx = int((out['semantic_occ'][0] != fx['out/semantic_occ']).sum())
        assert n_fx <= 0.001 * fx['out/semantic_occ'].size + 2
        geo = out['geo_occ'][0]
        assert ((geo == 0) == (out['semantic_occ'][0] != 17)).all()
    else:
        # density path: thresholded argmax over the semantic MLP
        for k in ('semantic_occ', 'geo_occ'):
            mism = int((out[k][0] != fx['out/' + k]).sum())
            assert mism <= 0.001 * fx['out/' + k].size + 2, (k, mism)


@pytest.mark.parametrize('name', ['tiny_traj', 'tiny_pretrain_traj'])
def test_preworld4d_forecasting_matches_reference(name, golden_dir):
    case = CASES[name]
    fx = np.load(os.path.join(golden_dir, name + '.npz'))
    model = _model(case).cuda()
    inputs, extra = build_case_inputs(case)
    dev_inputs = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        out = model(return_loss=False, img_inputs=[dev_inputs],
                    img_metas=[None], **extra)
    keys = [k for k in fx.files if k.startswith('out/')]
    assert len(keys) == 14                      # 7 grids x (semantic, geo)
    for key in keys:
        k = key.split('/', 1)[1]
        mism = int((out[k][0] != fx[key]).sum())
        # six recursive steps amplify fp32 reordering; near ties only
        assert mism <= 0.002 * fx[key].size + 2, (k, mism)
    with torch.no_grad():
        vf, st = _run_stages(model, dev_inputs)
    _check_samples(fx, {k: v.cpu() for k, v in st.items()}, name)


def test_forecast_step_matches_oracle():
    case = CASES['tiny_traj']
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    g = torch.Generator().manual_seed(0)
    vf = torch.randn(2, 40, 40, 16, 32, generator=g)         # [B,X,Y,Z,C]
    ego = torch.randn(2, 1, 21, generator=g)
    want = torch_ref.forecast_step(sd, vf, ego)
    vf_cl = vf.permute(0, 3, 2, 1, 4).contiguous().cuda()
    with torch.no_grad():
        got = model.forecast_step(vf_cl, ego.cuda())
    got = got.permute(0, 3, 2, 1, 4).cpu()
    assert (got - want).abs().max().item() / want.abs().max().item() < 1e-5


def test_plan_trajectory_matches_oracle():
    """a18, the planning branch (preworld_temporal_traj.py:454-472 +
    DownScaleModule3DCustom, occupancy_head.py:180-200): three stride-2 2x2x2 convs on
    the library's [Z,Y,X] volume (spatially transposed kernels), global pool, ego fusion
    MLP, trajectory head -- against the CPU restatement, 1e-5 of max|ref|."""
    case = CASES['tiny_traj']
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    g = torch.Generator().manual_seed(5)
    vf = torch.randn(2, 40, 40, 16, 32, generator=g)         # [B,X,Y,Z,C]
    ego = torch.randn(2, 1, 21, generator=g)
    fused = torch_ref.forecast_step(sd, vf, ego)
    want = torch_ref.plan_trajectory(sd, fused, ego)
    assert want.shape == (2, 2)
    with torch.no_grad():
        fused_cl = model.forecast_step(vf.permute(0, 3, 2, 1, 4).contiguous().cuda(), ego.cuda())
        got = model.plan_trajectory(fused_cl, ego.cuda()).cpu()
        # the module's own forward takes the reference's [b,X,Y,Z,C] layout
        pooled = model.downscale(fused.cuda()).cpu()
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    x = fused.permute(0, 4, 1, 2, 3)
    for i in (1, 2, 3):
        x = torch.nn.functional.conv3d(x, sd[f'downscale.downscale{i}.weight'],
                                       sd[f'downscale.downscale{i}.bias'], stride=2)
    ref_pool = x.mean(dim=(2, 3, 4))
    assert pooled.shape == (2, 1, 1, 1, 128)
    assert (pooled.view(2, -1) - ref_pool).abs().max().item() <= 1e-5 * ref_pool.abs().max().item()


def test_attribute_projection_matches_oracle():
    case = CASES['tiny_pretrain']
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    g = torch.Generator().manual_seed(1)
    vf = torch.randn(1, 40, 40, 16, 32, generator=g) * 2
    d, s, c = torch_ref.attribute_projection(sd, vf)
    with torch.no_grad():
        attr = model.attributes_cl(vf.permute(0, 3, 2, 1, 4).contiguous().cuda())
    attr = attr.permute(0, 3, 2, 1, 4).cpu()
    for got, want in ((attr[..., 0], d), (attr[..., 2:19], s),
                      (attr[..., 19:22], c)):
        assert (got - want).abs().max().item() / want.abs().max().item() < 1e-5


def test_render_forward_runs_end_to_end():
    """Config 3 plumbing: trunk -> attribute projection -> ray march."""
    case = CASES['tiny_pretrain']
    # NerfHead hard-codes a 200x200x16 world (nerf_head.py:150): use the full
    # grid with the tiny image size
    from preworld_b200 import model_cfg
    cfg = model_cfg('pretrain', 'r50', (64, 176))
    model = build_model(cfg).eval()
    S.lively_init_(model, 3)
    model = model.cuda()
    inputs = S.make_img_inputs(1, (64, 176), seed=4)
    rays = S.make_rays(inputs, 512, seed=5)
    with torch.no_grad():
        res = model.render_forward(tuple(t.cuda() for t in inputs), rays.cuda())
    r = res[0]
    assert r['render_semantic'].shape == (512, 17)
    assert torch.isfinite(r['render_depth']).all()
    assert r['ray_mask'].all()


@pytest.mark.timeout(900)
def test_full_size_finetune_matches_reference(golden_dir):
    """BASELINE.json configs[0]/[1]: 6x3x256x704 -> 200x200x16, R50."""
    case = CASES['full_finetune']
    fx = np.load(os.path.join(golden_dir, 'full_finetune.npz'))
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    inputs, _ = build_case_inputs(case)
    dev_inputs = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        vf, st = _run_stages(model, dev_inputs)
        out = model.simple_test(None, None, img=dev_inputs)
    worst = _check_samples(fx, {k: v.cpu() for k, v in st.items()}, 'full')
    print('full_finetune worst relative stage errors', worst)
    occ, want = out['semantic_occ'][0], fx['out/semantic_occ']
    assert occ.shape == (200, 200, 16)
    n_bad = int((occ != want).sum())
    print('full_finetune argmax mismatches vs reference fixture:', n_bad)
    # round 2 (round-to-nearest split + accumulator-truncation gain):
    # 18 of 640 000; strict-fp32 cuDNN gives 1, torch's default TF32 cuDNN 3196
    # (test_reference_arithmetic_on_gpu_floor)
    assert n_bad <= 32
    # every mismatch must be a near tie of the oracle's logits (top-2 margin
    # below the 1e-3 float tolerance); the oracle runs here in ~5-15 s
    pc = torch_ref.PathConfig(model_cfg_for(case))
    ost = {}
    with torch.no_grad():
        want_o = torch_ref.preworld_simple_test(sd, pc, inputs, ost)
    err = (st['logits'].cpu() - ost['logits']).abs().max().item() \
        / ost['logits'].abs().max().item()
    # every flip sits on a near tie: measured top-2 margins <= 2e-6 of max|logit|
    n_o = _margin_ok(occ, want_o['semantic_occ'][0], ost['logits'],
                     tol=1e-5, verbose=True)
    print(f'full_finetune vs oracle: {n_o} near-tie voxels differ, '
          f'logits max rel err {err:.2e}')
    assert err < 1e-4                         # north_star bar: 1e-3
    # size-independent properties
    assert ((out['geo_occ'][0] == 0) == (occ != 17)).all()
    with torch.no_grad():
        out2 = model.simple_test(None, None, img=dev_inputs)
    assert (out2['semantic_occ'][0] == occ).all()      # deterministic
    lifted = st['lifted_0'].cpu()
    assert lifted.shape == (1, 32, 16, 200, 200)
    nz = (lifted.abs().sum(1) > 0).float().mean().item()
    assert 0.15 < nz < 0.35                   # ~142k of 640k voxels non-empty


def _oracle_with_gpu_convs(sd, pc, inputs, allow_tf32):
    """The CPU oracle with ONLY its conv / linear calls evaluated by cuDNN /
    cuBLAS on the GPU (everything else stays on the CPU): what the reference's
    own arithmetic gives when its convolutions run on a GPU, in strict fp32 or
    with torch's default TF32 convolutions."""
    import torch.nn.functional as F
    o_conv, o_lin = torch_ref._conv, torch_ref._linear
    sd_dev = {k: v.cuda() for k, v in sd.items() if v.is_floating_point()}

    def conv(sd_, p, x, stride=1, padding=0, dilation=1):
        w = sd_dev[p + '.weight']
        f = F.conv3d if w.dim() == 5 else F.conv2d
        return f(x.cuda(), w, sd_dev.get(p + '.bias'), stride, padding, dilation).cpu()

    def lin(sd_, p, x):
        return F.linear(x.cuda(), sd_dev[p + '.weight'], sd_dev.get(p + '.bias')).cpu()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch_ref._conv, torch_ref._linear = conv, lin
    try:
        st = {}
        with torch.no_grad():
            out = torch_ref.preworld_simple_test(sd, pc, inputs, st)
    finally:
        torch_ref._conv, torch_ref._linear = o_conv, o_lin
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out, st


@pytest.mark.timeout(900)
def test_reference_arithmetic_on_gpu_floor(golden_dir):
    """What 'bit-exact argmax against the CPU fixture' can mean on a GPU: the
    reference's own path with its convolutions evaluated by cuDNN (strict fp32,
    and torch's default TF32) differs from the CPU fixture too.  Reported next to
    this library's result (gpurun_out/argmax_floor.json -> profiles/)."""
    import json
    case = CASES['full_finetune']
    fx = np.load(os.path.join(golden_dir, 'full_finetune.npz'))
    want = fx['out/semantic_occ']
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    inputs, _ = build_case_inputs(case)
    pc = torch_ref.PathConfig(model_cfg_for(case))
    ost = {}
    with torch.no_grad():
        torch_ref.preworld_simple_test(sd, pc, inputs, ost)
    ref_logits = ost['logits']
    scale = ref_logits.abs().max().item()
    report = {}
    for name, tf32 in (('cudnn_fp32', False), ('cudnn_tf32_default', True)):
        out, st = _oracle_with_gpu_convs(sd, pc, inputs, tf32)
        report[name] = {
            'argmax_mismatch_vs_cpu_fixture': int((out['semantic_occ'][0] != want).sum()),
            'logits_max_rel_err_vs_cpu_oracle':
                (st['logits'] - ref_logits).abs().max().item() / scale}
    model = model.cuda()
    dev_inputs = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        _, st = _run_stages(model, dev_inputs)
        out = model.simple_test(None, None, img=dev_inputs)
    report['preworld_b200'] = {
        'argmax_mismatch_vs_cpu_fixture': int((out['semantic_occ'][0] != want).sum()),
        'logits_max_rel_err_vs_cpu_oracle':
            (st['logits'].cpu() - ref_logits).abs().max().item() / scale}
    print('argmax floor (640000 voxels):', json.dumps(report))
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/argmax_floor.json', 'w') as f:
        json.dump(report, f, indent=1)
    # this library must be at least as close to the CPU fixture as the
    # reference's own default GPU arithmetic (TF32 convolutions)
    assert report['preworld_b200']['argmax_mismatch_vs_cpu_fixture'] <= \
        report['cudnn_tf32_default']['argmax_mismatch_vs_cpu_fixture']


@pytest.mark.parametrize('name', ['tiny_finetune', 'tiny_pretrain'])
def test_cuda_graph_replay_equals_eager(name):
    """enable_cuda_graph(): the captured graph replays the same launches, so
    the occupancy grids are identical to the eager path -- also for a second
    sample fed through the same captured graph."""
    case = CASES[name]
    model = _model(case).cuda()
    samples = [build_case_inputs(case)[0],
               build_case_inputs(dict(case,
                                      input_seed=case['input_seed'] + 7))[0]]
    eager = []
    with torch.no_grad():
        for s in samples:
            out = model(return_loss=False,
                        img_inputs=[tuple(t.cuda() for t in s)], img_metas=[None])
            eager.append((out['semantic_occ'][0], out['geo_occ'][0]))
    model.enable_cuda_graph()
    with torch.no_grad():
        for rep in range(2):
            for s, want in zip(samples, eager):
                out = model(return_loss=False,
                            img_inputs=[tuple(t.cuda() for t in s)],
                            img_metas=[None])
                assert np.array_equal(out['semantic_occ'][0], want[0])
                assert np.array_equal(out['geo_occ'][0], want[1])
    assert len(model._graph_cache) == 1
    # host (pinned) tensors: image H2D straight into the static buffer, pose
    # chain on the CPU -- same occupancy
    with torch.no_grad():
        for s, want in zip(samples, eager):
            out = model(return_loss=False,
                        img_inputs=[tuple(t.pin_memory() for t in s)],
                        img_metas=[None])
            assert np.array_equal(out['semantic_occ'][0], want[0])


def test_cuda_graph_follows_weight_updates():
    """a captured graph bakes in the packed weights.  After
    load_state_dict() with other weights (or .to()), the public call must give
    the eager result for the NEW weights, not replay the old ones."""
    case = CASES['tiny_finetune']
    model = _model(case).cuda()
    s = build_case_inputs(case)[0]
    call = lambda: model(return_loss=False,
                         img_inputs=[tuple(t.cuda() for t in s)],
                         img_metas=[None])['semantic_occ'][0]
    model.enable_cuda_graph()
    with torch.no_grad():
        first = call()
        other = build_model(model_cfg_for(case)).eval()
        S.lively_init_(other, case['seed'] + 11)
        model.load_state_dict(other.state_dict())
        graphed = call()
        model.enable_cuda_graph(False)
        eager = call()
    assert not np.array_equal(first, eager)          # the weights do matter
    assert np.array_equal(graphed, eager)
    model.enable_cuda_graph()
    with torch.no_grad():
        call()
        model.float()                                # _apply drops the captures
        assert len(model._graph_cache) == 0
        assert np.array_equal(call(), eager)


@pytest.mark.timeout(1200)
def test_full_size_traj_matches_oracle():
    """BASELINE.json configs[3] at full size: 6x3x256x704 -> 7 occupancy grids of
    200x200x16 (current + 6 state-conditioned forecasting steps) against the CPU oracle
    (preworld_temporal_traj.py:213-371).  Every differing voxel must be a near tie of the
    oracle's logits of ITS step (recomputed from the oracle's own fused features); the fused
    voxel features of the last step stay within 1e-4 of the oracle's."""
    case = dict(CASES['full_finetune'], variant='finetune-traj', detector='PreWorld4DTraj',
                seed=5, input_seed=6)
    model = _model(case)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    inputs, extra = build_case_inputs(case)
    dev_inputs = tuple(t.cuda() for t in inputs)
    ego = extra['temporal_ego_states'][0][0]
    with torch.no_grad():
        out = model(return_loss=False, img_inputs=[dev_inputs], img_metas=[None], **extra)
        vf = model.voxel_features_cl(dev_inputs)
        bias = model.ego_bias(ego, vf.device)
        for _ in range(6):
            vf = model.forecast_step(vf, ego_bias=bias)
    pc = torch_ref.PathConfig(model_cfg_for(case))
    ost = {}
    with torch.no_grad():
        want = torch_ref.preworld4d_simple_test(sd, pc, inputs, [ego], ost)
    total = 0
    for k in range(7):
        key = f'semantic_occ_{k}s'
        occ, ref = out[key][0], want[key][0]
        assert occ.shape == (200, 200, 16)
        vf_o = ost['voxel_feats'] if k == 0 else ost[f'voxel_feats_{k}s']
        with torch.no_grad():
            logits_o = torch_ref.occ_from_head(sd, pc, vf_o, True)[2]
        n = _margin_ok(occ, ref, logits_o, tol=2e-5)
        total += n
        assert n <= 64, (key, n)
        assert ((out[f'geo_occ_{k}s'][0] == 0) == (occ != 17)).all()
    print(f'full-size traj: {total} near-tie voxels differ over 7 x 640 000')
    got_last = vf.permute(0, 3, 2, 1, 4).cpu()               # [B,Z,Y,X,C] -> [B,X,Y,Z,C]
    ref_last = ost['voxel_feats_6s']
    err = (got_last - ref_last).abs().max().item() / ref_last.abs().max().item()
    print(f'full-size traj: fused voxel features after 6 steps, max rel err {err:.2e}')
    assert err < 1e-4


@pytest.mark.timeout(1200)
def test_full_size_pretrain_render_matches_oracle():
    """BASELINE.json configs[2] at full size: 6x3x256x704 trunk -> attribute projection
    (density / semantic / colour of 640 000 voxels) -> volume rendering of 38 400 rays x 417
    samples.  The config-3-specific part is checked at full size against the CPU oracle
    (preworld.py:251-254, nerf_head.py:165-269,332-407) fed with the SAME voxel features
    (the trunk's own full-size parity is test_full_size_finetune_matches_reference): the
    attribute volumes within 1e-5, every rendering within the north-star's 1e-3 (median
    1e-5), the ray mask identical; then the loss reduction against the oracle's."""
    from preworld_b200 import model_cfg
    cfg = model_cfg('pretrain', 'r50', (256, 704))
    model = build_model(cfg).eval()
    S.lively_init_(model, 3)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    inputs = S.make_img_inputs(1, (256, 704), seed=4)
    rays = S.make_rays(inputs, 38400, seed=5)
    dev_inputs = tuple(t.cuda() for t in inputs)
    with torch.no_grad():
        vf = model.voxel_features_cl(dev_inputs)
        attr = model.attributes_cl(vf)
        res = model.render_forward(dev_inputs, rays.cuda())[0]
        losses = model.nerf_head.compute_loss(res, rays[0].cuda())
    vf_ref = vf.permute(0, 3, 2, 1, 4).cpu()                 # [B,X,Y,Z,C]
    with torch.no_grad():
        d, s, c = torch_ref.attribute_projection(sd, vf_ref)
    a = attr.permute(0, 3, 2, 1, 4).cpu()
    for got, want in ((a[..., 0], d), (a[..., 2:19], s), (a[..., 19:22], c)):
        assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    ng = torch_ref.NerfGeometry([-40., -40., -1., 40., 40., 5.4])
    mask = res['ray_mask'].cpu()
    want_mask = (rays[0, :, 2] > 0) & (rays[0, :, 2] <= 52)
    assert torch.equal(mask, want_mask) and int(mask.sum()) > 30000
    outs = {k: [] for k in ('render_depth', 'render_semantic', 'render_color', 'alphainv_last')}
    for lo in range(0, 38400, 4800):                          # bounded oracle memory
        o = torch_ref.render_rays(ng, rays[0, lo:lo + 4800], inputs[6][0], d[0], s[0], c[0])
        for k in outs:
            outs[k].append(o[k])
    ref = {k: torch.cat(v) for k, v in outs.items()}
    for k, want in ref.items():
        got = res[k].cpu()[mask]
        scale = max(1.0, want.abs().max().item())
        err = (got - want).abs().max().item() / scale
        med = (got - want).abs().median().item() / scale
        print(f'full-size render {k}: max {err:.2e} median {med:.2e}')
        assert err < 1e-3 and med < 1e-5, (k, err, med)
    want_l = torch_ref.nerf_compute_loss(ref, rays[0, mask, 2], rays[0, mask, 3],
                                         rays[0, mask, 13:16], model.nerf_head.class_weights)
    for k, v in want_l.items():
        assert abs(float(losses[k]) - float(v)) <= 1e-4 * abs(float(v)), (k, losses[k], v)

"""Model dicts for the camera->voxel path, in the reference's config schema.

The reference selects every module through mmcv config dicts
(configs/preworld/nuscenes/{bevstereo-occ,preworld-7frame-finetune,
preworld-7frame-pretrain}.py and configs/preworld/nuscenes-temporal/*-traj.py).
Those files load unchanged through ``preworld_b200.config.Config.fromfile``;
this module re-states only their ``model = dict(...)`` part so that the GPU box
(which has no /root/reference) can build the same models.
``tests/test_configs.py`` checks, when the reference tree is present, that
``model_cfg(variant, backbone='swin')`` equals the reference file's model dict.

The ResNet-50/101 @ 256x704 variant named by BASELINE.json is a *derived*
config: the reference ships only Swin-B @ 512x1408 (bevstereo-occ.py:16,45-67)
although its code supports mmdet ResNet + CustomFPN (detectors/bevdet.py:
577-588, necks/fpn.py:10-11).
"""
import copy

VARIANTS = ('finetune', 'pretrain', 'finetune-traj', 'pretrain-traj')

PC_RANGE = [-40., -40., -1., 40., 40., 5.4]


def grid_config(x=(-40, 40, 0.4), y=(-40, 40, 0.4), z=(-1, 5.4, 0.4),
                depth=(1.0, 45.0, 0.5)):
    return {'x': list(x), 'y': list(y), 'z': list(z), 'depth': list(depth)}


def _swin_image_side():
    backbone = dict(
        type='SwinTransformer', pretrain_img_size=224, patch_size=4,
        window_size=12, mlp_ratio=4, embed_dims=128, depths=[2, 2, 18, 2],
        num_heads=[4, 8, 16, 32], strides=(4, 2, 2, 2), out_indices=(2, 3),
        qkv_bias=True, qk_scale=None, patch_norm=True, drop_rate=0.,
        attn_drop_rate=0., drop_path_rate=0.1, use_abs_pos_embed=False,
        return_stereo_feat=True, act_cfg=dict(type='GELU'),
        norm_cfg=dict(type='LN', requires_grad=True),
        pretrain_style='official', output_missing_index_as_none=False)
    neck = dict(type='FPN_LSS', in_channels=512 + 1024, out_channels=512,
                extra_upsample=None, input_feature_index=(0, 1),
                scale_factor=2)
    return backbone, neck, 512, (512, 1408)


def _resnet_image_side(depth, input_size):
    backbone = dict(
        type='ResNet', depth=depth, num_stages=4, out_indices=(0, 2, 3),
        frozen_stages=-1, norm_cfg=dict(type='BN', requires_grad=True),
        norm_eval=False, with_cp=False, style='pytorch')
    neck = dict(type='CustomFPN', in_channels=[1024, 2048], out_channels=256,
                num_outs=1, start_level=0, out_ids=[0])
    return backbone, neck, 256, tuple(input_size)


def model_cfg(variant='finetune', backbone='r50', input_size=(256, 704),
              grid=None, num_trans=32):
    """The ``model`` dict of configs/preworld/**/preworld-7frame-<variant>.py
    with the requested image side ('swin' == as shipped; 'r50'/'r101' ==
    derived)."""
    assert variant in VARIANTS, variant
    grid = grid or grid_config()
    if backbone == 'swin':
        bb, neck, vt_in, input_size = _swin_image_side()
    else:
        bb, neck, vt_in, input_size = _resnet_image_side(
            int(backbone[1:]), input_size)
    adj = (1, 2, 1)                      # multi_adj_frame_id_cfg
    n_adj = len(range(*adj))
    pretrain = variant.startswith('pretrain')
    traj = variant.endswith('traj')
    render_w = 1.0 if pretrain else 0.0
    cfg = dict(
        type='PreWorld4DTraj' if traj else 'PreWorld',
        align_after_view_transfromation=False,
        num_adj=n_adj,
        img_backbone=bb,
        img_neck=neck,
        img_view_transformer=dict(
            type='LSSViewTransformerBEVStereo', grid_config=grid,
            input_size=input_size, in_channels=vt_in, out_channels=num_trans,
            sid=False, collapse_z=False, loss_depth_weight=0.05,
            depthnet_cfg=dict(use_dcn=False, aspp_mid_channels=96,
                              stereo=True, bias=5.),
            downsample=16),
        img_bev_encoder_backbone=dict(
            type='CustomResNet3D', numC_input=num_trans * (n_adj + 1),
            num_layer=[1, 2, 4], with_cp=False,
            num_channels=[num_trans, num_trans * 2, num_trans * 4],
            stride=[1, 2, 2], backbone_output_ids=[0, 1, 2]),
        img_bev_encoder_neck=dict(type='LSSFPN3D', in_channels=num_trans * 7,
                                  out_channels=num_trans),
        pre_process=dict(
            type='CustomResNet3D', numC_input=num_trans, with_cp=False,
            num_layer=[1, ], num_channels=[num_trans, ], stride=[1, ],
            backbone_output_ids=[0, ]),
        loss_occ=dict(type='CrossEntropyLoss', use_sigmoid=False,
                      loss_weight=1.0),
        final_softplus=True,
        use_lss_depth_loss=pretrain and not traj,
        use_3d_loss=False,
        if_render=pretrain,
        if_post_finetune=not pretrain,
        weight_voxel_ce=0.0 if pretrain else 1.0,
        weight_voxel_sem_scal=0.0 if pretrain else 1.0,
        weight_voxel_geo_scal=0.0 if pretrain else 1.0,
        weight_voxel_lovasz=0.0 if pretrain else 1.0,
        empty_idx=17,
        nerf_head=dict(
            type='NerfHead', point_cloud_range=list(PC_RANGE), voxel_size=0.4,
            scene_center=[0, 0, 2.2], radius=39, use_depth_sup=pretrain,
            weight_depth=render_w, weight_semantic=render_w,
            weight_color=render_w, weight_entropy_last=0.01,
            weight_distortion=0.01),
        occupancy_head=dict(
            type='OccHead', with_cp=False, use_deblock=False,
            norm_cfg=dict(type='SyncBN', requires_grad=True),
            soft_weights=True, final_occ_size=[200, 200, 16], empty_idx=17,
            num_level=1, in_channels=[32], out_channel=18,
            point_cloud_range=list(PC_RANGE)),
    )
    return copy.deepcopy(cfg)

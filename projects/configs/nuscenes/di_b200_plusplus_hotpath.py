# `imgpts_neck` / `pts_bbox_head` / `test_cfg` sections of the DeepInteraction++ model: the hyper-parameters the
# reference's projects/configs/nuscenes/Fusion_0075_plusplus.py:210-305 gives to FusionTransformerv4 and
# DeepInteractionPlusPlusDecoder.  The reference's own config
# file loads unchanged through projects.mmdet3d_plugin.registry.load_config / build_neck (tests/test_host_cpu.py); this
# file exists because the benchmark box has no copy of the reference tree.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'
hidden = 128
point_cloud_range = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
class_names = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle', 'bicycle', 'pedestrian',
               'traffic_cone']
voxel_size = [0.075, 0.075, 0.2]
out_size_factor = 8
num_views = 6
_msda = dict(type='MultiScaleDeformableAttention', embed_dims=hidden, num_levels=2, batch_first=True)
_ffn = dict(type='FFN', embed_dims=hidden, feedforward_channels=512, num_fcs=2, ffn_drop=0.1,
            act_cfg=dict(type='ReLU', inplace=True))

model = dict(
    type='DeepInteraction',
    imgpts_neck=dict(
        type='FusionTransformerv4', num_layers=2, in_channels_img=256, in_channels_pts=256, hidden_channel=hidden,
        bn_momentum=0.1, bias='auto',
        img_transformerlayers=dict(
            type='DeepInteractionLayer',
            attn_cfgs=[_msda, dict(type='MMRI_P2I', embed_dims=hidden, batch_first=True)],
            ffn_cfgs=_ffn,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm', 'ffn', 'norm')),
        pts_transformerlayers=dict(
            type='DeepInteractionLayer',
            attn_cfgs=[_msda, dict(type='MMRI_I2P_Polar', embed_dims=hidden, dropout=0.1, batch_first=True),
                       dict(type='MMRI_I2P', embed_dims=hidden, dropout=0.1, batch_first=True, fp16_enabled=True,
                            group_attn_enabled=True)],
            ffn_cfgs=_ffn,
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))),
    pts_bbox_head=dict(
        type='DeepInteractionPlusPlusDecoder', num_views=num_views, out_size_factor_img=4, num_proposals=200,
        auxiliary=True, hidden_channel=hidden, num_classes=len(class_names), num_mmpi=4, num_heads=8,
        learnable_query_pos=False, initialize_by_heatmap=True, nms_kernel_size=3, ffn_channel=256, dropout=0.1,
        bn_momentum=0.1, activation='relu',
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
        bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=point_cloud_range[:2], voxel_size=voxel_size[:2],
                        out_size_factor=out_size_factor, post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                        score_threshold=0.0, code_size=10),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0)),
    test_cfg=dict(pts=dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=out_size_factor,
                           pc_range=point_cloud_range[0:2], voxel_size=voxel_size[:2], nms_type=None)))

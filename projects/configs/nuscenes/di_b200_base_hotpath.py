# Hot-path section of the DeepInteraction-base model (the hyper-parameters the reference's
# projects/configs/nuscenes/Fusion_0075_refactor.py:185-251 gives to `imgpts_neck`, `pts_bbox_head` and
# `test_cfg.pts`).  The reference's own config file loads unchanged through
# projects.mmdet3d_plugin.registry.load_config; this file exists because the benchmark box has no copy of
# the reference tree.  Only the two hot-path modules are described -- the detector, backbones and data
# pipeline stay with the reference.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'

point_cloud_range = [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0]
class_names = ['car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'barrier', 'motorcycle', 'bicycle',
               'pedestrian', 'traffic_cone']
voxel_size = [0.075, 0.075, 0.2]
out_size_factor = 8
num_views = 6
hidden = 128

model = dict(
    type='DeepInteraction',
    imgpts_neck=dict(type='DeepInteractionEncoder', num_layers=2, in_channels_img=256, in_channels_pts=512,
                     hidden_channel=hidden, bn_momentum=0.1, bias='auto'),
    pts_bbox_head=dict(
        type='DeepInteractionDecoder', num_views=num_views, out_size_factor_img=4, num_proposals=200, auxiliary=True,
        hidden_channel=hidden, num_classes=len(class_names), num_mmpi=4, num_heads=8, learnable_query_pos=False,
        initialize_by_heatmap=True, nms_kernel_size=3, ffn_channel=256, dropout=0.1, bn_momentum=0.1, activation='relu',
        common_heads=dict(center=(2, 2), height=(1, 2), dim=(3, 2), rot=(2, 2), vel=(2, 2)),
        bbox_coder=dict(type='TransFusionBBoxCoder', pc_range=point_cloud_range[:2], voxel_size=voxel_size[:2],
                        out_size_factor=out_size_factor, post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
                        score_threshold=0.0, code_size=10),
        loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2, alpha=0.25, reduction='mean', loss_weight=1.0),
        loss_bbox=dict(type='L1Loss', reduction='mean', loss_weight=0.25),
        loss_heatmap=dict(type='GaussianFocalLoss', reduction='mean', loss_weight=1.0)),
    test_cfg=dict(pts=dict(dataset='nuScenes', grid_size=[1440, 1440, 40], out_size_factor=out_size_factor,
                           pc_range=point_cloud_range[0:2], voxel_size=voxel_size[:2], nms_type=None)))

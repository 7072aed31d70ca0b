# `imgpts_neck` section of the DeepInteraction++ model: the hyper-parameters the reference's
# projects/configs/nuscenes/Fusion_0075_plusplus.py:210-271 gives to FusionTransformerv4.  The reference's own config
# file loads unchanged through projects.mmdet3d_plugin.registry.load_config / build_neck (tests/test_host_cpu.py); this
# file exists because the benchmark box has no copy of the reference tree.
plugin = True
plugin_dir = 'projects/mmdet3d_plugin/'
hidden = 128
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
            operation_order=('self_attn', 'norm', 'cross_attn', 'norm', 'cross_attn', 'norm', 'ffn', 'norm'))))

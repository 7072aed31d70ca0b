"""Registries for the plug-in classes.

With mmcv/mmdet/mmdet3d installed (the reference's environment, install.md) the classes are registered
into the REAL registries, so ``Config.fromfile`` + ``build_model`` pick them up through
``plugin_dir='projects/mmdet3d_plugin/'`` exactly like the reference (tools/train.py:106-127).
Without them (this repository's offline image) a small local registry with the same
``register_module()`` / ``build(cfg)`` surface is used, and ``load_config`` executes the reference's
plain-Python config files.
"""


class Registry:
    def __init__(self, name):
        self.name, self.module_dict = name, {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            key = name or cls.__name__
            if key in self.module_dict and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self.module_dict[key] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        cfg = dict(cfg)
        for k, v in default_args.items():
            cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {self.name} registry')
        return cls(**cfg)


def _real(path, attr):
    try:
        mod = __import__(path, fromlist=[attr])
        return getattr(mod, attr)
    except Exception:
        return None


HAVE_MMDET3D = _real('mmdet3d.models.builder', 'NECKS') is not None
NECKS = _real('mmdet3d.models.builder', 'NECKS') or Registry('neck')
HEADS = _real('mmdet3d.models.builder', 'HEADS') or Registry('head')
BBOX_CODERS = _real('mmdet.core.bbox.builder', 'BBOX_CODERS') or Registry('bbox_coder')
TRANSFORMER_LAYER = _real('mmcv.cnn.bricks.registry', 'TRANSFORMER_LAYER') or Registry('transformer layer')
ATTENTION = _real('mmcv.cnn.bricks.registry', 'ATTENTION') or Registry('attention')
BBOX_ASSIGNERS = _real('mmdet.core.bbox.builder', 'BBOX_ASSIGNERS') or Registry('bbox_assigner')
MATCH_COST = _real('mmdet.core.bbox.match_costs.builder', 'MATCH_COST') or Registry('match cost')


def load_config(path):
    """Execute an mmcv-style plain-Python config (no ``_base_``) and return its variables as a dict."""
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), ns)
    return {k: v for k, v in ns.items() if not k.startswith('__')}


def build_neck(cfg):
    """Build cfg.model.imgpts_neck alone."""
    model = cfg['model'] if 'model' in cfg else cfg
    return NECKS.build(model['imgpts_neck'])


def build_hot_path(cfg):
    """Build (imgpts_neck, pts_bbox_head) the way mmdet3d's MVXTwoStageDetector does: the head receives
    ``train_cfg=train_cfg.pts`` / ``test_cfg=test_cfg.pts`` (reference detectors/deepinteraction.py:19-58)."""
    model = cfg['model'] if 'model' in cfg else cfg
    neck = NECKS.build(model['imgpts_neck'])
    head_cfg = dict(model['pts_bbox_head'])
    test_cfg = model.get('test_cfg') or {}
    # folded into the cfg like MVXTwoStageDetector does (mmcv's Registry.build takes the cfg only)
    train_cfg = model.get('train_cfg') or {}
    head_cfg.update(train_cfg=train_cfg.get('pts'), test_cfg=test_cfg.get('pts'))
    head = HEADS.build(head_cfg)
    return neck, head

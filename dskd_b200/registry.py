"""LOSSES registry + build_loss, the reference's drop-in convention.

Reference: mmdet/models/builder.py:14 (`LOSSES = MODELS`), :43-45 (`build_loss`), used as
`@LOSSES.register_module()` (mse_loss.py:15, kd_loss.py:46) and `build_loss(dict(type=..., **kw))`
(gfl_deformable_detr_head_il.py:133-143).  mmcv (which owns the real Registry) is absent in this
image, so a minimal compatible registry lives here; when mmdet *is* importable the modules are also
registered into its registry so existing configs resolve `type='DSGFeatureDistillLoss'` unchanged.
"""
import inspect


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def __contains__(self, key):
        return key in self._modules

    def get(self, key):
        return self._modules.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f'{key} is already registered in {self.name}')
            self._modules[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise TypeError(f'cfg must be a dict with a "type" key, got {cfg!r}')
        args = dict(cfg)
        kind = args.pop('type')
        cls = self.get(kind) if isinstance(kind, str) else kind
        if cls is None:
            raise KeyError(f'{kind} is not in the {self.name} registry')
        if not inspect.isclass(cls):
            raise TypeError(f'type must be a str or a class, got {type(cls)}')
        return cls(**args)


LOSSES = Registry('loss')
ASSIGNERS = Registry('bbox_assigner')


def build_loss(cfg):
    """builder.py:43-45."""
    return LOSSES.build(cfg)


def build_assigner(cfg):
    return ASSIGNERS.build(cfg)


def register_into_mmdet():
    """Make the modules visible to an installed mmdet (no-op when mmdet / mmcv are absent)."""
    try:
        from mmdet.models.builder import LOSSES as MM_LOSSES  # type: ignore
    except Exception:
        return False
    for key, cls in LOSSES._modules.items():
        if key in ('MSELoss', 'KnowledgeDistillationKLDivLoss'):
            continue                      # never shadow mmdet's own modules of the same name
        MM_LOSSES.register_module(name=key, force=True, module=cls)
    return True

"""Import the UNMODIFIED reference (/root/reference, read-only) in this container.

The reference is an MMDetection 2.23 fork that hard-requires mmcv-full (absent, no network).
On the distillation hot path mmcv only supplies decorators (`force_fp32`, `mmcv.jit`),
registries and base classes, so a permissive stub package is enough to import and *run* the
reference's own `GFLDeformableDETRHead_il.loss`, `GFLHungarianAssigner.assign`, match costs and
loss modules on CPU.  Used only by gen_golden.py to produce tests/golden/*.npz; never imported
by tests, bench or the product (the GPU box has no /root/reference).
"""
import importlib.abc
import importlib.machinery
import sys
import types

import torch.nn as nn

REFERENCE_ROOT = '/root/reference'


class _Stub:
    """Usable as decorator, decorator factory, base class and attribute bag."""

    def __init__(self, name='stub'):
        self.__name__ = name
        self.__qualname__ = name

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and callable(a[0]) and not isinstance(a[0], _Stub):
            return a[0]                      # @decorator  /  @factory(...)(fn)
        return self

    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'):
            raise AttributeError(n)
        return _Stub(n)

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        has_module = any(isinstance(b, type) and issubclass(b, nn.Module) for b in bases)
        first = [b for b in bases if isinstance(b, _Stub)][0]
        base = object if (has_module or first is not self) else nn.Module
        return (type('Stub_' + self.__name__, (base,), {}),)


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, n):
        if n.startswith('__') and n.endswith('__'):
            raise AttributeError(n)
        return _Stub(n)


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ('mmcv', 'pycocotools', 'terminaltables', 'matplotlib', 'cv2')

    def find_spec(self, fullname, path, target=None):
        if fullname.split('.')[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install():
    if not any(isinstance(f, _Finder) for f in sys.meta_path):
        sys.meta_path.insert(0, _Finder())
        sys.path.insert(0, REFERENCE_ROOT)
        sys.dont_write_bytecode = True
    import mmcv
    mmcv.__version__ = '1.5.0'
    mmcv.digit_version = lambda v: tuple(int(x) for x in v.split('.')[:3])
    import mmdet  # noqa: F401
    return mmdet

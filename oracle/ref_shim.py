"""
ORACLE (test infrastructure, NOT product code) -- import the reference's OWN `EcgVit` verbatim.

Works only where `/root/reference` exists (the build container; never on the GPU box).  The reference
imports 11 third-party modules that are not installed offline (sty, colorama, matplotlib, seaborn,
h5py, wfdb, pytorch_lightning, icecream, loess, pywt, vit_pytorch).  Ten of them are plotting / IO /
logging helpers irrelevant to the hot path and are replaced by permissive stubs; `vit_pytorch` (the
arithmetic) is replaced by the restatement in `oracle/vit_restated.py`.

Used by `tests/golden/make_golden.py` to generate golden vectors and by the CPU tests that validate the
travelling oracle (`oracle/ecg_vit_oracle.py`) against the reference wrapper.
"""
import os
import sys
import types
import importlib

REFERENCE_ROOT = '/root/reference'

_STUB_TOP = ['sty', 'colorama', 'matplotlib', 'seaborn', 'h5py', 'wfdb', 'pytorch_lightning', 'icecream',
             'loess', 'pywt']


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Stub

    def __add__(cls, other):
        return other if isinstance(other, str) else cls

    __radd__ = __add__

    def __getitem__(cls, item):
        return _Stub

    def __str__(cls):
        return ''

    def __format__(cls, spec):
        return ''


class _Stub(metaclass=_StubMeta):
    """Class that can be subclassed, called, indexed and attribute-walked without doing anything."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Stub()

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _Stub()

    def __getitem__(self, item):
        return _Stub()

    def __setitem__(self, key, value):
        pass

    def __iter__(self):
        return iter(())

    def __add__(self, other):
        return other if isinstance(other, str) else self

    __radd__ = __add__

    def __str__(self):
        return ''


class _StubModule(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        self.__path__ = []  # behaves as a package so `import a.b.c` resolves
        self.__all__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name == 'rcParams':
            d = {}
            setattr(self, name, d)
            return d
        return _Stub


class _StubFinder:
    """meta-path finder that serves stub modules for the missing top-level packages."""

    def find_spec(self, fullname, path=None, target=None):
        top = fullname.split('.')[0]
        if top in _STUB_TOP:
            from importlib.machinery import ModuleSpec
            return ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Make `import ecg_transformer...` work from /root/reference; idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f'{REFERENCE_ROOT} is not present: the reference shim only works in the build container')
    sys.meta_path.append(_StubFinder())
    # the arithmetic: restated vit_pytorch
    from oracle import vit_restated
    vp = types.ModuleType('vit_pytorch')
    vp.__path__ = []
    vp.ViT = vit_restated.ViT
    vp_vit = types.ModuleType('vit_pytorch.vit')
    for n in ('ViT', 'Transformer', 'Attention', 'FeedForward', 'PreNorm'):
        setattr(vp_vit, n, getattr(vit_restated, n))
    vp_rec = types.ModuleType('vit_pytorch.recorder')
    vp_rec.Recorder = vit_restated.Recorder
    vp.vit, vp.recorder = vp_vit, vp_rec
    sys.modules['vit_pytorch'] = vp
    sys.modules['vit_pytorch.vit'] = vp_vit
    sys.modules['vit_pytorch.recorder'] = vp_rec
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'ecg_transformer'))


def load_reference():
    """Returns the reference's (EcgVit, EcgVitConfig, get_train_args, ModelOutput) objects, verbatim."""
    install()
    ecg_vit = importlib.import_module('ecg_transformer.models.ecg_vit')
    train = importlib.import_module('ecg_transformer.models.train')
    models_util = importlib.import_module('ecg_transformer.util.models')
    return ecg_vit.EcgVit, ecg_vit.EcgVitConfig, train.get_train_args, models_util.ModelOutput

"""Import shim that lets the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py to produce the fixtures under
tests/golden/; never imported by robovat_b200 and never needed on the GPU box (the
reference tree does not travel).  Follows SURVEY.md Appendix D: stub the third-party
modules the reference imports at module scope but that are not installed here (gym,
pybullet, matplotlib, easydict, h5py, pcl, scipy.misc) and patch two numpy-2
incompatibilities of third_party/transformations.py (`numpy.array(..., copy=False)`).
Nothing in the reference's own files is modified.
"""
import os
import sys
import types

import numpy as np

REFERENCE = os.environ.get('ROBOVAT_REFERENCE', '/root/reference')


class _Box(object):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is not None:
            low, high = np.full(shape, low), np.full(shape, high)
        self.low, self.high, self.dtype = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype), dtype
        self.shape = self.low.shape


class _Discrete(object):
    def __init__(self, n):
        self.n = n


class _Dict(object):
    def __init__(self, spaces):
        self.spaces = dict(spaces)


class _AttrDict(dict):
    def __init__(self, d=None, **kw):
        super(_AttrDict, self).__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        super(_AttrDict, self).__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    __setattr__ = __setitem__


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install(pybullet_stub=None):
    """Make `import robovat...` work.  Idempotent."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE):
        raise RuntimeError('%s is not present: the reference only exists in the authoring container' % REFERENCE)
    sys.dont_write_bytecode = True
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    spaces = _module('gym.spaces', Box=_Box, Discrete=_Discrete, Dict=_Dict)
    _module('gym', spaces=spaces, Env=object)
    plt = _module('matplotlib.pyplot', figure=lambda *a, **k: None, ion=lambda: None, show=lambda: None,
                  pause=lambda *a: None)
    _module('matplotlib', pyplot=plt)
    _module('easydict', EasyDict=_AttrDict)
    _module('h5py')
    pb = pybullet_stub or types.SimpleNamespace()
    _module('pybullet', **{k: getattr(pb, k) for k in dir(pb) if not k.startswith('_')})
    # scipy.misc lost imresize & co; perception/image_utils imports them at module scope
    import scipy
    if not hasattr(scipy, 'misc') or not hasattr(getattr(scipy, 'misc', None), 'imresize'):
        misc = _module('scipy.misc', imresize=None, imrotate=None, imread=None, imsave=None)
        scipy.misc = misc
    import third_party.transformations as T
    np_proxy = types.ModuleType('numpy_proxy')
    np_proxy.__dict__.update(np.__dict__)

    def array(obj, *args, **kwargs):
        if kwargs.get('copy') is False:
            kwargs.pop('copy')
            return np.asarray(obj, *args, **kwargs)
        return np.array(obj, *args, **kwargs)
    np_proxy.array = array
    T.numpy = np_proxy
    if not hasattr(np, 'float'):
        np.float = float                 # robovat/math/pose.py:206 uses the removed alias
    _installed = True


def patch_point_cloud_utils():
    """`-1 * np.ones_like(uint8 segmask)` (perception/point_cloud_utils.py:123) relied on numpy-1 value-based
    promotion to int16; under NEP 50 it overflows.  Give that module a numpy whose ones_like widens uint8."""
    import robovat.perception.point_cloud_utils as pc
    proxy = types.ModuleType('numpy_proxy_pc')
    proxy.__dict__.update(np.__dict__)

    def ones_like(a, *args, **kwargs):
        a = np.asarray(a)
        if a.dtype == np.uint8 and 'dtype' not in kwargs:
            kwargs['dtype'] = np.int16
        return np.ones_like(a, *args, **kwargs)
    proxy.ones_like = ones_like
    pc.np = proxy


def attrdict(d):
    return _AttrDict(d)

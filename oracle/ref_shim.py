"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Container-only harness that imports the *unmodified* reference
(`/root/reference/src`) on CPU so that (a) the restated oracle in
`oracle/sgsac_oracle.py` can be pinned against it and (b) golden vectors can
be generated (`oracle/make_golden.py`).  `/root/reference` does not exist on
the GPU box; nothing under tests -m gpu, smoke() or bench.py may import this.

Shim list:
  * stub modules for absent GUI / sim deps imported at module scope
    (turtle: modules.py:1; pygame, PyQt5, pyqtgraph: utils.py:9,343-345,384-385;
    kornia: augmentations.py:4; captum: rl_utils.py:4)
  * captum==0.5.0 GuidedBackprop re-stated (un-vendored third-party, pinned in
    setup/sgqn-carla.yml:12): every nn.ReLU back-propagates relu(grad_in).
  * kornia==0.6.6 RandomCrop re-stated for random_shift (augmentations.py:229-233)
    as an integer crop at host-supplied offsets.
  * numpy>=2 fix of ReplayBuffer._encode_obses (utils.py:129-135).
  * .cuda() neutralised (CPU run).
  * all four RNG streams (numpy idxs utils.py:127, python random sgsac.py:68 /
    augmentations.py:70, torch randn_like modules.py:219, torch crop offsets
    augmentations.py:255-256) replaced by host-supplied values (`Tape`).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


class Tape:
    """Host-supplied randomness consumed by the reference in call order."""

    def __init__(self):
        self.idxs = []      # list of int arrays  -> np.random.randint
        self.noise = []     # list of (B,A) float tensors -> torch.randn_like
        self.u = []         # list of floats in [0,1) -> random.uniform(lo,hi)
        self.overlay = []   # list of int arrays -> sample_frames_from_carla_dataset
        self.crop = []      # list of (w1,h1) long tensors -> random_crop / random_shift
        self.pool = None    # uint8 (N,3,84,84) overlay pool
        self.places = []    # list of float (B,3,84,84) in [0,1] -> _get_places_batch


TAPE = Tape()
_loaded = {}


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Q:
    def __init__(self, *a, **k):
        pass


class GuidedBackprop:
    """captum==0.5.0 `GuidedBackprop.attribute` semantics used at rl_utils.py:35-39."""

    def __init__(self, model):
        self.model = model

    def attribute(self, inputs):
        hooks = []

        def pre(mod, inp):
            x = inp[0].clone()
            if x.requires_grad:
                x.register_hook(lambda g: F.relu(g))
            return x

        for m in self.model.modules():
            if isinstance(m, torch.nn.ReLU):
                hooks.append(m.register_forward_pre_hook(pre))
        had = inputs.requires_grad
        inputs.requires_grad_()
        try:
            with torch.enable_grad():
                out = self.model(inputs)
                assert out[0].numel() == 1
                (g,) = torch.autograd.grad(torch.unbind(out), inputs)
        finally:
            for h in hooks:
                h.remove()
        if not had:
            inputs.requires_grad_(False)
        return g


class _RandomCrop:
    """kornia==0.6.6 RandomCrop((h,w)) on an already padded batch: integer crop at
    per-sample offsets; offsets come from TAPE.crop as (dy, dx)."""

    def __init__(self, size):
        self.size = size

    def __call__(self, x):
        h, w = self.size
        dy, dx = TAPE.crop.pop(0)
        out = torch.empty(x.shape[0], x.shape[1], h, w, dtype=x.dtype)
        for b in range(x.shape[0]):
            out[b] = x[b, :, int(dy[b]):int(dy[b]) + h, int(dx[b]):int(dx[b]) + w]
        return out


def load():
    """Import the reference with shims; returns namespace dict of its modules."""
    if _loaded:
        return _loaded
    assert available(), "reference not mounted"
    _stub("turtle", forward=None)
    _stub("pygame")
    _stub("pyqtgraph", PlotWidget=_Q, plot=None)
    qw = _stub("PyQt5.QtWidgets", QMainWindow=_Q, QLabel=_Q, QVBoxLayout=_Q, QWidget=_Q, QApplication=_Q)
    qc = _stub("PyQt5.QtCore", Qt=_Q)
    _stub("PyQt5", QtCore=qc, QtWidgets=qw)
    ka = _stub("kornia.augmentation", RandomCrop=_RandomCrop, RandomAffine=None, RandomErasing=None)
    _stub("kornia", augmentation=ka)
    _stub("captum", attr=_stub("captum.attr", GuidedBackprop=GuidedBackprop, GuidedGradCam=None))
    if "termcolor" not in sys.modules:
        try:
            import termcolor  # noqa: F401
        except ImportError:
            _stub("termcolor", colored=lambda s, *a, **k: s)

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF_SRC)
    import utils
    import augmentations
    import arguments
    import algorithms.modules as modules
    import algorithms.sgsac as sgsac
    from algorithms.factory import make_agent

    def _enc(self, idxs):
        o, n = zip(*[(np.asarray(self._obses[i][0]), np.asarray(self._obses[i][1])) for i in idxs])
        return np.array(o), np.array(n)

    utils.ReplayBuffer._encode_obses = _enc
    utils.ReplayBuffer._get_idxs = lambda self, n=None: np.asarray(TAPE.idxs.pop(0))

    def _frames(x):
        ids = torch.as_tensor(np.asarray(TAPE.overlay.pop(0)), dtype=torch.long)
        return TAPE.pool[ids].repeat(1, 3, 1, 1)

    augmentations.sample_frames_from_carla_dataset = _frames
    augmentations.places_dataloader = object()
    augmentations._get_places_batch = lambda batch_size: TAPE.places.pop(0)

    _orig_crop = augmentations.random_crop

    def _crop(x, size=84, w1=None, h1=None, return_w1_h1=False):
        # bypass only the `x.is_cuda` assert (augmentations.py:241); same slicing
        if x.shape[-1] - size <= 0:
            return (x, None, None) if return_w1_h1 else x
        if w1 is None:
            w1, h1 = TAPE.crop.pop(0)
        n = x.shape[0]
        xp = x.permute(0, 2, 3, 1)
        windows = augmentations.view_as_windows_cuda(xp, (1, size, size, 1))[..., 0, :, :, 0]
        cropped = windows[torch.arange(n), w1, h1]
        return (cropped, w1, h1) if return_w1_h1 else cropped

    augmentations.random_crop = _crop
    modules.torch = _TorchProxy(torch)          # randn_like -> tape (modules.py:219)
    sgsac.random = _RandomProxy()               # random.uniform -> tape (sgsac.py:68)
    _loaded.update(utils=utils, augmentations=augmentations, arguments=arguments, modules=modules,
                   sgsac=sgsac, make_agent=make_agent)
    return _loaded


class _TorchProxy:
    def __init__(self, t):
        self._t = t

    def __getattr__(self, k):
        return getattr(self._t, k)

    def randn_like(self, x):
        n = TAPE.noise.pop(0)
        assert n.shape == x.shape
        return n.clone()


class _RandomProxy:
    def uniform(self, a, b):
        return a + (b - a) * TAPE.u.pop(0)     # CPython: a + (b-a)*random()


def parse_args(argv):
    ns = load()
    old = sys.argv
    sys.argv = ["x"] + list(argv)
    try:
        return ns["arguments"].parse_args()
    finally:
        sys.argv = old


class NullLogger:
    def __init__(self):
        self.rows = []

    def log(self, key, value, step, n=1):
        if isinstance(value, torch.Tensor):
            value = value.item()
        self.rows.append((key, float(value), step))

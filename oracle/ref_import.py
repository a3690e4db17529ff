"""Import the reference's OWN hot-path modules, unchanged, from its checkout (ORACLE tooling).

Only usable where the checkout exists (this container: /root/reference; never on the GPU box).
Used by tests/golden/make_golden.py to freeze golden vectors and by
tests/test_oracle_golden.py to pin the restatement against the live reference.

The reference imports four packages that are not installed here; they are replaced by the
smallest stubs that let `core/utils/hyperbolic.py`, `core/active/floating_region.py` and
`core/active/build.py` import and run on CPU:
  geoopt.manifolds.stereographic.math -> oracle.geoopt_math (restatement; hyperbolic.py:8)
  matplotlib.pyplot, mpl_toolkits.axes_grid1 -> empty modules (hyperbolic.py:10, build.py:16-17)
  yacs.config.CfgNode -> attribute dict (core/configs/defaults.py:3)
and `torch.Tensor.cuda` is neutralised while a reference function runs, because
floating_region.py:87,183-198 hard-code `.cuda()`.
"""
import contextlib
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("HALO_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "utils", "hyperbolic.py"))


class _CfgNode(dict):
    """Minimal yacs.config.CfgNode: attribute access on a dict, plus the no-op mutators the repo calls."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy

        return copy.deepcopy(self)

    def set_new_allowed(self, *_):
        pass

    def merge_from_file(self, *_):
        pass

    def merge_from_list(self, *_):
        pass

    def freeze(self):
        pass

    def defrost(self):
        pass


def _install_stubs():
    from . import geoopt_math

    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    if "geoopt" not in sys.modules:
        g = mod("geoopt")
        g.manifolds = mod("geoopt.manifolds")
        g.manifolds.stereographic = mod("geoopt.manifolds.stereographic")
        sys.modules["geoopt.manifolds.stereographic.math"] = geoopt_math
        g.manifolds.stereographic.math = geoopt_math
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
        tk = mod("mpl_toolkits")
        tk.axes_grid1 = mod("mpl_toolkits.axes_grid1")
        tk.axes_grid1.make_axes_locatable = lambda *a, **k: None
    if "yacs" not in sys.modules:
        y = mod("yacs")
        y.config = mod("yacs.config")
        y.config.CfgNode = _CfgNode


_loaded = None


def load():
    """Returns a namespace with the reference's HyperMapper, HyperMLR, FloatingRegionScore,
    select_pixels_to_label, RegionSelection and its global cfg."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with cpu_only():
        from core.configs import cfg  # noqa
        from core.utils import hyperbolic as hyp  # noqa
        from core.active import floating_region as fr  # noqa
        from core.active import build as ab  # noqa
    ns = types.SimpleNamespace(
        cfg=cfg,
        HyperMapper=hyp.HyperMapper,
        HyperMLR=hyp.HyperMLR,
        FloatingRegionScore=fr.FloatingRegionScore,
        select_pixels_to_label=ab.select_pixels_to_label,
        RegionSelection=ab.RegionSelection,
        hyperbolic=hyp,
        floating_region=fr,
        build=ab,
    )
    _loaded = ns
    return ns


def load_training_side():
    """The reference's classifier module and negative-learning loss, for the rows next to the hot path (SURVEY 8f rows 3,
    4).  core/models/__init__.py pulls mmcv in through resnet.py, so classifier.py is loaded on its own; it only needs torch
    and core.utils.hyperbolic (already stubbed by `load`)."""
    import importlib.util

    load()
    with cpu_only():
        if "core.models.classifier" not in sys.modules:
            if "core.models" not in sys.modules:
                pkg = types.ModuleType("core.models")
                pkg.__path__ = [os.path.join(REFERENCE_ROOT, "core", "models")]
                sys.modules["core.models"] = pkg
            spec = importlib.util.spec_from_file_location("core.models.classifier",
                                                          os.path.join(REFERENCE_ROOT, "core", "models", "classifier.py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules["core.models.classifier"] = m
            spec.loader.exec_module(m)
        from core.loss.negative_learning_loss import NegativeLearningLoss  # noqa
    return types.SimpleNamespace(classifier=sys.modules["core.models.classifier"], NegativeLearningLoss=NegativeLearningLoss)


@contextlib.contextmanager
def cpu_only():
    """Run reference code on a CUDA-less host: `.cuda()` becomes the identity."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig

"""Install this package under the reference's own module paths.

    import halo_b200; halo_b200.install()          # before `from core.models import build_classifier`

After `install()` the reference's unmodified `core/models/classifier.py`, `core/train_learners.py` and
`core/utils/visualize.py` pick up the CUDA implementations through their usual imports:

    from ..utils.hyperbolic import HyperMapper, HyperMLR                 (classifier.py:4)
    from core.active.build import RegionSelection                        (train_learners.py:12)
    from core.active.floating_region import FloatingRegionScore          (visualize.py:6)
"""
import sys
import types


def install(strict=False):
    """Register `core.utils.hyperbolic`, `core.active.floating_region`, `core.active.build` (and
    `core.active`) so that they resolve to halo_b200.  If the reference modules were already imported their
    public names are rebound in place.  Returns the list of module names that were patched."""
    from . import active, floating_region, hyperbolic

    mapping = {
        "core.utils.hyperbolic": (hyperbolic, ("HyperMapper", "HyperMLR", "PROJ_EPS")),
        "core.active.floating_region": (floating_region, ("FloatingRegionScore",)),
        "core.active.build": (active, ("select_pixels_to_label", "RegionSelection", "to_np_array")),
    }
    patched = []
    for name, (mod, symbols) in mapping.items():
        existing = sys.modules.get(name)
        if existing is not None and existing is not mod:
            for s in symbols:
                setattr(existing, s, getattr(mod, s))
        else:
            parent_name, _, leaf = name.rpartition(".")
            parent = sys.modules.get(parent_name)
            if parent is None and strict:
                raise RuntimeError("install(strict=True): package %s is not importable" % parent_name)
            if parent is None:
                # make a namespace chain so `import core.utils.hyperbolic` works even without the checkout
                chain = parent_name.split(".")
                for i in range(1, len(chain) + 1):
                    pn = ".".join(chain[:i])
                    if pn not in sys.modules:
                        m = types.ModuleType(pn)
                        m.__path__ = []
                        sys.modules[pn] = m
                        if i > 1:
                            setattr(sys.modules[".".join(chain[:i - 1])], chain[i - 1], m)
                parent = sys.modules[parent_name]
            sys.modules[name] = mod
            setattr(parent, leaf, mod)
        patched.append(name)
    core_active = sys.modules.get("core.active")
    if core_active is not None:  # reference core/active/__init__.py does `from .build import *`
        for s in ("select_pixels_to_label", "RegionSelection"):
            setattr(core_active, s, getattr(active, s))
    return patched

"""Import the *unmodified* reference modules from /root/reference (build container only).

Used by `make_golden.py` (fixture generation), by the `needs_reference` tests that pin the oracle
against the live reference, and by `baseline/ref_step.py` (the reference arm of bench.py, which
points REF_ROOT at the git-ignored reference install `baseline/_ref/DPItorch` on the GPU box).
Never imported by the product path or any `-m gpu` test.

Third-party imports the reference makes at module scope (matplotlib, torchkbnufft, ehtim,
astropy, pynfft - none installed here) are replaced by stub modules; the two ehtim helpers the
operator setup calls (`ftmatrix`, `NFFTInfo`) are provided by the restatements in
`dpi_b200.synthetic`.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REF_ROOT = "/root/reference/DPItorch"


def reference_available() -> bool:
    return os.path.isdir(REF_ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_loaded = {}


def import_reference():
    """Returns dict(realnvp=<module>, interferometry=<module>, mri=<module>, geometric=<module>)."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    from dpi_b200 import synthetic as syn

    if not hasattr(np, "int"):
        np.int = int  # removed in numpy>=1.24; used at interferometry_helpers.py:124-128

    class _Any:
        def __init__(self, *a, **k):
            pass

        def to(self, *a, **k):
            return self

    _stub("matplotlib")
    _stub("matplotlib.pyplot", ion=lambda: None)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    _stub("torchkbnufft", KbNufft=_Any, AdjKbNufft=_Any)
    _stub("torchkbnufft.mri")
    _stub("torchkbnufft.mri.dcomp_calc", calculate_radial_dcomp_pytorch=None)
    _stub("torchkbnufft.math", absolute=None)
    _stub("ehtim", RADPERUAS=syn.RADPERUAS)
    _stub("ehtim.observing")
    _stub("ehtim.observing.obs_helpers", ftmatrix=syn.ftmatrix, NFFTInfo=syn.NFFTInfoLike)
    _stub("ehtim.const_def", FFT_PAD_DEFAULT=2, GRIDDER_P_RAD_DEFAULT=2)
    _stub("astropy")
    _stub("astropy.io")
    _stub("astropy.io.fits")
    _stub("pynfft")
    _stub("pynfft.nfft", NFFT=_Any)

    # The reference packages are imported under their own top-level names; make sure none of
    # ours shadows them (ours live under the `dpi_b200.` namespace).
    for k in ("generative_model", "interferometry_helpers", "MRI_helpers", "geometric_model"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_ROOT)
    try:
        import generative_model.realnvpfc_model as realnvp
        import generative_model.glow_model as glow
        import interferometry_helpers as interferometry
        import MRI_helpers as mri
        import geometric_model as geometric
    finally:
        sys.path.remove(REF_ROOT)
    # park them under private names so a later `import interferometry_helpers` of the drop-in
    # package cannot collide with the reference copy
    for k in ("generative_model", "generative_model.realnvpfc_model", "generative_model.glow_model",
              "interferometry_helpers", "MRI_helpers", "geometric_model"):
        mod = sys.modules.pop(k, None)
        if mod is not None:
            sys.modules["_dpi_reference." + k] = mod
    _loaded.update(realnvp=realnvp, glow=glow, interferometry=interferometry, mri=mri,
                   geometric=geometric)
    return _loaded

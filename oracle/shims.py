"""Oracle shims: register the restatements as ``timm`` / ``segmentation_models_pytorch``.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  With these in ``sys.modules`` the
reference's own ``code/models/*.py`` and ``code/losses/*.py`` import and run
unmodified (SURVEY §8c feasibility probe).  Only names the reference touches are
provided: ``timm.create_model`` (encoders.py:53), ``smp.decoders.fpn.decoder.FPNDecoder``
(decoders.py:6), ``smp.base.{SegmentationHead,ClassificationHead}`` (heads.py:33,85,...),
``smp.losses.DiceLoss`` (loss_functions.py:169), ``smp.encoders.get_encoder``
(encoders.py:774, non-Swin names only -> raises).
"""

import sys
import types

from . import swin, fpn, heads


def install(force: bool = False) -> None:
    if "timm" in sys.modules and not force and not getattr(sys.modules["timm"], "__oracle_shim__", False):
        return  # a real timm is present: never shadow it
    timm = types.ModuleType("timm")
    timm.__oracle_shim__ = True
    timm.create_model = swin.create_model
    sys.modules["timm"] = timm

    smp = types.ModuleType("segmentation_models_pytorch")
    smp.__oracle_shim__ = True
    smp.__path__ = []
    base = types.ModuleType("segmentation_models_pytorch.base")
    base.SegmentationHead = heads.SegmentationHead
    base.ClassificationHead = heads.ClassificationHead
    losses = types.ModuleType("segmentation_models_pytorch.losses")
    losses.DiceLoss = heads.DiceLoss
    decoders = types.ModuleType("segmentation_models_pytorch.decoders")
    decoders.__path__ = []
    dfpn = types.ModuleType("segmentation_models_pytorch.decoders.fpn")
    dfpn.__path__ = []
    dec = types.ModuleType("segmentation_models_pytorch.decoders.fpn.decoder")
    dec.FPNDecoder = fpn.FPNDecoder
    encoders = types.ModuleType("segmentation_models_pytorch.encoders")

    def get_encoder(name, *a, **k):
        raise NotImplementedError(f"oracle shim: smp encoder '{name}' is outside the hot path (Swin only)")

    encoders.get_encoder = get_encoder
    smp.base, smp.losses, smp.decoders, smp.encoders = base, losses, decoders, encoders
    decoders.fpn = dfpn
    dfpn.decoder = dec
    for m in (smp, base, losses, decoders, dfpn, dec, encoders):
        sys.modules[m.__name__] = m
    # importlib.util.find_spec() (used by libraries that probe for optional dependencies, e.g. transformers probing for
    # timm) raises on a module whose __spec__ is None: give every stub a spec
    import importlib.machinery
    for m in (timm, smp, base, losses, decoders, dfpn, dec, encoders):
        m.__spec__ = importlib.machinery.ModuleSpec(m.__name__, loader=None)


def import_reference_models(reference_root: str = "/root/reference"):
    """Import the reference's own ``models`` / ``losses`` / ``configs`` packages on top of the shims.

    Only usable where /root/reference exists (the authoring container); returns
    (models, losses, configs) modules.
    """
    import importlib
    import os
    code = os.path.join(reference_root, "code")
    if not os.path.isdir(code):
        raise FileNotFoundError(code)
    install()
    if code not in sys.path:
        sys.path.insert(0, code)
    return (importlib.import_module("models"), importlib.import_module("losses"),
            importlib.import_module("configs"))

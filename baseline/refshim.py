"""Import the UNMODIFIED reference (baseline/_ref/, else /root/reference) in-process -- MEASUREMENT / TEST INFRASTRUCTURE.

The reference needs four packages this image lacks (albumentations, matplotlib, imagesize, torchmetrics); none of them is
touched by the hot path (they serve augmentation, plotting, image-size probing, mAP), so they are stubbed with MagicMock as
in SURVEY.md Appendix C.  Only bench.py's reference / cpu_baseline legs and tests/ use this module.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = ["albumentations", "albumentations.pytorch", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "imagesize",
          "torchmetrics", "torchmetrics.detection", "torchmetrics.detection.mean_ap"]
_REF_MODULES = ["config", "model", "loss", "ultralytics_loss", "dataset", "utils", "utils.bboxes_utils", "utils.plot_utils",
                "utils.utils", "utils.training_utils", "utils.validation_utils"]


def ref_path():
    for p in (os.path.join(HERE, "_ref"), os.environ.get("YOLO_REF", "/root/reference")):
        if p and os.path.isfile(os.path.join(p, "model.py")):
            return p
    return None


def import_reference(device="cpu"):
    """returns a namespace with the reference's modules (config, model, ultralytics_loss, loss, bboxes_utils, plot_utils,
    training_utils, validation_utils, utils_utils); raises ImportError when the reference is not available"""
    path = ref_path()
    if path is None:
        raise ImportError("reference not available (neither baseline/_ref nor /root/reference)")
    for m in _STUBS:
        if m not in sys.modules:
            try:
                importlib.import_module(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if path not in sys.path:
        sys.path.insert(0, path)
    import config as rconfig
    if os.path.dirname(os.path.abspath(rconfig.__file__)) != os.path.abspath(path):
        raise ImportError(f"another module named `config` shadows the reference's ({rconfig.__file__})")
    rconfig.DEVICE = device
    ns = types.SimpleNamespace(path=path, config=rconfig)
    ns.model = importlib.import_module("model")
    ns.ultralytics_loss = importlib.import_module("ultralytics_loss")
    ns.loss = importlib.import_module("loss")
    ns.bboxes_utils = importlib.import_module("utils.bboxes_utils")
    ns.plot_utils = importlib.import_module("utils.plot_utils")
    ns.training_utils = importlib.import_module("utils.training_utils")
    ns.validation_utils = importlib.import_module("utils.validation_utils")
    ns.utils_utils = importlib.import_module("utils.utils")
    return ns

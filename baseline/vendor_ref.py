"""Vendor the reference's Python sources into the git-ignored baseline/_ref/ (SURVEY.md 8c).

The reference (AlessandroMondin/YOLOV5m) is pure Python with no setup.py / pyproject.toml, so there is nothing to
`pip install`: "installing" it means making its modules importable.  /root/reference does not exist on the GPU box; the
git-ignored (but gpurun-shipped) directory baseline/_ref/ carries the UNMODIFIED files there, so that
`bench.py --impl reference`, `cpu_baseline` and tests/test_reference_integration.py run the reference itself, not a port.
Nothing under baseline/_ref/ is tracked and no product code imports it.

    python baseline/vendor_ref.py            # called by __graft_entry__.build() when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("YOLO_REF", "/root/reference")
FILES = ["model.py", "ultralytics_loss.py", "loss.py", "config.py", "dataset.py", "coco.py", "train.py", "detect.py",
         "utils/__init__.py", "utils/bboxes_utils.py", "utils/plot_utils.py", "utils/utils.py", "utils/training_utils.py",
         "utils/validation_utils.py"]


def vendor(src=SRC, dst=DST):
    """copy (never edit) the reference modules; returns the destination or None when the reference is not present"""
    if not os.path.isfile(os.path.join(src, "model.py")):
        return dst if os.path.isfile(os.path.join(dst, "model.py")) else None
    for f in FILES:
        s, d = os.path.join(src, f), os.path.join(dst, f)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.isfile(d) or open(s, "rb").read() != open(d, "rb").read():
            shutil.copyfile(s, d)
    return dst


if __name__ == "__main__":
    out = vendor()
    print(out if out else "reference not present; baseline/_ref not populated")
    sys.exit(0)

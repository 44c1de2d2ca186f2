"""CPU oracle for the YOLOV5m hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain PyTorch-CPU / numpy restatement of the reference
algorithms (AlessandroMondin/YOLOV5m: model.py, ultralytics_loss.py,
utils/bboxes_utils.py, utils/plot_utils.py).  It is the *checker* for the
CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``yolov5m_b200`` never imports anything from here.

Parity pin: every function here is checked against golden vectors produced by
importing the real reference (``/root/reference``) in the build container --
see ``tests/golden/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""

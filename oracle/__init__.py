"""CPU oracle for the DeepInteraction MMRI+MMPI forward path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It restates, in plain
PyTorch/numpy/cv2 (fp32, CPU-runnable), the algorithm of the reference modules
listed in SURVEY.md section 8(a); each function cites the reference file:line it
follows (paths relative to the reference checkout).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``deepinteraction_b200`` and
``projects.mmdet3d_plugin``) never imports it and fails loudly if the CUDA
library is missing.

Parity pinning: the reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference's own
Python files, stub-loaded in the build container by
``tools/make_goldens.py`` and committed under ``tests/golden/``.
Third-party semantics that are NOT under the reference tree (mmdet3d 0.17.1
``apply_3d_transformation`` / ``LiDARInstance3DBoxes.corners``, detectron2
``ROIAlignV2``) are restated from their published behaviour in
``oracle/geometry.py`` -- for those three boundaries parity is "unpinned" in the
sense of the task statement (the goldens use the same restatement as stub).
"""
from . import geometry, depth_completion, mmri, mmpi  # noqa: F401

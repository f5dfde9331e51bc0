"""CPU oracle for the face-landmark hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path
(``face-landmark-detector_b200/``) never imports it and fails loudly when the
CUDA library is missing.

Parity pinning status (see DESIGN.md "Oracle"):
  * host-side functions (box math, crop/resize, regression decode, argmax map,
    get_average_xy family, get_image_array) are PINNED: the restatements here are
    checked against the reference's own functions executed from /root/reference
    under stubbed heavy imports (tests/make_golden.py -> tests/golden/*.npz).
  * cv2.resize / cv2.warpAffine integer models are PINNED against OpenCV 4.13.0
    (the version in this image; the reference leaves OpenCV unpinned).
  * CNN arithmetic (Keras/TensorFlow, unpinned, not installable here) is
    "parity unpinned": restated from networks/*.py + Keras layer semantics, and
    cross-checked by two independent restatements (torch functional vs numpy
    einsum) only.
  * alignment (Umeyama + warp) does not exist in the reference: build-defined,
    oracle = fp64 closed form cross-checked against an SVD formulation, warp =
    cv2.warpAffine 4.13.
"""

"""CPU oracle for the detect -> landmark -> iris path of rs-face-detection-tflite.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker / the timed CPU baseline.  The product (``libfdl.so`` + the ctypes
mirror in ``rs-face-detection-tflite_b200/``) never imports this package and
fails loudly when the CUDA library is missing.

What it restates (all citations relative to /root/reference):

* Rust glue  -- ``src/face_detection_lite/{face_detection,transform,nms,types,
  face_landmark,iris_landmark}.rs`` in numpy (``oracle/glue.py``).
* OpenCV ops -- ``resize``/``warpPerspective``/``copyMakeBorder``/``flip`` used by
  ``transform.rs:222-297``: integer-exact numpy restatement
  (``oracle/cv_ops.py``), itself checked bit-for-bit against cv2 4.13 (the same
  library family the reference links through crate ``opencv 0.93.1``).
* TFLite     -- the float reference semantics of the 11 builtin ops the five
  dense ``.tflite`` graphs use, executed with torch-CPU fp32
  (``oracle/tflite_reader.py`` + ``oracle/graph_exec.py``).  The TFLite runtime
  (crate ``tflite 0.9.8``) is an un-vendored dependency and cannot be built here.

Parity pin: the reference's own tests assert nothing, so the
oracle is pinned to the only golden artefacts the reference holds -- the three
rendered PNGs in ``assets/`` written by ``src/lib.rs:42-83`` -- pixel-exactly
(``tests/test_oracle_kat.py``; ``tests/test_golden.py`` re-renders the oracle's
results with ``oracle/render.py``, a restatement of ``render.rs`` + imageproc's
drawing routines, and requires the painted pixel sets to equal the PNGs'), and
to cv2 for the OpenCV half and the JPEG ingest (``oracle/jpeg_decode.py``).  Beyond that
parity is "unpinned by the reference's tests" and DESIGN.md says so.
"""

"""Generates the fixtures in this directory.  Run in the build container from the repo root: `python tests/golden/make_golden.py`.

Two kinds of fixture, kept apart:

* `k1_reference_assets.json` -- facts decoded from the REFERENCE's own rendered outputs, the only result artefacts its repository
  holds for this path (`/root/reference/assets/man_{bbox,landmark,iris}.png`, written by src/lib.rs:42-83 through render.rs): the
  extent of the pure-green detection rectangle, the extents and the exact pixel sets of the pure-red landmark / iris drawings.
  These pin the oracle (tests/test_oracle_kat.py, tests/test_golden.py) and, through it or directly, the CUDA path.
* `man_pipeline_oracle.npz` -- the oracle's outputs on test_data/man.jpg (detections, face ROI, 468 landmarks, eye ROIs, eye
  contours and irises, plus the refined landmark set).  The reference itself cannot run here (Rust, un-vendored native
  dependencies), so these are ORACLE vectors, pinned by the K1 facts above; they let the `-m gpu` tests compare the CUDA path with
  committed numbers and flag any drift of the oracle itself (another torch / numpy / OpenCV build).

Nothing under tests/ reads /root/reference at run time; only this script does.
"""
import hashlib
import json
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_ASSETS = "/root/reference/assets"


def _mask(name, rgb):
    bgr = cv2.imread(os.path.join(REF_ASSETS, name), cv2.IMREAD_COLOR)
    img = bgr[:, :, ::-1]
    return (img[:, :, 0] == rgb[0]) & (img[:, :, 1] == rgb[1]) & (img[:, :, 2] == rgb[2])


def _extent(m):
    ys, xs = np.nonzero(m)
    return [int(xs.min()), int(xs.max()), int(ys.min()), int(ys.max())]


def main():
    facts = {"source": "reference assets/*.png (src/lib.rs:42-83, render.rs:361-479)", "image": "test_data/man.jpg", "size": [540, 360]}
    green = _mask("man_bbox.png", (0, 255, 0))
    facts["bbox_green_extent_x0x1y0y1"] = _extent(green)
    facts["bbox_green_pixels"] = int(green.sum())
    for key, name in (("landmark", "man_landmark.png"), ("iris", "man_iris.png")):
        red = _mask(name, (255, 0, 0))
        facts[key + "_red_extent_x0x1y0y1"] = _extent(red)
        facts[key + "_red_pixels"] = int(red.sum())
        facts[key + "_red_mask_sha256"] = hashlib.sha256(np.packbits(red).tobytes()).hexdigest()
        np.save(os.path.join(HERE, "k1_%s_red_mask.npy" % key), np.packbits(red, axis=1))
    with open(os.path.join(HERE, "k1_reference_assets.json"), "w") as f:
        json.dump(facts, f, indent=1, sort_keys=True)

    import synth_frames
    from oracle import glue, pipeline
    man = synth_frames.load_rgb("man.jpg")
    p = pipeline.Pipeline(glue.BACK_CAMERA, os.path.join(ROOT, "models"))
    faces, out = p.run(man)
    o = out[0]
    roi = lambda r: np.array([r.x_center, r.y_center, r.width, r.height, r.rotation], np.float64)
    lm = np.asarray(o["landmarks"], np.float64)
    arrays = dict(
        image_sha256_16=np.frombuffer(hashlib.sha256(man.tobytes()).hexdigest()[:16].encode(), np.uint8),
        det_data=np.stack([np.asarray(f.data, np.float32) for f in faces]), det_score=np.array([f.score for f in faces], np.float32),
        face_roi=roi(o["roi"]), landmarks=lm, left_roi=roi(o["left_roi"]), right_roi=roi(o["right_roi"]),
        left_contour=np.asarray(o["left"][0], np.float64), left_iris=np.asarray(o["left"][1], np.float64),
        right_contour=np.asarray(o["right"][0], np.float64), right_iris=np.asarray(o["right"][1], np.float64),
        refined_landmarks=np.asarray(glue.update_face_landmarks_with_iris_results(lm, np.asarray(o["left"][0]), np.asarray(o["right"][0])), np.float64),
    )
    np.savez_compressed(os.path.join(HERE, "man_pipeline_oracle.npz"), **arrays)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()

"""Burst classification of detector C.

The reference (meteor_detect_class/detector_and_classification.py:7-91) finds
ORB keypoints in a *rendered JPG* of the spectrogram, clusters them with DBSCAN
and calls a cluster critical when its bounding box is >= 5 px wide ("about
0.5 s", :50).  That image pipeline is out of scope of the GPU hot path
(SURVEY.md section 2 #8 / section 8 C-classify: it works on a lossy rendering,
not on signal arithmetic).  What is kept:

* the classification RULE, applied to event durations measured on the signal
  itself (``classify_events``) -- this is what feeds the ``Kritisch`` column;
* the call SIGNATURE and 5-tuple return of ``detect_and_cluster_bursts`` so
  existing callers keep working: it runs the same third-party library stage
  (OpenCV ORB + scikit-learn DBSCAN, unchanged CPU libraries, exactly as in the
  reference) when those libraries are installed.  It is not part of the
  measured path and no GPU kernel stands behind it.
"""
from __future__ import annotations

import numpy as np

CRITICAL_MIN_WIDTH_PX = 5          # detector_and_classification.py:50
CRITICAL_MIN_DUR_S = 0.5           # "entspricht ca. 0.5s" (:50); README.md:75-76


def classify_events(durations_s, critical_min_dur_s: float = CRITICAL_MIN_DUR_S):
    """(critical, non_critical) index lists for events of the given durations."""
    d = np.asarray(durations_s, dtype=np.float64)
    crit = np.nonzero(d >= critical_min_dur_s)[0].tolist()
    non = np.nonzero(d < critical_min_dur_s)[0].tolist()
    return crit, non


def detect_and_cluster_bursts(image_path, eps=30, min_samples=5, display=True, output_path=None):
    """Same signature and return tuple as the reference (:7, :91):
    ``(bursts, unique_labels, burst_positions, critical_bursts, non_critical_bursts)``."""
    try:
        import cv2
        from sklearn.cluster import DBSCAN
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("detect_and_cluster_bursts needs OpenCV and scikit-learn (image stage, out of scope of "
                           "the GPU path); use classify_events() on detector output instead") from e
    image = cv2.imread(image_path, cv2.IMREAD_COLOR)
    orb = cv2.ORB_create(nfeatures=500, edgeThreshold=0, scaleFactor=1.2)      # :12
    keypoints, _ = orb.detectAndCompute(image, None)
    pts = np.array([kp.pt for kp in keypoints])
    labels = DBSCAN(eps=eps, min_samples=min_samples).fit(pts).labels_ if len(pts) else []   # :19-23
    unique_labels = set(labels)
    critical, non_critical = [], []
    for label in unique_labels:
        if label == -1:
            continue
        cl = pts[labels == label]
        (x_min, y_min), (x_max, y_max) = cl.min(axis=0), cl.max(axis=0)
        is_crit = (x_max - x_min) >= CRITICAL_MIN_WIDTH_PX                       # :48-50
        (critical if is_crit else non_critical).append(label)
        if output_path:
            cv2.rectangle(image, (int(x_min), int(y_min)), (int(x_max), int(y_max)),
                          (0, 255, 0) if is_crit else (0, 0, 255), 2)
    if output_path:
        cv2.imwrite(output_path, image)
    return [], unique_labels, [], critical, non_critical

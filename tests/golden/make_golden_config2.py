"""One-time golden of BASELINE configs[2] at FULL size, written by THE REFERENCE.

    python tests/golden/make_golden_config2.py        (authoring container only: needs /root/reference; ~5-10 min)

Runs the reference's `mean_shift_segmentation` (`cellulus/utils/mean_shift.py:6-45`, loaded by file path,
scikit-learn MeanShift underneath, hill climb on one core as shipped) on the bench volume of `bench.py`:
`synthetic.blob_scene((128, 256, 256), 400, radius=10, seed=0)`, bandwidth 7, reduction_probability 0.1,
threshold 0.5, `np.random.seed(0)` immediately before the call.  Stores the int32 label volume as uint16
(K < 65536) in `config2_labels.npz` together with the wall time of the reference call on this host.
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_stubs, load_by_path, synthetic  # noqa: E402

SHAPE, OBJECTS, RADIUS, BW, THR, RP = (128, 256, 256), 400, 10.0, 7.0, 0.5, 0.1


def main():
    import sklearn

    install_stubs()
    ms = load_by_path("ref_mean_shift", os.path.join(REF, "cellulus/utils/mean_shift.py"))
    emb, _, ids = synthetic.blob_scene(SHAPE, OBJECTS, radius=RADIUS, seed=0)
    emb64 = emb.astype(np.float64)
    np.random.seed(0)
    t0 = time.perf_counter()
    labels = ms.mean_shift_segmentation(emb64[np.newaxis, :3].copy(), emb64[3], bandwidth=BW, min_size=0,
                                        reduction_probability=RP, threshold=THR, seeds=None)
    dt = time.perf_counter() - t0
    assert labels.max() < 65536
    np.savez_compressed(os.path.join(HERE, "config2_labels.npz"), labels=labels.astype(np.uint16),
                        cfg=np.array([BW, RP, THR, RADIUS, OBJECTS], dtype=np.float64), shape=np.array(SHAPE),
                        seconds=np.array(dt), sklearn_version=np.array(sklearn.__version__),
                        foreground=np.array(int((labels > 0).sum())))
    print(f"reference mean_shift_segmentation on {SHAPE}: {dt:.1f} s, K = {int(labels.max())}, "
          f"fg = {int((labels > 0).sum())}")


if __name__ == "__main__":
    main()

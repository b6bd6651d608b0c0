"""Generates tests/golden/extract_golden.npz by EXECUTING THE UNMODIFIED REFERENCE function
scripts/extract_data.py:120 `run_correspondence_extraction` (sklearn brute-force kneighbors, k = 2) on seeded synthetic
feature files written to a temporary directory.  Run in the build container only:

    python tests/golden/make_extract_golden.py

The function subsamples every scan with np.random.choice (scripts/extract_data.py:160-168); the permutations are
recovered by replaying the legacy RandomState from the same seed and are stored next to the outputs so that the test
can rebuild the inputs from oracle.lmpcr_oracle.synth_scene.
"""
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lmpcr_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N_SCANS, N_PTS, SEED, NP_SEED = 3, 1000, 23, 1234


def main():
    refimport.import_reference()
    o3d = sys.modules["open3d"]
    o3d.utility = types.SimpleNamespace(set_verbosity_level=lambda *a, **k: None, VerbosityLevel=types.SimpleNamespace(Error=0))
    import importlib
    ext = importlib.import_module("scripts.extract_data")

    feats, xyz, _ = O.synth_scene(N_SCANS, N_PTS, seed=SEED)
    out = {"cfg": np.array([N_SCANS, N_PTS, SEED, NP_SEED])}
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "src")
        dst = os.path.join(tmp, "dst")
        os.makedirs(os.path.join(src, "scene0"))
        fdir = os.path.join(dst, "synth", "features", "scene0")
        os.makedirs(fdir)
        for s in range(N_SCANS):
            np.savez_compressed(os.path.join(fdir, "scene0_%03d.npz" % s), feature=feats[s], xyz=xyz[s])
        np.random.seed(NP_SEED)
        ext.run_correspondence_extraction("synth", src, dst, N_PTS, 0)
        # replay the sampling
        np.random.seed(NP_SEED)
        for a in range(N_SCANS):
            for b in range(a + 1, N_SCANS):
                i1 = np.random.choice(N_PTS, N_PTS, replace=False)
                i2 = np.random.choice(N_PTS, N_PTS, replace=False)
                d = np.load(os.path.join(dst, "synth", "correspondences", "scene0", "scene0_%03d_%03d.npz" % (a, b)))
                key = "p%d_%d_" % (a, b)
                assert np.array_equal(d["x"][:, 3:], xyz[b][i2]), "sampling replay out of step"
                out[key + "inds1"] = i1.astype(np.int16)
                out[key + "inds2"] = i2.astype(np.int16)
                out[key + "x"] = d["x"]
                out[key + "mutuals"] = d["mutuals"].astype(np.uint8)
                out[key + "ratios"] = d["ratios"]
                out[key + "dtypes"] = np.array([str(d["x"].dtype), str(d["mutuals"].dtype), str(d["ratios"].dtype)])
                out[key + "shapes"] = np.array([d["x"].shape[1], d["mutuals"].shape[1], d["ratios"].ndim])
    np.savez_compressed(os.path.join(OUT, "extract_golden.npz"), **out)
    print("extract_golden.npz", os.path.getsize(os.path.join(OUT, "extract_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

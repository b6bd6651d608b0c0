"""Generates tests/golden/overlap_golden.npz by EXECUTING THE UNMODIFIED REFERENCE lib/utils.py:713
`compute_overlap_ratio(method='3DMatch')` (sklearn KD-tree) on seeded synthetic clouds.  The 'FCGF' method needs Open3D's
voxel_down_sample, which is not installed here: for that method the golden holds the reference's KD-tree counting applied
to clouds down-sampled by the oracle's restatement of Open3D (documented as unpinned in oracle/lmpcr_oracle.py).

    python tests/golden/make_overlap_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lmpcr_oracle as O  # noqa: E402
from oracle import refimport  # noqa: E402
import synthdata  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


CASES = [("a", 11, 3000, 2500, 0.5), ("b", 12, 5000, 5000, 0.2), ("c", 13, 777, 4000, 0.8), ("d", 14, 2000, 2000, 0.0)]


def main():
    lib = refimport.import_reference()
    out = {}
    for name, seed, ni, nj, ov in CASES:
        pi, pj, T = synthdata.synth_cloud_pair(seed, ni, nj, ov)
        out[name + "_cfg"] = np.array([seed, ni, nj, ov])
        out[name + "_T"] = T
        out[name + "_3dmatch"] = np.array(lib.utils.compute_overlap_ratio(pi, pj, T, method="3DMatch"))
        di, dj = O.voxel_down_sample(pi, 0.025), O.voxel_down_sample(pj, 0.025)
        # the reference's counting (KD-tree, 3 voxels) on the down-sampled clouds: '3DMatch' code path with a patched radius is
        # not available, so reproduce :764-775 with the reference's own NearestNeighbors call sequence
        from sklearn.neighbors import NearestNeighbors
        neigh = NearestNeighbors(n_neighbors=1, algorithm="kd_tree")
        Ti = np.linalg.inv(T)
        di_t = (Ti[:3, :3] @ di.T + Ti[:3, 3].reshape(-1, 1)).T
        dj_t = (T[:3, :3] @ dj.T + T[:3, 3].reshape(-1, 1)).T
        neigh.fit(dj_t); d01, _ = neigh.kneighbors(di, return_distance=True)
        neigh.fit(di_t); d10, _ = neigh.kneighbors(dj, return_distance=True)
        out[name + "_fcgf"] = np.array(max((d01 < 0.075).sum() / len(di), (d10 < 0.075).sum() / len(dj)))
        out[name + "_fcgf_sizes"] = np.array([len(di), len(dj)])
        print(name, float(out[name + "_3dmatch"]), float(out[name + "_fcgf"]), len(di), len(dj))
    np.savez_compressed(os.path.join(OUT, "overlap_golden.npz"), **out)


if __name__ == "__main__":
    main()

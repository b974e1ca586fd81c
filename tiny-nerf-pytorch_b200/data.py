"""data.load_tiny_nerf_npz -- same contract as the reference's src/data.py:4-13."""
from typing import Any, Dict

import numpy as np


def load_tiny_nerf_npz(path: str = "data/tiny_nerf_data.npz") -> Dict[str, Any]:
    """Arrays of the .npz by key ('images' (N,H,W,3), 'poses' (N,4,4), 'focal'); float64 -> float32."""
    with np.load(path) as archive:
        out = {}
        for key in archive.files:
            arr = archive[key]
            out[key] = arr.astype(np.float32) if getattr(arr, "dtype", None) == np.float64 else arr
    return out

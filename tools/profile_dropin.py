"""cProfile of the reference's UNMODIFIED train.py running on this repo's modules (host-side cost of the drop-in route); developer tool
for the GPU box:  python tools/profile_dropin.py [iters]"""
import cProfile
import os
import pstats
import runpy
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tiny-nerf-pytorch_b200")
sys.path[:0] = [os.path.join(ROOT, "tools", "_timing_shims"), PKG, os.path.join(PKG, "_shims")]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600
work = tempfile.mkdtemp(prefix="tnerf_prof_")
os.makedirs(os.path.join(work, "data"))
rng = np.random.default_rng(0)
poses = np.tile(np.eye(4, dtype=np.float32), (8, 1, 1)); poses[:, 2, 3] = 4.0
np.savez(os.path.join(work, "data", "tiny_nerf_data.npz"), images=rng.random((8, 100, 100, 3), dtype=np.float32), poses=poses, focal=np.float32(138.9))
os.chdir(work)
far = str(10 ** 9)
sys.argv = ["train.py", "--iters", str(iters), "--n-rand", "4096", "--n-samples", "64", "--log-every", far, "--preview-every", far, "--ckpt-every", far, "--no-resume"]
script = os.path.join(ROOT, "baseline", "_ref", "src", "train.py")
pr = cProfile.Profile()
pr.enable()
runpy.run_path(script, run_name="__main__")
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats("tiny-nerf-pytorch_b200|train.py", 40)
st.sort_stats("tottime").print_stats(25)

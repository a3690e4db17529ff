"""Where the wall time of the drop-in RegionSelection goes (real pipeline shape, fake loader): cProfile of the host side."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
pr = cProfile.Profile()
bench.run_dropin(dev, n_images=4)   # warm
pr.enable()
t0 = time.perf_counter()
res = bench.run_dropin(dev, n_images=int(sys.argv[1]) if len(sys.argv) > 1 else 24)
pr.disable()
print(res)
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)

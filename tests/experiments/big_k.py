"""One GPU, unsharded K = 32768: per-kernel times of the update kernels at the C5 global size (ncu launch list)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "m3p2i-aip_b200"), ROOT]
from m3p2i_b200 import _abi as A, native, scene as S
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
conf = dict(bench.CONFIGS[name]); conf["K"] = K
p, _, dof, root = bench.build_planner(conf, 1, 0, 0, "none")
for _ in range(6):
    i = p.command_resident(sync=True)
print(name, K, "command ms", i.kernel_ms, "rollout ms", i.rollout_ms, "beta iters", i.beta_iters, "lanes", i.rollout_lanes)

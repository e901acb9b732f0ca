"""One 2-D refinement (N = 10^4 points, ladam, K = 50) for use under ncu:
    ncu --set full --clock-control none --import-source on -k regex:mlp2d_refine -c 1 -o gpurun_out/prof_2d python tools/profile_2d.py
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "collaborative-gan-sampling_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cgs import synthetic as S  # noqa: E402
from sampling.refiner_cpu import MlpSpec, Refiner  # noqa: E402


class _Data:
    def __init__(self, pts):
        self.pts = pts

    def next_batch(self, n):
        return self.pts[:n]


dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
ws = S.init_mlp2d(64, 6, seed=2019, gain=1.5)
mlp = MlpSpec(ws, dev)
rng = np.random.RandomState(1)
x0 = (rng.randn(n, 2) * 4).astype(np.float32)
real = (rng.randn(n, 2) * 3).astype(np.float32)
ref = Refiner(types.SimpleNamespace(rollout_steps=50, rollout_rate=0.1, rollout_method="ladam"))
ref.set_env(mlp, None, _Data(real))
for _ in range(2):
    out = ref.manipulate_sample(x0, "deterministic")
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
out = ref.manipulate_sample(x0, "deterministic")
e.record()
e.synchronize()
print("n", n, "ms incl. host side", s.elapsed_time(e))

"""Small end-to-end run of the image path for compute-sanitizer (memcheck / racecheck): every round-2 kernel once --
class-fused tcgen05 tiles (forced), tcgen05 edge kernels on the s2d layout, device-side early exit, graph replay."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "collaborative-gan-sampling_b200"))
import torch
from cgs import lib as L, nets as N, synthetic as S
from sampling.collaborator import Refiner

lib = L.load()
dev = torch.device("cuda", 0)
# default rules / class-fused tiles whenever legal / class fusion + M-tile pairs / CTA pairs (cta_group::2) whenever legal
for flags, name, B, thr in ((0, "dcgan64_l1", 5, None), (1048576, "dcgan64_l1", 5, None), (1048576, "dcgan32_l2", 6, None),
                            (1048576 | 8388608, "dcgan64_l3", 3, 0.0), (8388608, "dcgan32_l4", 4, None),
                            (1048576 | 33554432, "dcgan64_l1", 5, None), (33554432, "dcgan32_l2", 7, 0.0),
                            (1048576, "mnist", 20, 0.1)):
    lib.cgs_debug_set_flags(flags)
    arch = N.get_arch(name)
    spec = N.NetSpec(arch, S.init_weights(arch, gain=3.0 if name == "mnist" else 2.5), dev)
    r = Refiner(3, 0.1)
    r.set_env(N.discriminator_spec(spec), N.feature_to_data_spec(spec), N.loss_refine)
    r.early_exit_logit = thr
    h0 = torch.from_numpy(S.proposal_features(arch, B, seed=1)).to(dev)
    x = r.build_refiner(h0)
    torch.cuda.synchronize()
    print(flags, name, "ok", tuple(x.shape), float(r.optimal_logit.mean()))
print("launches", lib.cgs_launch_count())

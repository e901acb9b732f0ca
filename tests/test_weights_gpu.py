"""Weight interchange on the GPU: a spec built from an .npz in ANOTHER layout (PyTorch-ordered
kernels, Saver-decorated names) packs to the same device weights and matches the oracle fed the TF-layout variables."""
import numpy as np
import pytest
import torch

from oracle import graph_refiner as gr
from oracle import nets as onets

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["mnist", "dcgan32_l2"])
def test_spec_from_torch_layout_npz_matches_oracle(cgs_lib, cuda_device, tmp_path, name):
    from cgs import nets as N, weights as W
    from sampling.collaborator import Refiner
    arch = N.get_arch(name)
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=6), 3.0)
    path = tmp_path / "external.npz"
    W.save_npz(path, {k + ":0": v for k, v in W.to_torch_layout(arch, w, include_head=True).items()})
    spec_ext = N.NetSpec.from_npz(arch, path, cuda_device, math="fp32", layout="torch")
    spec_ref = N.NetSpec(arch, w, cuda_device, math="fp32")
    for a, b in zip(spec_ext.gtail.tensors + spec_ext.d.tensors, spec_ref.gtail.tensors + spec_ref.d.tensors):
        assert torch.equal(a, b)                      # same packed device weights, bit for bit
    h0 = torch.relu(torch.randn(5, *arch["feature_shape"], generator=torch.Generator().manual_seed(1)))
    ref = Refiner(2, 0.1)
    ref.set_env(N.discriminator_spec(spec_ext), N.feature_to_data_spec(spec_ext), N.loss_refine)
    x = ref.build_refiner(h0.to(cuda_device))
    o = gr.build_refiner(h0, arch, w, 2, 0.1)
    assert np.abs(x.cpu().numpy() - o["refined"].numpy()).max() <= 1e-4
    with pytest.raises(ValueError, match="another order"):
        N.NetSpec.from_npz(arch, path, cuda_device, math="fp32", layout="tf")      # torch-ordered file read as TF

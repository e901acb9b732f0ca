"""Oracle self-consistency (CPU): TF SAME padding identities, BN folding, best-image-kept shortcut."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import graph_refiner as gr
from oracle import nets as onets


@pytest.mark.parametrize("k", [4, 5])
def test_same_padding_identities(k):
    """conv SAME s2: pad-before 1 (k4: after 1, k5: after 2); deconv = data-gradient of that conv (App. A8)."""
    torch.manual_seed(0)
    x = torch.randn(2, 8, 8, 3)
    w = torch.randn(k, k, 3, 5)
    y = onets.conv2d_same(x, w, torch.zeros(5))
    assert y.shape == (2, 4, 4, 5)
    xp = F.pad(x.permute(0, 3, 1, 2), (1, k - 3, 1, k - 3))
    assert torch.allclose(y, F.conv2d(xp, w.permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1), atol=1e-5)
    # deconv with filter [k,k,Cout,Cin] is the vjp of the conv with the same filter read as [k,k,Cin',Cout']
    wd = torch.randn(k, k, 3, 5)                      # deconv: 5 -> 3 channels
    g = torch.randn(2, 4, 4, 5)
    xr = torch.zeros(2, 8, 8, 3, requires_grad=True)
    (vjp,) = torch.autograd.grad((onets.conv2d_same(xr, wd, torch.zeros(5)) * g).sum(), xr)
    assert torch.allclose(onets.deconv2d_same(g, wd, torch.zeros(3)), vjp, atol=1e-4)
    if k == 5:   # the 'obvious' torch arguments are wrong for k5
        bad = F.conv_transpose2d(g.permute(0, 3, 1, 2), wd.permute(3, 2, 0, 1), stride=2, padding=2, output_padding=1)
        assert not torch.allclose(bad.permute(0, 2, 3, 1), vjp, atol=1e-4)


def test_bn_folding_equals_inference_bn(cgs_lib):
    from cgs import nets as N
    arch = onets.get_arch("mnist")
    w = onets.init_weights(arch, seed=3)
    layer = arch["d"][1]                              # d_conv2 + d_bn2
    x = torch.randn(2, 14, 14, 64)
    ref = onets.run_layers(x, [dict(layer, act="none")], "discriminator", w, "inference")
    wf, bf = N.fold_layer(layer, "discriminator", w)
    got = onets.conv2d_same(x, wf, bf)
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5)


def test_best_image_kept_equals_rerun_of_tail():
    """collaborator.py:88 re-runs feature_to_data(optimal_feature); keeping the best image is identical."""
    arch = onets.get_arch("mnist")
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=1), 3.0)
    h0 = torch.relu(torch.randn(5, 7, 7, 128, generator=torch.Generator().manual_seed(0)))
    o = gr.build_refiner(h0, arch, w, 4, 0.1)
    assert torch.equal(o["refined"], o["best_img_kept"])
    assert (o["optimal_logit"] >= o["default_logit"]).all()


def test_batch_stat_bn_mode_differs_and_couples_samples():
    """The reference refines through D with is_training=True (GAN.py:175); the oracle exposes both modes."""
    arch = onets.get_arch("mnist")
    w = onets.scale_weights_for_signal(arch, onets.init_weights(arch, seed=1), 3.0)
    h0 = torch.relu(torch.randn(6, 7, 7, 128, generator=torch.Generator().manual_seed(0)))
    a = gr.build_refiner(h0, arch, w, 2, 0.1, d_bn="inference")
    b = gr.build_refiner(h0, arch, w, 2, 0.1, d_bn="batch")
    assert not torch.allclose(a["optimal_logit"], b["optimal_logit"])
    # inference mode: per-sample independence (what makes sharding exact)
    c = gr.build_refiner(h0[:3], arch, w, 2, 0.1, d_bn="inference")
    assert torch.allclose(c["refined"], a["refined"][:3], atol=1e-6)


def test_ladam_rejected_like_reference():
    arch = onets.get_arch("mnist")
    w = onets.init_weights(arch, seed=1)
    with pytest.raises(NotImplementedError):
        gr.build_refiner(torch.zeros(1, 7, 7, 128), arch, w, 1, 0.1, method="ladam")

"""Host-side lowering + weight packing, checked on the CPU against the oracle's TF-semantics layers."""
import numpy as np
import pytest
import torch

from oracle import nets as onets

import emulate


def _arches():
    return ["mnist", "dcgan32_l1", "dcgan64_l3"]


def test_arch_builders_agree(cgs_lib):
    from cgs import nets as N
    for name in _arches() + ["dcgan64_l1", "dcgan64_l4", "dcgan32_l2"]:
        assert N.get_arch(name) == onets.get_arch(name)


@pytest.mark.parametrize("arch_name", _arches())
def test_layer_forward_and_dgrad_lowering(cgs_lib, arch_name):
    from cgs import nets as N
    arch = N.get_arch(arch_name)
    w = onets.init_weights(arch, seed=7)
    rng = np.random.RandomState(0)
    B = 2
    for scope, layers in (("generator", arch["gtail"]), ("discriminator", arch["d"])):
        for layer in layers:
            if layer["type"] == "fc" and layer["cout"] == 1:
                continue
            wf, bf = N.fold_layer(layer, scope, w)
            w_fwd, w_bwd, bias = N.pack_layer(layer, wf, bf)
            cin, cout = layer["cin"], layer["cout"]
            if layer["type"] == "fc":
                x = rng.standard_normal((B, cin)).astype(np.float32)
            else:
                x = rng.standard_normal((B, layer["hin"], layer["win"], cin)).astype(np.float32)
            # oracle forward (pre-activation, BN inference) and its data-gradient
            xt = torch.from_numpy(x).requires_grad_(True)
            Lnoact = dict(layer, act="none")
            y = onets.run_layers(xt, [Lnoact], scope, w, "inference")
            dy = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32))
            (dx,) = torch.autograd.grad((y * dy).sum(), xt)
            # replay forward
            cs_in, cs_out = N.cstride(cin), N.cstride(cout)
            hin, win = (layer["hin"], layer["win"]) if layer["type"] != "fc" else (1, 1)
            xp = np.zeros((B, hin, win, cs_in))
            xp[..., :cin] = x.reshape(B, hin, win, cin)
            acc = emulate.layer_pass(layer, False, B, xp, w_fwd.numpy())
            got = acc[..., :cout] + bias.numpy()[:cout]
            ref = y.detach().numpy().reshape(got.shape)
            assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), (layer["name"], "fwd")
            assert np.all(acc[..., cout:] == 0)
            # replay backward
            hout, wout = (y.shape[1], y.shape[2]) if layer["type"] != "fc" else (1, 1)
            dyp = np.zeros((B, hout, wout, cs_out))
            dyp[..., :cout] = dy.numpy().reshape(B, hout, wout, cout)
            gacc = emulate.layer_pass(layer, True, B, dyp, w_bwd.numpy())
            gref = dx.numpy().reshape(B, hin, win, cin)
            assert np.abs(gacc[..., :cin] - gref).max() <= 1e-4 * max(1.0, np.abs(gref).max()), (layer["name"], "bwd")


def test_multiply_high_division_is_exact():
    """csrc/conv_gemm.cuh fast_div: n // d == (n * ceil(2^64 / d)) >> 64 for every 32-bit n (magic 0 encodes d == 1).
    The persistent tile index of every GEMM / edge kernel is decomposed with it, so it must be exact, not close."""
    import random
    rng = random.Random(7)
    ds = [1, 2, 3, 4, 7, 8, 14, 25, 49, 98, 147, 148, 196, 1024, 1568, 2048, 65535, 65536, 1 << 20, (1 << 31) - 1]
    ds += [rng.randrange(1, 1 << 31) for _ in range(200)]
    for d in ds:
        magic = 0 if d <= 1 else ((1 << 64) - 1) // d + 1
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 31) - 1, (1 << 32) - 1] + [rng.randrange(0, 1 << 32) for _ in range(200)]
        for n in ns:
            if n < 0 or n >= (1 << 32):
                continue
            got = n if magic == 0 else (n * magic) >> 64
            assert got == n // d, (n, d)

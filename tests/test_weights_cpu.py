"""Weight interchange on the CPU: Saver-style names, npz round trip, layout permutations of
nsgan/ops.py:39 ([kh,kw,Cin,Cout]) / :51 ([kh,kw,Cout,Cin]) / :75-79 ([in,out]) and the errors for bad files."""
import numpy as np
import pytest
import torch

from oracle import nets as onets


def _arch_and_weights(name="dcgan32_l2"):
    from cgs import nets as N
    arch = N.get_arch(name)
    return arch, onets.init_weights(arch, seed=4)


def test_npz_round_trip_with_saver_decorations(tmp_path):
    from cgs import weights as W
    arch, w = _arch_and_weights()
    decorated = {k + ":0": v for k, v in w.items()}
    decorated["generator/g_h2/w/Adam:0"] = np.zeros(3)
    decorated["generator/g_h2/w/Adam_1:0"] = np.zeros(3)
    decorated["beta1_power:0"] = np.float32(0.5)
    path = tmp_path / "ckpt.npz"
    W.save_npz(path, decorated)
    got = W.load_npz(path)
    assert set(got) == set(w) and all(np.array_equal(got[k], w[k]) for k in w)
    W.validate(arch, got, include_head=True)


def test_torch_layout_permutation_is_the_inverse_of_the_oracle_permutes():
    """conv: TF [kh,kw,Cin,Cout] <-> torch [Cout,Cin,kh,kw]; deconv: TF [kh,kw,Cout,Cin] <-> torch [Cin,Cout,kh,kw]
    (the permutes oracle/nets.py applies); a file in torch order gives the same nets."""
    from cgs import weights as W
    arch, w = _arch_and_weights()
    tw = W.to_torch_layout(arch, w, include_head=True)
    conv = next(L for L in arch["d"] if L["type"] == "conv")
    dec = arch["gtail"][0]
    wc, wd = w["discriminator/%s/w" % conv["name"]], w["generator/%s/w" % dec["name"]]
    assert np.array_equal(tw["discriminator/%s/w" % conv["name"]], torch.from_numpy(wc).permute(3, 2, 0, 1).numpy())
    assert np.array_equal(tw["generator/%s/w" % dec["name"]], torch.from_numpy(wd).permute(3, 2, 0, 1).numpy())
    assert tw["generator/%s/w" % dec["name"]].shape == (dec["cin"], dec["cout"], dec["k"], dec["k"])
    back = W.from_torch_layout(arch, tw, include_head=True)
    assert all(np.array_equal(back[k], w[k]) for k in w)


def test_validation_names_the_offending_variable():
    from cgs import weights as W
    arch, w = _arch_and_weights()
    missing = {k: v for k, v in w.items() if k != "discriminator/d_bn2/moving_variance"}
    with pytest.raises(KeyError, match="d_bn2/moving_variance"):
        W.validate(arch, missing)
    wrong = dict(w)
    k = "generator/%s/w" % arch["gtail"][0]["name"]
    wrong[k] = np.ascontiguousarray(np.transpose(w[k], (0, 1, 3, 2)))     # conv order where deconv order is expected
    with pytest.raises(ValueError, match="another order"):
        W.validate(arch, wrong)

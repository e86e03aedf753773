"""Pins the CPU oracle against golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import torch

from oracle import sudormrf_oracle as O

# The oracle issues the same torch CPU ops as the reference, so agreement is to
# fp32 reassociation noise (thread count / oneDNN version); SURVEY §8c measured
# the reference's own fp32 noise floor at 1.1e-6.
TOL = 2e-5


def test_state_dict_inventory(golden):
    meta, sd, x, outs, taps = golden
    cfg = O.Config(variant=meta["variant"], **meta["kwargs"])
    shapes = O.param_shapes(cfg)
    assert list(shapes.keys()) == list(sd.keys())
    for k, s in shapes.items():
        assert tuple(sd[k].shape) == tuple(s), k


def test_forward_matches_reference_golden(golden):
    meta, sd, x, outs, taps = golden
    cfg = O.Config(variant=meta["variant"], **meta["kwargs"])
    got_taps = {}
    y = O.forward(cfg, sd, x, taps=got_taps)
    assert y.shape == outs["output"].shape
    assert y.shape[-1] == x.shape[-1]         # groupcomm_sudormrf_v2.py:442
    rel_max, rel_l2 = O.parity_errors(y, outs["output"])
    assert rel_max < TOL and rel_l2 < TOL, (rel_max, rel_l2)
    # intermediate taps captured with forward hooks on the reference
    name_map = {"encoder": "encoder", "bottleneck": "bottleneck",
                "sm.0.proj_1x1.conv": "sm.0.proj_1x1.conv",
                "sm.0.spp_dw.0.conv": "sm.0.spp_dw.0.conv",
                "sm.0.spp_dw.1.conv": "sm.0.spp_dw.1.conv",
                "sm.0.UBlock.proj_1x1.conv": "sm.0.UBlock.proj_1x1.conv",
                "sm.0.UBlock.spp_dw.1.conv": "sm.0.UBlock.spp_dw.1.conv",
                "sm.0.TAC": "sm.0.TAC", "sm.0": "sm.0.out", "sm.1": "sm.1.out",
                "l1": "l1", "sm.0.conv_1x1_exp.conv": "sm.0.conv_1x1_exp.conv", "m": "m"}
    checked = 0
    for ref_name, ours in name_map.items():
        if ref_name in taps:
            if meta["variant"] == "groupcomm" and ours.endswith(".out"):
                ours = ours.replace(".out", ".UBlock.out")
            a = got_taps[ours].reshape(taps[ref_name].shape)
            e = O.parity_errors(a, taps[ref_name])
            assert max(e) < TOL, (ref_name, e)
            checked += 1
    assert checked >= 5


def test_mixture_consistency_golden(golden):
    meta, sd, x, outs, taps = golden
    if "mc_uniform" not in outs:
        return
    y = outs["output"]
    assert max(O.parity_errors(O.mixture_consistency(y, x), outs["mc_uniform"])) < 1e-6
    assert max(O.parity_errors(O.mixture_consistency(y, x, "magsq"), outs["mc_magsq"])) < 1e-6
    # defining property: the corrected estimates sum to the mixture
    s = O.mixture_consistency(y, x).sum(1, keepdim=True)
    assert torch.allclose(s, x, atol=1e-5)


def test_fp64_oracle_agrees_with_fp32(golden):
    meta, sd, x, outs, taps = golden
    cfg = O.Config(variant=meta["variant"], **meta["kwargs"])
    y64 = O.forward(cfg, sd, x, dtype=torch.float64)
    assert max(O.parity_errors(y64, outs["output"])) < 1e-4


def test_padded_length_rule():
    cfg = O.Config(upsampling_depth=5)            # hop*2^D = 320
    assert O.padded_length(cfg, 1) == 320
    assert O.padded_length(cfg, 320) == 320
    assert O.padded_length(cfg, 321) == 640
    assert O.padded_length(cfg, 32000) == 32000
    assert O.padded_length(cfg, 32079) == 32320


def test_bad_mc_type_raises():
    import pytest
    with pytest.raises(ValueError):
        O.mixture_consistency(torch.zeros(1, 2, 4), torch.zeros(1, 1, 4), "nope")

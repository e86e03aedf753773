"""Live comparison oracle <-> unmodified reference.  Only runs where
/root/reference exists (the build container); the GPU box relies on the
committed golden fixtures instead."""
import os
import sys
import warnings

import pytest
import torch

from oracle import sudormrf_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sudo_rm_rf")),
                                reason="reference tree not present")


def _ref_modules():
    warnings.filterwarnings("ignore")
    # our drop-in overlay must not shadow the reference here
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] == "sudo_rm_rf"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        import sudo_rm_rf.dnn.models.improved_sudormrf as ri
        import sudo_rm_rf.dnn.models.groupcomm_sudormrf_v2 as rg
        import sudo_rm_rf.dnn.models.causal_improved_sudormrf_v3 as rc
        import sudo_rm_rf.dnn.models.sudormrf as ro
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split(".")[0] == "sudo_rm_rf"]:
            del sys.modules[k]
        sys.modules.update(saved)
    return {"improved": ri.SuDORMRF, "groupcomm": rg.GroupCommSudoRmRf, "causal": rc.CausalSuDORMRF,
            "original": ro.SuDORMRF}


@pytest.mark.parametrize("variant,kw,T", [
    ("improved", dict(out_channels=64, in_channels=128, num_blocks=4, upsampling_depth=5,
                      enc_kernel_size=21, enc_num_basis=128, num_sources=2), 3333),
    ("improved", dict(out_channels=32, in_channels=32, num_blocks=1, upsampling_depth=1,
                      enc_kernel_size=21, enc_num_basis=16, num_sources=1), 7),
    ("groupcomm", dict(out_channels=64, in_channels=128, num_blocks=2, upsampling_depth=4,
                       enc_kernel_size=21, enc_num_basis=64, num_sources=2,
                       group_size=16), 2000),
    ("causal", dict(in_audio_channels=1, out_channels=32, in_channels=64, num_blocks=3, upsampling_depth=4,
                    enc_kernel_size=21, enc_num_basis=64, num_sources=2), 2000),
    ("original", dict(out_channels=64, in_channels=128, num_blocks=3, upsampling_depth=4,
                      enc_kernel_size=21, enc_num_basis=128, num_sources=2), 3333),
    ("original", dict(out_channels=32, in_channels=64, num_blocks=2, upsampling_depth=5,
                      enc_kernel_size=21, enc_num_basis=32, num_sources=4), 1600),      # no reshape layer, T a multiple of the lcm
])
def test_live(variant, kw, T):
    cfg = O.Config(variant=variant, **kw)
    sd = O.make_state_dict(cfg, seed=11)
    m = _ref_modules()[variant](**kw).eval()
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    x = torch.randn(2, 1, T, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = m(x)
    assert max(O.parity_errors(O.forward(cfg, sd, x), ref)) < 2e-5

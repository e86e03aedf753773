"""The reference's import paths resolve to the B200-native modules through the overlay, and a
whole-module pickle written against those paths loads (README.md:75 `torch.load(model.pt)`)."""
import io
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import io, sys, torch
sys.path.insert(0, {repo!r}); sys.path.insert(0, {repo!r} + "/sudo_rm_rf_b200/dropin")
import sudo_rm_rf.dnn.models.improved_sudormrf as improved_sudormrf
import sudo_rm_rf.dnn.models.groupcomm_sudormrf_v2 as sudormrf_gc_v2
import sudo_rm_rf.dnn.experiments.utils.mixture_consistency as mixture_consistency
import sudo_rm_rf_b200 as P
assert improved_sudormrf.SuDORMRF is P.SuDORMRF
assert sudormrf_gc_v2.GroupCommSudoRmRf is P.GroupCommSudoRmRf
assert mixture_consistency.apply is P.mixture_consistency.apply
m = improved_sudormrf.SuDORMRF(16, 32, 2, 3, 21, 24, 2)
m.__class__.__module__ = "sudo_rm_rf.dnn.models.improved_sudormrf"   # what a reference pickle records
buf = io.BytesIO(); torch.save(m, buf); buf.seek(0)
m2 = torch.load(buf, weights_only=False)
assert type(m2).__name__ == "SuDORMRF" and m2.enc_num_basis == 24
assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
print("overlay ok")
"""


def test_overlay_import_paths_and_pickle():
    out = subprocess.run([sys.executable, "-c", SCRIPT.format(repo=REPO)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "overlay ok" in out.stdout

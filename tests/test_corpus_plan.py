"""Host logic of the bucketed corpus inference (sudo_rm_rf_b200/corpus.py): CPU only."""
import random

import pytest

from oracle import sudormrf_oracle as O
from sudo_rm_rf_b200.corpus import padded_length, plan_buckets


def test_padded_length_is_the_reference_rule():
    for depth in (1, 3, 5, 6):
        for k in (11, 21, 41):
            cfg = O.Config(enc_kernel_size=k, upsampling_depth=depth)
            q = (k // 2) * 2 ** depth
            for T in (1, 7, q - 1, q, q + 1, 3 * q, 3 * q + 5, 32000, 32079):
                assert padded_length(T, q) == O.padded_length(cfg, T)
    with pytest.raises(ValueError):
        padded_length(0, 320)


def test_plan_covers_every_utterance_once_and_buckets_share_a_padded_length():
    rnd = random.Random(3)
    lengths = [rnd.randint(1, 40000) for _ in range(500)] + [320, 640, 640, 319, 321]
    plan = plan_buckets(lengths, 320, 32)
    seen = []
    last_tp = 0
    for tp, idx in plan:
        assert 1 <= len(idx) <= 32
        assert tp % 320 == 0 and tp >= last_tp
        last_tp = tp
        assert all(padded_length(lengths[i], 320) == tp for i in idx)
        assert idx == sorted(idx)                      # corpus order inside a batch
        seen += idx
    assert sorted(seen) == list(range(len(lengths)))


def test_plan_splits_large_buckets_and_is_deterministic():
    lengths = [1000] * 70 + [5] * 3
    plan = plan_buckets(lengths, 320, 32)
    assert [(tp, len(idx)) for tp, idx in plan] == [(320, 3), (1280, 32), (1280, 32), (1280, 6)]
    assert plan == plan_buckets(lengths, 320, 32)
    assert plan_buckets([], 320, 8) == []
    with pytest.raises(ValueError):
        plan_buckets([10], 320, 0)


def test_wav_roundtrip(tmp_path):
    """load_wav follows torchaudio.load's conventions ([channels, T] float32 in [-1, 1], sample rate)."""
    import numpy as np
    import torch
    from scipy.io import wavfile
    from sudo_rm_rf_b200 import corpus as Cp
    x = (torch.randn(2, 1234) * 0.3).clamp(-1, 1)
    p = str(tmp_path / "f32.wav")
    Cp.save_wav(p, x, 8000)
    y, sr = Cp.load_wav(p)
    assert sr == 8000 and torch.equal(x, y)
    p16 = str(tmp_path / "i16.wav")
    wavfile.write(p16, 16000, (x[0].numpy() * 32767).astype(np.int16))
    y, sr = Cp.load_wav(p16)
    assert sr == 16000 and y.shape == (1, 1234) and float((y[0] - x[0]).abs().max()) < 1e-4


def test_plan_with_a_model_specific_padding_rule():
    """The original model pads to multiples of lcm(hop, 2^D) and leaves exact multiples alone (sudormrf.py:283-293)."""
    cfg = O.Config(variant="original", enc_kernel_size=21, upsampling_depth=4)
    rule = lambda T: O.padded_length(cfg, T)
    lengths = [1, 79, 80, 81, 160, 161, 4000, 4001]
    plan = plan_buckets(lengths, rule, 8)
    assert [(tp, idx) for tp, idx in plan] == [(80, [0, 1, 2]), (160, [3, 4]), (240, [5]), (4000, [6]), (4080, [7])]
    with pytest.raises(ValueError):
        plan_buckets([5, 0], rule, 8)

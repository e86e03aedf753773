"""Bucketed inference over a corpus of variable-length mixtures.

The reference evaluates a corpus one utterance at a time
(``sudo_rm_rf/utils/simple_whamr_evaluation.py:138-148``: normalise, ``model(x.unsqueeze(1))``,
mixture consistency).  Utterances cannot simply be stacked: the model pads every input to a
multiple of ``hop * 2**upsampling_depth`` samples (improved_sudormrf.py:303-314) and its global
layer norms see that padded length, so the estimate of an utterance depends on ITS padded length.
Utterances that share a padded length, however, are computed identically alone or side by side
(every reduction of the forward is per sample).  ``separate_corpus`` therefore buckets the corpus
by padded length, runs each bucket as zero-padded batches through ``sdr_separate_ragged``
(per-utterance statistics over the true length) and crops: every result equals what the
one-at-a-time loop of the reference produces for that utterance.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence, Tuple

import torch

from . import _engine
from . import _native as N


def padded_length(T: int, quantum: int) -> int:
    """improved_sudormrf.py:303-310 with quantum = hop * 2**upsampling_depth (= n_least_samples_req)."""
    if T <= 0:
        raise ValueError("empty utterance")
    return quantum if T < quantum else -(-T // quantum) * quantum


def plan_buckets(lengths: Sequence[int], quantum: int, max_batch: int) -> List[Tuple[int, List[int]]]:
    """Groups utterance indices into batches that share a padded length.

    Returns ``[(padded_length, [indices...]), ...]``: buckets in increasing padded length, inside a
    bucket the corpus order is kept and batches hold at most ``max_batch`` utterances.  Pure host
    logic (deterministic, no torch)."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    buckets = {}
    for i, T in enumerate(lengths):
        buckets.setdefault(padded_length(int(T), quantum), []).append(i)
    plan = []
    for Tp in sorted(buckets):
        idx = buckets[Tp]
        for k in range(0, len(idx), max_batch):
            plan.append((Tp, idx[k:k + max_batch]))
    return plan


def separate_corpus(model, wavs: Iterable[torch.Tensor], max_batch: int = 32,
                    mixture_consistency: bool = False, rescale: bool = True) -> List[torch.Tensor]:
    """Separates a corpus of mono mixtures of different lengths.

    ``wavs``: 1-D tensors (CPU or CUDA, any float dtype).  Returns, in corpus order, one ``[S, T_i]``
    fp32 CUDA tensor per utterance = ``model.separate(w[None], normalize=True, ...)[0]``.
    ``rescale=False`` returns the estimates of the normalised mixture, as the reference's evaluation
    script scores them (simple_whamr_evaluation.py:141-148); ``mixture_consistency`` as in
    README.md:113-114."""
    wavs = list(wavs)
    if not wavs:
        return []
    for w in wavs:
        if w.dim() != 1:
            raise RuntimeError("separate_corpus expects 1-D waveforms")
    lib = N.lib()
    cfg = _engine.make_config(model)
    if cfg.in_audio_channels != 1:
        raise RuntimeError("separate_corpus follows the README recipe, which is written for mono mixtures")
    device = _engine._fetch(model, "encoder.weight").device
    if device.type != "cuda":
        raise RuntimeError("sudo_rm_rf_b200 runs on CUDA (sm_100a) only: move the model to a B200")
    if torch.is_grad_enabled() and model.training:
        raise RuntimeError("sudo_rm_rf_b200 implements the inference forward only: call model.eval()")
    quantum = (cfg.enc_kernel_size // 2) * (2 ** cfg.upsampling_depth)
    plan = plan_buckets([int(w.shape[0]) for w in wavs], quantum, max_batch)
    results: List[torch.Tensor] = [None] * len(wavs)
    S = cfg.num_sources
    with torch.cuda.device(device), torch.no_grad():
        packed = _engine.packed_weights(model, cfg, device)
        st = _engine._state(model, device)
        for Tp, idx in plan:
            B = len(idx)
            batch = torch.zeros((B, 1, Tp), dtype=torch.float32, device=device)
            for r, i in enumerate(idx):
                batch[r, 0, :wavs[i].shape[0]] = wavs[i].detach().to(device=device, dtype=torch.float32)
            lengths = torch.tensor([int(wavs[i].shape[0]) for i in idx], dtype=torch.int64, device=device)
            ws_bytes = lib.sdr_separate_workspace_bytes(C.byref(cfg), B, Tp)
            if ws_bytes == 0:
                raise N.NativeError("bad model configuration (sdr_separate_workspace_bytes returned 0)")
            if st.workspace is None or st.workspace.numel() < ws_bytes:
                st.workspace = None
                st.graphs.clear()
                st.workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
            out = torch.empty((B, S, Tp), dtype=torch.float32, device=device)
            N.check(lib.sdr_separate_ragged(
                C.byref(cfg), C.c_void_p(packed.data_ptr()), C.c_void_p(batch.data_ptr()),
                C.c_void_p(lengths.data_ptr()), C.c_void_p(out.data_ptr()), B, Tp,
                1 if mixture_consistency else 0, 1 if rescale else 0,
                C.c_void_p(st.workspace.data_ptr()), st.workspace.numel(),
                _engine._stream_ptr(device)), "sdr_separate_ragged")
            for r, i in enumerate(idx):
                results[i] = out[r, :, :wavs[i].shape[0]].clone()
    return results

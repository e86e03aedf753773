"""B200-native mirror of the evaluation metric in ``sudo_rm_rf/dnn/losses/sisdr.py``.

``PermInvariantSISDR`` keeps the reference's constructor arguments, ``forward``
signature and return conventions (sisdr.py:66-194), but it is the *metric* the
validation loops compute right after ``model(...)``
(run_improved_sudormrf.py:82-85,201-205): inference only, no autograd.  The
arithmetic (one fp64 Gram pass over the batch + a permutation search per item)
runs in ``libsudormrf_b200.so`` (``sdr_pit_sisdr``).
"""
import ctypes as C
import itertools

import torch
import torch.nn as nn

from . import _native as N


class PermInvariantSISDR(nn.Module):
    """Permutation-invariant SI-SDR / SI-SDRi of a batch of estimates (sisdr.py:66-194)."""

    def __init__(self, batch_size=None, zero_mean=False, n_sources=None, backward_loss=True,
                 improvement=False, return_individual_results=False):
        super().__init__()
        self.bs = batch_size
        self.perform_zero_mean = zero_mean
        self.backward_loss = backward_loss
        self.permutations = list(itertools.permutations(torch.arange(n_sources)))
        self.permutations_tensor = torch.LongTensor(self.permutations)
        self.improvement = improvement
        self.n_sources = n_sources
        self.return_individual_results = return_individual_results

    def forward(self, pr_batch, t_batch, eps=1e-9, initial_mixtures=None,
                return_best_permutation=False):
        """pr_batch, t_batch ``[B, n_sources, T]``; initial_mixtures ``[B, 1, T]`` (SI-SDRi only).
        Returns what the reference returns: the (negated if ``backward_loss``) best SI-SDR(i), per
        item if ``return_individual_results`` else its batch mean, and optionally the best
        permutations ``[B, n_sources]``."""
        if pr_batch.dim() != 3 or t_batch.dim() != 3 or pr_batch.shape[:2] != t_batch.shape[:2] \
                or pr_batch.shape[1] != self.n_sources:
            raise RuntimeError("expected pr_batch and t_batch of shape [B, n_sources, T]")
        if not (pr_batch.is_cuda and t_batch.is_cuda):
            raise RuntimeError("sudo_rm_rf_b200.sisdr runs on CUDA tensors only (no CPU path)")
        if torch.is_grad_enabled() and (pr_batch.requires_grad or t_batch.requires_grad):
            raise RuntimeError("sudo_rm_rf_b200.sisdr is the evaluation metric only (no autograd): "
                               "wrap the call in torch.no_grad()")
        if self.improvement and initial_mixtures is None:
            raise RuntimeError("improvement=True needs initial_mixtures")
        # normalize_input (sisdr.py:95-112): crop everything to the shortest length
        min_len = min(pr_batch.shape[-1], t_batch.shape[-1])
        if initial_mixtures is not None:
            min_len = min(min_len, initial_mixtures.shape[-1])
        dev = pr_batch.device
        est = pr_batch.detach()[:, :, :min_len].to(torch.float32).contiguous()
        tgt = t_batch.detach()[:, :, :min_len].to(device=dev, dtype=torch.float32).contiguous()
        mix = None
        if initial_mixtures is not None:
            if initial_mixtures.dim() != 3 or initial_mixtures.shape[1] != 1 \
                    or initial_mixtures.shape[0] != est.shape[0]:
                raise RuntimeError("expected initial_mixtures of shape [B, 1, T]")
            mix = initial_mixtures.detach()[:, :, :min_len].to(device=dev, dtype=torch.float32).contiguous()
        B, S, T = est.shape
        lib = N.lib()
        nbytes = lib.sdr_pit_sisdr_scratch_bytes(B, S)
        if nbytes == 0:
            raise N.NativeError("sdr_pit_sisdr supports 1..4 sources")
        with torch.cuda.device(dev):
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            best = torch.empty(B, dtype=torch.float32, device=dev)
            perm = torch.empty(B, dtype=torch.int32, device=dev)
            N.check(lib.sdr_pit_sisdr(
                C.c_void_p(est.data_ptr()), C.c_void_p(tgt.data_ptr()),
                C.c_void_p(mix.data_ptr() if (mix is not None and self.improvement) else 0),
                C.c_void_p(best.data_ptr()), C.c_void_p(perm.data_ptr()), B, S, T,
                1 if self.perform_zero_mean else 0, 1 if self.improvement else 0, float(eps),
                C.c_void_p(scratch.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "sdr_pit_sisdr")
        result = best if self.return_individual_results else best.mean()
        if self.backward_loss:
            result = -result
        if return_best_permutation:
            return result, self.permutations_tensor.to(dev)[perm.long()]
        return result

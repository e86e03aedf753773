"""CPU oracle for the SuDoRM-RF forward inference path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``sudo_rm_rf_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs use it, and only as the checker or the
timed CPU baseline, never as the product path.

It is a *functional restatement* (plain functions over a ``state_dict``; no
``nn.Module``) of the reference's forward pass.  Citations are file:line in the
reference tree (``/root/reference``):

  improved_sudormrf.py   = sudo_rm_rf/dnn/models/improved_sudormrf.py
  groupcomm_sudormrf_v2.py = sudo_rm_rf/dnn/models/groupcomm_sudormrf_v2.py
  causal_improved_sudormrf_v3.py = sudo_rm_rf/dnn/models/causal_improved_sudormrf_v3.py
  sudormrf.py            = sudo_rm_rf/dnn/models/sudormrf.py (the original SuDoRM-RF, variant "original")
  mixture_consistency.py = sudo_rm_rf/dnn/experiments/utils/mixture_consistency.py
  README.md              = the reference's README (inference recipe, lines 100-114)
  sisdr.py               = sudo_rm_rf/dnn/losses/sisdr.py (validation metrics)

Parity pinning: the reference ships NO golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``tests/golden/make_golden.py`` (which imports the
unmodified reference from /root/reference) and committed as
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them, and
``tests/test_oracle_vs_reference.py`` compares live when /root/reference exists.

All arithmetic is torch CPU, fp32 by default (the reference's dtype); pass
``dtype=torch.float64`` for a high-precision run of the same formulas.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Config:
    """Constructor arguments of the reference models.

    improved_sudormrf.py:224-231 (SuDORMRF),
    groupcomm_sudormrf_v2.py:232-241 (GroupCommSudoRmRf) and
    causal_improved_sudormrf_v3.py:121-129 (CausalSuDORMRF),
    sudormrf.py:186-193 (the original SuDORMRF).
    """
    variant: str = "improved"          # "improved" | "groupcomm" | "causal" | "original"
    out_channels: int = 128
    in_channels: int = 512
    num_blocks: int = 16
    upsampling_depth: int = 4
    enc_kernel_size: int = 21
    enc_num_basis: int = 512
    num_sources: int = 2
    in_audio_channels: int = 1         # groupcomm and causal only
    group_size: int = 16               # groupcomm only

    @property
    def hop(self) -> int:
        return self.enc_kernel_size // 2

    @property
    def n_least_samples_req(self) -> int:
        # improved_sudormrf.py:244
        return self.hop * 2 ** self.upsampling_depth

    def as_dict(self):
        return asdict(self)


def padded_length(cfg: Config, T: int) -> int:
    """improved_sudormrf.py:303-310: round T up to a multiple of hop*2^D
    (at least one multiple).  The original model (sudormrf.py:206-209,283-293) rounds up to a multiple of
    lcm(hop, 2^D) instead and leaves a length that already is one alone."""
    if cfg.variant == "original":
        q = cfg.hop * 2 ** cfg.upsampling_depth // math.gcd(cfg.hop, 2 ** cfg.upsampling_depth)
        return T if T % q == 0 else T + q - T % q
    q = cfg.n_least_samples_req
    if T < q:
        return q
    return ((T + q - 1) // q) * q


# --------------------------------------------------------------------------
# parameter inventory (names/shapes are API: SURVEY §8 a1/a5/a8)
# --------------------------------------------------------------------------
def _ublock_shapes(prefix: str, co: int, ci: int, depth: int) -> Dict[str, tuple]:
    # improved_sudormrf.py:170-196
    s = {
        f"{prefix}proj_1x1.conv.weight": (ci, co, 1),
        f"{prefix}proj_1x1.conv.bias": (ci,),
        f"{prefix}proj_1x1.norm.gamma": (ci,),
        f"{prefix}proj_1x1.norm.beta": (ci,),
        f"{prefix}proj_1x1.act.weight": (1,),
    }
    for d in range(depth):
        s[f"{prefix}spp_dw.{d}.conv.weight"] = (ci, 1, 5)
        s[f"{prefix}spp_dw.{d}.conv.bias"] = (ci,)
        s[f"{prefix}spp_dw.{d}.norm.gamma"] = (ci,)
        s[f"{prefix}spp_dw.{d}.norm.beta"] = (ci,)
    s[f"{prefix}final_norm.norm.gamma"] = (ci,)
    s[f"{prefix}final_norm.norm.beta"] = (ci,)
    s[f"{prefix}final_norm.act.weight"] = (1,)
    s[f"{prefix}res_conv.weight"] = (co, ci, 1)
    s[f"{prefix}res_conv.bias"] = (co,)
    return s


def param_shapes(cfg: Config) -> Dict[str, tuple]:
    """Ordered name -> shape map, in the reference's ``state_dict()`` order."""
    N, Co, Ci = cfg.enc_num_basis, cfg.out_channels, cfg.in_channels
    S, D, k = cfg.num_sources, cfg.upsampling_depth, cfg.enc_kernel_size
    if cfg.variant == "causal":
        return _causal_param_shapes(cfg)
    if cfg.variant == "original":
        return _original_param_shapes(cfg)
    A = cfg.in_audio_channels if cfg.variant == "groupcomm" else 1
    s: Dict[str, tuple] = {}
    s["encoder.weight"] = (N, A, k)                  # improved_sudormrf.py:247-251
    s["ln.gamma"] = (N,)
    s["ln.beta"] = (N,)
    s["bottleneck.weight"] = (Co, N, 1)
    s["bottleneck.bias"] = (Co,)
    for i in range(cfg.num_blocks):
        if cfg.variant == "improved":
            s.update(_ublock_shapes(f"sm.{i}.", Co, Ci, D))
        else:
            G = cfg.group_size
            n, H = Co // G, Co * 3 // G              # groupcomm_sudormrf_v2.py:402
            s[f"sm.{i}.TAC.TAC_input.0.weight"] = (H, n)
            s[f"sm.{i}.TAC.TAC_input.0.bias"] = (H,)
            s[f"sm.{i}.TAC.TAC_input.1.weight"] = (1,)
            s[f"sm.{i}.TAC.TAC_mean.0.weight"] = (H, H)
            s[f"sm.{i}.TAC.TAC_mean.0.bias"] = (H,)
            s[f"sm.{i}.TAC.TAC_mean.1.weight"] = (1,)
            s[f"sm.{i}.TAC.TAC_output.0.weight"] = (n, 2 * H)
            s[f"sm.{i}.TAC.TAC_output.0.bias"] = (n,)
            s[f"sm.{i}.TAC.TAC_output.1.weight"] = (1,)
            s[f"sm.{i}.TAC.TAC_norm.gamma"] = (n,)
            s[f"sm.{i}.TAC.TAC_norm.beta"] = (n,)
            s.update(_ublock_shapes(f"sm.{i}.UBlock.", Co // G, Ci // G, D))
    s["mask_net.0.weight"] = (1,)
    s["mask_net.1.weight"] = (S * N * A, Co, 1)
    s["mask_net.1.bias"] = (S * N * A,)
    s["decoder.weight"] = (N * S * A, S * A, k)     # improved_sudormrf.py:272-279
    return s


def _causal_param_shapes(cfg: Config) -> Dict[str, tuple]:
    """causal_improved_sudormrf_v3.py:146-197 (model), :71-95 (block): no norms, a scalar ``skipinit_gain`` per
    block, one PReLU slope per ConvAct, 21-tap depthwise filters, a (2k-1)-tap encoder, a PReLU on the masks."""
    N, Co, Ci = cfg.enc_num_basis, cfg.out_channels, cfg.in_channels
    S, D, k, A = cfg.num_sources, cfg.upsampling_depth, cfg.enc_kernel_size, cfg.in_audio_channels
    s: Dict[str, tuple] = {}
    s["encoder.weight"] = (N, A, 2 * k - 1)          # :146-151
    s["bottleneck.weight"] = (Co, N, 1)              # :155-158
    s["bottleneck.bias"] = (Co,)
    for i in range(cfg.num_blocks):
        p = f"sm.{i}."
        s[p + "skipinit_gain"] = ()                  # :73
        s[p + "proj_1x1.conv.weight"] = (Ci, Co, 1)
        s[p + "proj_1x1.conv.bias"] = (Ci,)
        s[p + "proj_1x1.act.weight"] = (1,)
        for d in range(D):
            s[p + f"spp_dw.{d}.conv.weight"] = (Ci, 1, 21)     # :78-90
            s[p + f"spp_dw.{d}.conv.bias"] = (Ci,)
            s[p + f"spp_dw.{d}.act.weight"] = (1,)
        s[p + "res_conv.weight"] = (Co, Ci, 1)       # :96
        s[p + "res_conv.bias"] = (Co,)
    s["mask_net.0.weight"] = (1,)                    # :175-177
    s["mask_net.1.weight"] = (S * N * A, Co, 1)
    s["mask_net.1.bias"] = (S * N * A,)
    s["decoder.weight"] = (N * S * A, S * A, k)      # :180-187
    s["mask_nl_class.weight"] = (1,)                 # :189
    return s


def _original_param_shapes(cfg: Config) -> Dict[str, tuple]:
    """sudormrf.py:211-264 (model), :134-157 (UBlock): GroupNorm(1, C) norms (``weight`` / ``bias``), one PReLU
    slope per channel, a biased encoder and decoder, a ``conv_1x1_exp`` + ``module_act`` tail in every block, the
    (N+1) x 1 Conv2d that produces the mask logits, a grouped decoder, and ``ln_mask_in`` (registered last, :264,
    never used by forward)."""
    N, Co, Ci = cfg.enc_num_basis, cfg.out_channels, cfg.in_channels
    S, D, k = cfg.num_sources, cfg.upsampling_depth, cfg.enc_kernel_size
    s: Dict[str, tuple] = {}
    s["encoder.0.weight"] = (N, 1, k)                # :212-218
    s["encoder.0.bias"] = (N,)
    s["ln.weight"] = (N,)                            # :221
    s["ln.bias"] = (N,)
    s["l1.weight"] = (Co, N, 1)                      # :222-224
    s["l1.bias"] = (Co,)
    for i in range(cfg.num_blocks):
        p = f"sm.{i}."
        s[p + "proj_1x1.conv.weight"] = (Ci, Co, 1)  # :138-139
        s[p + "proj_1x1.conv.bias"] = (Ci,)
        s[p + "proj_1x1.norm.weight"] = (Ci,)
        s[p + "proj_1x1.norm.bias"] = (Ci,)
        s[p + "proj_1x1.act.weight"] = (Ci,)
        for d in range(D):                           # :141-154
            s[p + f"spp_dw.{d}.conv.weight"] = (Ci, 1, 5)
            s[p + f"spp_dw.{d}.conv.bias"] = (Ci,)
            s[p + f"spp_dw.{d}.norm.weight"] = (Ci,)
            s[p + f"spp_dw.{d}.norm.bias"] = (Ci,)
        s[p + "conv_1x1_exp.conv.weight"] = (Co, Ci, 1)   # :160
        s[p + "conv_1x1_exp.conv.bias"] = (Co,)
        s[p + "conv_1x1_exp.norm.weight"] = (Co,)
        s[p + "conv_1x1_exp.norm.bias"] = (Co,)
        s[p + "final_norm.norm.weight"] = (Ci,)      # :161
        s[p + "final_norm.norm.bias"] = (Ci,)
        s[p + "final_norm.act.weight"] = (Ci,)
        s[p + "module_act.norm.weight"] = (Co,)      # :162
        s[p + "module_act.norm.bias"] = (Co,)
        s[p + "module_act.act.weight"] = (Co,)
    if Co != N:                                      # :233-236
        s["reshape_before_masks.weight"] = (N, Co, 1)
        s["reshape_before_masks.bias"] = (N,)
    s["m.weight"] = (S, 1, N + 1, 1)                 # :239-242
    s["m.bias"] = (S,)
    s["decoder.weight"] = (N * S, 1, k)              # :245-252 (groups = S)
    s["decoder.bias"] = (S,)
    s["ln_mask_in.weight"] = (N,)                    # :253
    s["ln_mask_in.bias"] = (N,)
    return s


def _is_groupnorm(name: str, leaf: str) -> bool:
    """GroupNorm parameters of the original model (``weight`` = gamma, ``bias`` = beta)."""
    return name.endswith("norm." + leaf) or name in ("ln." + leaf, "ln_mask_in." + leaf)


def make_state_dict(cfg: Config, seed: int = 0, perturbed: bool = True,
                    dtype=torch.float32) -> Dict[str, Tensor]:
    """Seeded synthetic weights with the reference's names and shapes.

    ``perturbed=True``: every parameter is non-trivial so that
    gamma/beta/bias/PReLU slopes are exercised (the reference's default init has
    gamma=1, beta=0, slope=.25 which hides bugs).  Matrices are scaled by
    1/sqrt(fan_in) so activations stay O(1) through deep stacks.
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("skipinit_gain"):
            # the reference initialises it to 0 (every block the identity); the synthetic weights use a gain that keeps
            # the residual stream O(1) while making every block matter
            t = 0.5 + 0.2 * torch.rand(shape, generator=g) if perturbed else torch.zeros(shape)
        elif not perturbed:
            if name.endswith("gamma") or _is_groupnorm(name, "weight"):
                t = torch.ones(shape)
            elif name.endswith("beta") or _is_groupnorm(name, "bias"):
                t = torch.zeros(shape)
            elif name.endswith("act.weight") or name in ("mask_net.0.weight", "mask_nl_class.weight") or \
                    (name.endswith(".1.weight") and "TAC" in name):
                t = torch.full(shape, 0.25)
            else:
                fan_in = max(1, int(math.prod(shape[1:])))
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
        else:
            if name.endswith("gamma") or _is_groupnorm(name, "weight"):
                t = 1.0 + 0.3 * torch.randn(shape, generator=g)
            elif name.endswith("beta") or _is_groupnorm(name, "bias"):
                t = 0.2 * torch.randn(shape, generator=g)
            elif name.endswith("act.weight") and len(shape) == 1 and shape[0] > 1:
                # per-channel PReLU slopes: most in [0.1, 0.6), about a tenth negative and a seventh above 1
                r = torch.rand(shape, generator=g)
                t = 0.1 + 0.5 * torch.rand(shape, generator=g)
                t = torch.where(r < 0.1, torch.full_like(t, -0.2), torch.where(r > 0.85, torch.full_like(t, 1.3), t))
            elif len(shape) == 1 and shape[0] == 1:
                t = 0.25 + 0.15 * torch.rand(shape, generator=g)   # PReLU slopes
            elif len(shape) == 1:
                t = 0.1 * torch.randn(shape, generator=g)          # biases
            else:
                fan_in = max(1, int(math.prod(shape[1:])))
                t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
        sd[name] = t.to(dtype)
    return sd


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def glob_ln(x: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    """Global layer norm, improved_sudormrf.py:30-47: biased two-pass variance
    over every non-batch dim, eps=1e-8 inside the sqrt, per-channel affine."""
    dims = tuple(range(1, x.dim()))
    mu = x.mean(dim=dims, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=dims, keepdim=True)
    xn = (x - mu) / torch.sqrt(var + 1e-8)
    shape = [1, -1] + [1] * (x.dim() - 2)
    return xn * gamma.view(shape) + beta.view(shape)


def prelu1(x: Tensor, slope: Tensor) -> Tensor:
    """nn.PReLU() with its single shared slope (improved_sudormrf.py:68,111)."""
    return torch.where(x >= 0, x, x * slope.reshape(()))


def pad_wave(cfg: Config, wav: Tensor, dtype) -> Tensor:
    """improved_sudormrf.py:303-314: zero-pad on the right to padded_length;
    the reference materialises the pad in fp32."""
    T = wav.shape[-1]
    Tp = padded_length(cfg, T)
    out = torch.zeros(list(wav.shape[:-1]) + [Tp], dtype=dtype, device=wav.device)
    out[..., :T] = wav.to(dtype)
    return out


def uconv_block(x: Tensor, sd: Dict[str, Tensor], p: str, depth: int,
                taps: Optional[dict] = None) -> Tensor:
    """One U-ConvBlock, improved_sudormrf.py:198-220.

    y   = PReLU(GLN(W1 x + b1))                               (:205, ConvNormAct :70-73)
    z_0 = GLN(dw5_s1(y)), z_d = GLN(dw5_s2(z_{d-1}))          (:206-211)
    m   = z_0 + up2(z_1 + up2(z_2 + ...))  (nearest)          (:214-216)
    out = W2 PReLU(GLN(m)) + b2 + x                           (:218-220)
    """
    ci = sd[p + "proj_1x1.conv.weight"].shape[0]
    y = F.conv1d(x, sd[p + "proj_1x1.conv.weight"], sd[p + "proj_1x1.conv.bias"])
    if taps is not None:
        taps[p + "proj_1x1.conv"] = y
    y = glob_ln(y, sd[p + "proj_1x1.norm.gamma"], sd[p + "proj_1x1.norm.beta"])
    y = prelu1(y, sd[p + "proj_1x1.act.weight"])
    levels = []
    cur = y
    for d in range(depth):
        stride = 1 if d == 0 else 2                             # :181-189
        z = F.conv1d(cur, sd[p + f"spp_dw.{d}.conv.weight"],
                     sd[p + f"spp_dw.{d}.conv.bias"],
                     stride=stride, padding=2, groups=ci)
        if taps is not None:
            taps[p + f"spp_dw.{d}.conv"] = z
        cur = glob_ln(z, sd[p + f"spp_dw.{d}.norm.gamma"],
                      sd[p + f"spp_dw.{d}.norm.beta"])
        levels.append(cur)
    for _ in range(depth - 1):
        top = levels.pop()
        levels[-1] = levels[-1] + F.interpolate(top, scale_factor=2, mode="nearest")
    m = levels[0]
    if taps is not None:
        taps[p + "merge"] = m
    e = glob_ln(m, sd[p + "final_norm.norm.gamma"], sd[p + "final_norm.norm.beta"])
    e = prelu1(e, sd[p + "final_norm.act.weight"])
    out = F.conv1d(e, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"]) + x
    if taps is not None:
        taps[p + "out"] = out
    return out


def tac(x4: Tensor, sd: Dict[str, Tensor], p: str,
        taps: Optional[dict] = None) -> Tensor:
    """Transform-average-concatenate, groupcomm_sudormrf_v2.py:356-384.

    x4: [B, G, n, L].  Per (b, t): h_g = PReLU(W1 x_g + b1); mean over groups;
    q = PReLU(W2 mean + b2); o_g = PReLU(W3 [h_g; q] + b3); then GlobLN over
    (n, L) per (b, g) and a residual add.
    """
    B, G, n, L = x4.shape
    xin = x4.permute(0, 3, 1, 2).reshape(B * L * G, n)
    h = prelu1(F.linear(xin, sd[p + "TAC_input.0.weight"], sd[p + "TAC_input.0.bias"]),
               sd[p + "TAC_input.1.weight"]).view(B, L, G, -1)
    mean = h.mean(2).reshape(B * L, -1)
    q = prelu1(F.linear(mean, sd[p + "TAC_mean.0.weight"], sd[p + "TAC_mean.0.bias"]),
               sd[p + "TAC_mean.1.weight"])
    q = q.unsqueeze(1).expand(B * L, G, q.shape[-1])
    cat = torch.cat([h.view(B * L, G, -1), q], dim=2).reshape(B * L * G, -1)
    o = prelu1(F.linear(cat, sd[p + "TAC_output.0.weight"], sd[p + "TAC_output.0.bias"]),
               sd[p + "TAC_output.1.weight"])
    o = o.view(B, L, G, n).permute(0, 2, 3, 1).contiguous()      # B, G, n, L
    if taps is not None:
        taps[p + "TAC_output"] = o
    o = glob_ln(o.view(B * G, n, L), sd[p + "TAC_norm.gamma"], sd[p + "TAC_norm.beta"])
    return x4 + o.view(B, G, n, L)


def gc_uconv_block(x: Tensor, sd: Dict[str, Tensor], p: str, depth: int, G: int,
                   taps: Optional[dict] = None) -> Tensor:
    """groupcomm_sudormrf_v2.py:405-418: TAC over the group view, then ONE small
    U-ConvBlock shared by all groups, groups folded into the batch."""
    B, C, L = x.shape
    t = tac(x.view(B, G, C // G, L), sd, p + "TAC.", taps)
    if taps is not None:
        taps[p + "TAC"] = t.reshape(B, C, L)
    y = uconv_block(t.reshape(B * G, C // G, L), sd, p + "UBlock.", depth, taps)
    return y.view(B, C, L)


# --------------------------------------------------------------------------
# whole-model forwards
# --------------------------------------------------------------------------
def forward(cfg: Config, sd: Dict[str, Tensor], wav: Tensor,
            taps: Optional[dict] = None, dtype=torch.float32) -> Tensor:
    """SuDORMRF.forward (improved_sudormrf.py:283-301) /
    GroupCommSudoRmRf.forward (groupcomm_sudormrf_v2.py:302-322).

    wav: [B, A, T] (A = 1, or in_audio_channels for groupcomm) -> [B, S*A, T].
    """
    if cfg.variant == "causal":
        return causal_forward(cfg, sd, wav, taps=taps, dtype=dtype)
    if cfg.variant == "original":
        return original_forward(cfg, sd, wav, taps=taps, dtype=dtype)
    if wav.dim() != 3:
        raise RuntimeError("expected a 3-D input [batch, audio_channels, time]")
    sd = {k: v.to(device=wav.device, dtype=dtype) for k, v in sd.items()}   # (bench.py times this op sequence on cuda too)
    T = wav.shape[-1]
    k, hop = cfg.enc_kernel_size, cfg.hop
    S, N = cfg.num_sources, cfg.enc_num_basis
    A = cfg.in_audio_channels if cfg.variant == "groupcomm" else 1
    x = pad_wave(cfg, wav, dtype)
    x = F.conv1d(x, sd["encoder.weight"], None, stride=hop, padding=hop)   # :286
    if taps is not None:
        taps["encoder"] = x
    s = x                                                                   # :289
    x = glob_ln(x, sd["ln.gamma"], sd["ln.beta"])                           # :291
    x = F.conv1d(x, sd["bottleneck.weight"], sd["bottleneck.bias"])         # :292
    if taps is not None:
        taps["bottleneck"] = x
    for i in range(cfg.num_blocks):                                         # :293
        if cfg.variant == "improved":
            x = uconv_block(x, sd, f"sm.{i}.", cfg.upsampling_depth, taps)
        else:
            x = gc_uconv_block(x, sd, f"sm.{i}.", cfg.upsampling_depth,
                               cfg.group_size, taps)
    x = prelu1(x, sd["mask_net.0.weight"])                                  # :295
    x = F.conv1d(x, sd["mask_net.1.weight"], sd["mask_net.1.bias"])
    x = x.view(x.shape[0], S * A, N, -1)                                    # :296
    x = torch.relu(x) * s.unsqueeze(1)                                      # :297-298
    if taps is not None:
        taps["masked"] = x.reshape(x.shape[0], -1, x.shape[-1])
    y = F.conv_transpose1d(x.reshape(x.shape[0], -1, x.shape[-1]),
                           sd["decoder.weight"], None, stride=hop, padding=hop,
                           output_padding=hop - 1)                          # :300
    return y[..., :T]                                                       # :301


def causal_weight(w: Tensor) -> Tensor:
    """ScaledWSConv1d.get_weight, causal_improved_sudormrf_v3.py:21-27: the last ``kernel_size // 2`` taps (the
    future samples under the symmetric padding) are zeroed; kernels shorter than 3 are left alone."""
    ks = w.shape[-1]
    if ks < 3:
        return w
    m = torch.ones_like(w)
    m[..., -(ks // 2):] = 0.
    return w * m


def causal_uconv_block(x: Tensor, sd: Dict[str, Tensor], p: str, depth: int, alpha: float = 1.0, beta: float = 1.0,
                       taps: Optional[dict] = None) -> Tensor:
    """causal_improved_sudormrf_v3.py:98-118."""
    Ci = sd[p + "proj_1x1.conv.weight"].shape[0]
    residual = x
    o = F.conv1d(x / beta, sd[p + "proj_1x1.conv.weight"], sd[p + "proj_1x1.conv.bias"])       # :105
    if taps is not None:
        taps[p + "proj_1x1.conv"] = o
    o = prelu1(o, sd[p + "proj_1x1.act.weight"])
    outs = []
    for d in range(depth):                                                                      # :106-111
        o = F.conv1d(o, causal_weight(sd[p + f"spp_dw.{d}.conv.weight"]), sd[p + f"spp_dw.{d}.conv.bias"],
                     stride=1 if d == 0 else 2, padding=10, groups=Ci)
        if taps is not None:
            taps[p + f"spp_dw.{d}.conv"] = o
        o = prelu1(o, sd[p + f"spp_dw.{d}.act.weight"])
        outs.append(o)
    for _ in range(depth - 1):                                                                  # :114-116
        up = F.interpolate(outs.pop(-1), scale_factor=2, mode="nearest")
        outs[-1] = outs[-1] + up
    if taps is not None:
        taps[p + "merged"] = outs[-1]
    y = F.conv1d(outs[-1], sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return y * sd[p + "skipinit_gain"] * alpha + residual                                       # :118


def causal_forward(cfg: Config, sd: Dict[str, Tensor], wav: Tensor, taps: Optional[dict] = None,
                   dtype=torch.float32) -> Tensor:
    """CausalSuDORMRF.forward, causal_improved_sudormrf_v3.py:191-211.  wav [B, A, T] -> [B, S*A, T]."""
    if wav.dim() != 3:
        raise RuntimeError("expected a 3-D input [batch, audio_channels, time]")
    sd = {k: v.to(device=wav.device, dtype=dtype) for k, v in sd.items()}
    T = wav.shape[-1]
    k, hop = cfg.enc_kernel_size, cfg.hop
    x = pad_wave(cfg, wav, dtype)                                                               # :193
    x = F.conv1d(x, causal_weight(sd["encoder.weight"]), None, stride=hop, padding=(2 * k - 1) // 2)   # :194
    if taps is not None:
        taps["encoder"] = x
    x = F.conv1d(x, sd["bottleneck.weight"], sd["bottleneck.bias"])                             # :199
    if taps is not None:
        taps["bottleneck"] = x
    for i in range(cfg.num_blocks):                                                             # :200
        x = causal_uconv_block(x, sd, f"sm.{i}.", cfg.upsampling_depth, taps=taps)
        if taps is not None:
            taps[f"sm.{i}.out"] = x
    x = prelu1(x, sd["mask_net.0.weight"])                                                      # :202
    x = F.conv1d(x, sd["mask_net.1.weight"], sd["mask_net.1.bias"])
    if taps is not None:
        taps["mask_net.1"] = x
    x = prelu1(x, sd["mask_nl_class.weight"])                                                   # :206 (no product with the encoder output, :207)
    y = F.conv_transpose1d(x, sd["decoder.weight"], None, stride=hop, padding=hop, output_padding=hop - 1)   # :209
    return y[..., :T]                                                                           # :210


def prelu_c(x: Tensor, slope: Tensor) -> Tensor:
    """nn.PReLU(C): one slope per channel (sudormrf.py:33,71)."""
    return torch.where(x >= 0, x, x * slope.view(1, -1, 1))


def original_ublock(x: Tensor, sd: Dict[str, Tensor], p: str, depth: int, taps: Optional[dict] = None) -> Tensor:
    """UBlock.forward, sudormrf.py:164-185.  GroupNorm(1, C, eps=1e-8) (:32,56,70,117) normalises over (C, L) with
    the biased variance and the eps inside the square root, i.e. the arithmetic of ``glob_ln``.

    y   = PReLU_c(GN(W1 x + b1))                               (:171)
    z_0 = GN(dw5_s1(y)), z_d = GN(dw5_s2(z_{d-1}))             (:172-177)
    m   = z_0 + up2(z_1 + up2(z_2 + ...))  (nearest)           (:180-182)
    e   = GN(Wexp PReLU_c(GN(m)) + bexp)                       (:184)
    out = PReLU_c(GN(e + x))                                   (:186)
    """
    ci = sd[p + "proj_1x1.conv.weight"].shape[0]
    y = F.conv1d(x, sd[p + "proj_1x1.conv.weight"], sd[p + "proj_1x1.conv.bias"])
    if taps is not None:
        taps[p + "proj_1x1.conv"] = y
    y = prelu_c(glob_ln(y, sd[p + "proj_1x1.norm.weight"], sd[p + "proj_1x1.norm.bias"]), sd[p + "proj_1x1.act.weight"])
    levels = []
    cur = y
    for d in range(depth):
        z = F.conv1d(cur, sd[p + f"spp_dw.{d}.conv.weight"], sd[p + f"spp_dw.{d}.conv.bias"],
                     stride=1 if d == 0 else 2, padding=2, groups=ci)       # :141-154: 5 taps at every level
        if taps is not None:
            taps[p + f"spp_dw.{d}.conv"] = z
        cur = glob_ln(z, sd[p + f"spp_dw.{d}.norm.weight"], sd[p + f"spp_dw.{d}.norm.bias"])
        levels.append(cur)
    for _ in range(depth - 1):
        top = levels.pop()
        levels[-1] = levels[-1] + F.interpolate(top, scale_factor=2, mode="nearest")
    m = levels[0]
    if taps is not None:
        taps[p + "merge"] = m
    e = prelu_c(glob_ln(m, sd[p + "final_norm.norm.weight"], sd[p + "final_norm.norm.bias"]),
                sd[p + "final_norm.act.weight"])
    e = F.conv1d(e, sd[p + "conv_1x1_exp.conv.weight"], sd[p + "conv_1x1_exp.conv.bias"])
    if taps is not None:
        taps[p + "conv_1x1_exp.conv"] = e
    u = glob_ln(e, sd[p + "conv_1x1_exp.norm.weight"], sd[p + "conv_1x1_exp.norm.bias"]) + x
    if taps is not None:
        taps[p + "sum"] = u
    out = prelu_c(glob_ln(u, sd[p + "module_act.norm.weight"], sd[p + "module_act.norm.bias"]),
                  sd[p + "module_act.act.weight"])
    if taps is not None:
        taps[p + "out"] = out
    return out


def original_forward(cfg: Config, sd: Dict[str, Tensor], wav: Tensor, taps: Optional[dict] = None,
                     dtype=torch.float32) -> Tensor:
    """SuDORMRF.forward of the ORIGINAL model, sudormrf.py:266-291.  wav [B, 1, T] -> [B, S, T]."""
    if wav.dim() != 3:
        raise RuntimeError("expected a 3-D input [batch, 1, time]")
    sd = {k: v.to(device=wav.device, dtype=dtype) for k, v in sd.items()}
    T = wav.shape[-1]
    hop, S, N = cfg.hop, cfg.num_sources, cfg.enc_num_basis
    x = pad_wave(cfg, wav, dtype)                                                               # :268, :283-293
    x = torch.relu(F.conv1d(x, sd["encoder.0.weight"], sd["encoder.0.bias"], stride=hop, padding=hop))   # :269
    if taps is not None:
        taps["encoder"] = x
    s = x                                                                                       # :272
    x = glob_ln(x, sd["ln.weight"], sd["ln.bias"])                                              # :275
    x = F.conv1d(x, sd["l1.weight"], sd["l1.bias"])                                             # :276
    if taps is not None:
        taps["l1"] = x
    for i in range(cfg.num_blocks):                                                             # :277
        x = original_ublock(x, sd, f"sm.{i}.", cfg.upsampling_depth, taps)
    if cfg.out_channels != N:                                                                   # :279-281
        x = F.conv1d(x, sd["reshape_before_masks.weight"], sd["reshape_before_masks.bias"])
        if taps is not None:
            taps["reshape_before_masks"] = x
    x = F.conv2d(x.unsqueeze(1), sd["m.weight"], sd["m.bias"], padding=(N - N // 2, 0))         # :284, :239-242
    if taps is not None:
        taps["m"] = x
    x = torch.sigmoid(x) if S == 1 else torch.softmax(x, dim=1)                                 # :285-288
    x = x * s.unsqueeze(1)                                                                      # :289
    if taps is not None:
        taps["masked"] = x.reshape(x.shape[0], -1, x.shape[-1])
    y = F.conv_transpose1d(x.reshape(x.shape[0], -1, x.shape[-1]), sd["decoder.weight"], sd["decoder.bias"],
                           stride=hop, padding=hop, output_padding=hop - 1, groups=S)           # :291, :245-252
    return y[..., :T]                                                                           # :292


def mixture_consistency(est: Tensor, mix: Tensor,
                        mix_weights_type: str = "uniform") -> Tensor:
    """mixture_consistency.py:14-36."""
    S = est.shape[1]
    resid = mix - est.sum(1, keepdim=True)
    if mix_weights_type == "uniform":
        w = 1.0 / S
    elif mix_weights_type == "magsq":
        w = (est ** 2).mean(-1, keepdim=True)
        w = w / (w.sum(1, keepdim=True) + 1e-9)
    else:
        raise ValueError(
            "Invalid mixture consistency weight type: {}".format(mix_weights_type))
    return est + w * resid


# --------------------------------------------------------------------------
# the steps either side of the forward
# --------------------------------------------------------------------------
def separate(cfg: Config, sd: Dict[str, Tensor], wav: Tensor,
             apply_mixture_consistency: bool = False, dtype=torch.float32) -> Tensor:
    """The README inference recipe, README.md:100-114.  wav [B, T] -> [B, S, T]."""
    wav = wav.to(dtype)
    std = wav.std(-1, keepdim=True)                      # README.md:101 (unbiased)
    mean = wav.mean(-1, keepdim=True)                    # README.md:102
    x = (wav - mean) / (std + 1e-9)                      # README.md:103
    rec = forward(cfg, sd, x.unsqueeze(1), dtype=dtype)  # README.md:106
    rec = (rec * std.unsqueeze(1)) + mean.unsqueeze(1)   # README.md:109
    if apply_mixture_consistency:                        # README.md:113-114
        rec = mixture_consistency(rec, x.unsqueeze(1))
    return rec


def pit_sisdr(pr: Tensor, tgt: Tensor, mix: Optional[Tensor] = None, zero_mean: bool = False,
              improvement: bool = False, eps: float = 1e-9):
    """PermInvariantSISDR.forward with backward_loss=False, return_individual_results=True
    (sisdr.py:95-194).  Returns (best [B], index of the best permutation [B]) with the
    permutations in itertools.permutations(range(S)) order (sisdr.py:87-89)."""
    import itertools
    n = min(pr.shape[-1], tgt.shape[-1])                                  # sisdr.py:96-102
    if mix is not None:
        n = min(n, mix.shape[-1])
        mix = mix[:, :, :n]
    pr, tgt = pr[:, :, :n], tgt[:, :, :n]
    if zero_mean:                                                         # sisdr.py:104-111
        pr = pr - pr.mean(-1, keepdim=True)
        tgt = tgt - tgt.mean(-1, keepdim=True)
        if mix is not None:
            mix = mix - mix.mean(-1, keepdim=True)

    def dot(a, b):                                                        # sisdr.py:114-116
        return torch.sum(a * b, dim=-1, keepdim=True)

    def permuted(p, t, tt):                                               # sisdr.py:118-126
        s_t = dot(p, t) / (tt + eps) * t
        e_t = p - s_t
        return 10 * torch.log10(dot(s_t, s_t) / (dot(e_t, e_t) + eps))

    S = pr.shape[1]
    tt = dot(tgt, tgt)                                                    # sisdr.py:134
    cols = [permuted(pr[:, list(perm), :], tgt, tt)                       # sisdr.py:136-142
            for perm in itertools.permutations(range(S))]
    allp = torch.cat(cols, -1)
    best, idx = torch.max(allp.mean(-2), -1)                              # sisdr.py:143
    if improvement:                                                       # sisdr.py:145-150
        base = permuted(mix.repeat(1, S, 1), tgt, tt)
        best = best - base.mean()
    return best, idx


def stabilized_pit_sisdr(pr: Tensor, tgt: Tensor, zero_mean: bool = False, single_source: bool = False,
                         improvement: bool = False, eps: float = 1e-9):
    """StabilizedPermInvSISDRMetric.forward with backward_loss=False, return_individual_results=True
    (sisdr.py:460-591): n_estimated = pr.shape[1] >= n_actual = tgt.shape[1]; the metric of every validation set of
    run_fuss_separation.py:111-131.  Returns (best [B], index of the best assignment [B]) with the assignments in
    itertools.permutations(range(n_estimated), r=n_actual) order (sisdr.py:490-492)."""
    import itertools
    if single_source:                                                     # sisdr.py:576-577
        pr = torch.sum(pr, -2, keepdim=True)
    if zero_mean:                                                         # sisdr.py:498-502, 579-580
        pr = pr - pr.mean(-1, keepdim=True)
        tgt = tgt - tgt.mean(-1, keepdim=True)

    def dot(a, b):                                                        # sisdr.py:504-506
        return torch.sum(a * b, dim=-1, keepdim=True)

    def stabilized(p, t, tt):                                             # sisdr.py:508-515
        pp = dot(p, p)
        rho_sq = dot(p, t) ** 2 / (pp * tt + eps)
        return 10 * torch.log10((rho_sq + eps) / (1. - rho_sq + eps))

    n_est, n_act = pr.shape[1], tgt.shape[1]
    assert n_est >= n_act
    tt = dot(tgt, tgt)                                                    # sisdr.py:524
    cols = [stabilized(pr[:, list(perm), :], tgt, tt)                     # sisdr.py:526-532
            for perm in itertools.permutations(range(n_est), r=n_act)]
    best, idx = torch.max(torch.cat(cols, -1).mean(-2), -1)               # sisdr.py:533
    if improvement:                                                       # sisdr.py:535-541: the mixture is the sum of the targets
        mix = torch.sum(tgt, -2, keepdim=True)
        if zero_mean:
            mix = mix - mix.mean(-1, keepdim=True)
        best = best - stabilized(mix.repeat(1, n_act, 1), tgt, tt).mean()
    return best, idx


def pairwise_neg_sdr(est: Tensor, tgt: Tensor, sdr_type: str = "sisdr", zero_mean: bool = True,
                     take_log: bool = True) -> Tensor:
    """PairwiseNegSDR.forward (sisdr.py:416-457): [B, S, S], entry [b, i, j] = -sdr(estimate i, target j)."""
    assert sdr_type in ("snr", "sisdr", "sdsdr")
    if zero_mean:                                                         # sisdr.py:419-423
        tgt = tgt - tgt.mean(dim=2, keepdim=True)
        est = est - est.mean(dim=2, keepdim=True)
    s_t = tgt.unsqueeze(1)                                                # sisdr.py:425-426
    s_e = est.unsqueeze(2)
    if sdr_type in ("sisdr", "sdsdr"):                                    # sisdr.py:428-435
        dot = torch.sum(s_e * s_t, dim=3, keepdim=True)
        energy = torch.sum(s_t ** 2, dim=3, keepdim=True) + 1e-8
        proj = dot * s_t / energy
    else:
        proj = s_t.repeat(1, s_t.shape[2], 1, 1)                          # sisdr.py:438
    noise = s_e - s_t if sdr_type in ("sdsdr", "snr") else s_e - proj     # sisdr.py:439-442
    sdr = torch.sum(proj ** 2, dim=3) / (torch.sum(noise ** 2, dim=3) + 1e-8)   # sisdr.py:444-445
    if take_log:
        sdr = 10 * torch.log10(sdr + 1e-8)                                # sisdr.py:446-447
    return -sdr


def pit_from_pairwise(pw: Tensor):
    """PITLossWrapper.find_best_perm with perm_reduce=None (sisdr.py:326-369): (min loss [B, 1], permutation index [B])."""
    import itertools
    n_src = pw.shape[1]
    pwl = pw.transpose(-1, -2)
    perms = list(itertools.permutations(range(n_src)))
    loss_set = torch.stack([sum(pwl[:, j, p[j]] for j in range(n_src)) / n_src for p in perms], dim=1)
    idx = torch.argmin(loss_set, dim=1)
    return loss_set.min(dim=1, keepdim=True)[0], idx


# --------------------------------------------------------------------------
# tolerance used by every parity test
# --------------------------------------------------------------------------
def parity_errors(y: Tensor, ref: Tensor):
    """Returns (max over samples of max|y-ref|/max|ref|, rel-L2)."""
    y = y.double().cpu()
    ref = ref.double().cpu()
    B = ref.shape[0]
    d = (y - ref).reshape(B, -1).abs().amax(1)
    m = ref.reshape(B, -1).abs().amax(1).clamp_min(1e-30)
    rel_max = float((d / m).max())
    rel_l2 = float((y - ref).norm() / ref.norm().clamp_min(1e-30))
    return rel_max, rel_l2

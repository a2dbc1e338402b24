"""CPU oracle for the NVF per-leaf-block decoder hot path.

TEST INFRASTRUCTURE ONLY.  This module is a restatement, in plain PyTorch fp32
CPU ops, of the reference's algorithm for the hot path.  It exists so that the
CUDA path can be checked against it.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it; the
product package `nvfpcc_b200` never does.

Parity status: PINNED.  `tests/golden/*.npz` were produced by importing the
unmodified reference modules from /root/reference (oracle/gen_golden.py, which
uses oracle/ref_import.py) and `tests/test_oracle_golden.py` checks every
function below against them.

Every function cites the reference lines it follows (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# straight-through / bounded helpers
# --------------------------------------------------------------------------


class _RoundSTE(torch.autograd.Function):
    """utils/network.py:25-32 BypassRound: round forward, identity backward."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _Round16STE(torch.autograd.Function):
    """utils/network.py:43-50 BypassRound16: round(x*16)/16, identity backward."""

    @staticmethod
    def forward(ctx, x):
        return torch.round(x * 16) / 16

    @staticmethod
    def backward(ctx, g):
        return g


class _LowerBound(torch.autograd.Function):
    """utils/network.py:56-72 and gdn_3d.py:13-29: clamp(min) whose gradient
    passes where x >= bound or the gradient pushes x upwards (g < 0)."""

    @staticmethod
    def forward(ctx, x, bound: float):
        ctx.save_for_backward(x)
        ctx.bound = bound
        return torch.clamp(x, min=bound)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        ok = torch.logical_or(x >= ctx.bound, g < 0)
        return g * ok, None


round_ste = _RoundSTE.apply
round16_ste = _Round16STE.apply
lower_bound = _LowerBound.apply

# --------------------------------------------------------------------------
# parameter construction (state_dict-compatible plain dicts of tensors)
# --------------------------------------------------------------------------

REPARAM_OFFSET = 2.0 ** -18
PEDESTAL = REPARAM_OFFSET ** 2          # gdn_3d.py:51
BETA_BOUND = (1e-6 + PEDESTAL) ** 0.5   # gdn_3d.py:52
GAMMA_BOUND = REPARAM_OFFSET            # gdn_3d.py:53


def kaiming_from_seed(shape: Sequence[int], seed: np.ndarray) -> torch.Tensor:
    """utils/network.py:377-400.  fan_in = size(1) * receptive field, for BOTH
    conv (Cout,Cin,k,k,k) and convT (Cin,Cout,k,k,k) kernels (the reference uses
    dim 1 regardless of layout); uniform(-bound, bound) driven by the seed."""
    shape = tuple(int(s) for s in shape)
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    fan_in = shape[1] * rf
    std = np.sqrt(2.0) / np.sqrt(fan_in)
    bound = np.sqrt(3.0) * std
    n = int(np.prod(shape))
    s = torch.from_numpy(np.asarray(seed[:n]).reshape(shape)).float()
    return (s - 0.5) * 2 * bound


def _conv_params(prefix: str, kshape, n_bias: int, bias_fan: int, seed: np.ndarray, ptr: int,
                 out: Dict[str, torch.Tensor]) -> int:
    """Shared body of QConvTranspose3d/QConv3d/IConv3d.__init__
    (utils/network.py:565-604, 625-667, 691-733): zero `kernel`/`b` parameters,
    `kernel_init`/`b_init` buffers cut from the seed stream; returns new ptr."""
    n = int(np.prod(kshape))
    out[prefix + ".kernel"] = torch.zeros(kshape, dtype=torch.float32)
    out[prefix + ".b"] = torch.zeros(n_bias, dtype=torch.float32)
    out[prefix + ".kernel_init"] = kaiming_from_seed(kshape, seed[ptr:ptr + n])
    b_seed = torch.from_numpy(np.asarray(seed[ptr + n:ptr + n + n_bias])).float()
    out[prefix + ".b_init"] = (b_seed - 0.5) * 2 * (1 / np.sqrt(bias_fan))
    return ptr + n + n_bias


def _gdn_params(prefix: str, ch: int, out: Dict[str, torch.Tensor]) -> None:
    """gdn_3d.py:50-69 (identical for GDN3d and IGDN3d)."""
    out[prefix + ".beta"] = torch.sqrt(torch.ones(ch) + PEDESTAL)
    out[prefix + ".gamma"] = torch.sqrt(0.1 * torch.eye(ch) + PEDESTAL)
    out[prefix + ".pedestal"] = torch.tensor([PEDESTAL], dtype=torch.float32)


def make_state(ch: int, channels: Sequence[int], seed: np.ndarray) -> Dict[str, torch.Tensor]:
    """Builds the full `Net` state_dict (NVFPCC.py:32-39) in construction
    order latent_gen -> entropy_coder -> reconstructor, consuming the SEED3
    stream exactly as utils/network.py:4605-4607 and 4671-4751 do."""
    c0, c1, c2, c3 = (int(c) for c in channels)
    sd: Dict[str, torch.Tensor] = {}
    p = 0
    # SingleLayerLatentGen: IConv3d(ch->ch, k=1)  (utils/network.py:4597-4608)
    p = _conv_params("latent_gen.h_analysis_2", (ch, ch, 1, 1, 1), ch, ch, seed, p, sd)
    _gdn_params("latent_gen.gdn_2", ch, sd)
    # QuantGaussianLikelihood (utils/network.py:4504-4512)
    sd["entropy_coder.sigma"] = torch.ones(1, ch, 1, 1, 1)
    sd["entropy_coder.mu"] = torch.zeros(1, ch, 1, 1, 1)
    # CompDecoder (utils/network.py:4658-4756)
    _gdn_params("reconstructor.activation", c0, sd)
    R = "reconstructor."
    p = _conv_params(R + "up0", (ch, c0, 5, 5, 5), c0, ch, seed, p, sd)       # convT: bias bound 1/sqrt(in)
    p = _conv_params(R + "conv0", (c0, c1, 5, 5, 5), c1, c0, seed, p, sd)
    p = _conv_params(R + "up1", (c1, c2, 5, 5, 5), c2, c1, seed, p, sd)
    p = _conv_params(R + "conv1", (c2, c2, 4, 4, 4), c2, c2, seed, p, sd)
    p = _conv_params(R + "up2", (c2, c3, 5, 5, 5), c3, c2, seed, p, sd)
    p = _conv_params(R + "conv2", (c3, c3, 4, 4, 4), c3, c3, seed, p, sd)
    p = _conv_params(R + "conv2_cls", (1, c3, 3, 3, 3), 1, c3, seed, p, sd)
    p = _conv_params(R + "conv1_cls", (1, c2, 3, 3, 3), 1, c2, seed, p, sd)
    p = _conv_params(R + "conv0_cls", (1, c1, 3, 3, 3), 1, c1, seed, p, sd)
    sd[R + "likelihood_model.sigma"] = torch.ones(1)
    sd[R + "likelihood_model.mu"] = torch.zeros(1)
    sd["_seed_consumed"] = torch.tensor(p)
    return sd


Q_KERNEL_LAYERS = ("up0", "conv0", "up1", "conv1", "up2", "conv2", "conv2_cls")  # get_q_params, network.py:4781-4792
I_KERNEL_LAYERS = ("conv1_cls", "conv0_cls")

# --------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------


def q_kernel(kernel: torch.Tensor, kernel_init: torch.Tensor, q: int, noise: torch.Tensor | None = None):
    """utils/network.py:606-620 / 669-686.  q==1: additive U(-1/32,1/32) noise
    ((rand-0.5)/16); q==2: straight-through round16; else identity; then
    + kernel_init."""
    if q == 1:
        if noise is None:
            noise = torch.rand_like(kernel)
        k = kernel + (noise - 0.5) * (1 / 16)
    elif q == 2:
        k = round16_ste(kernel)
    else:
        k = kernel
    return k + kernel_init


def gdn_effective(beta: torch.Tensor, gamma: torch.Tensor, pedestal: torch.Tensor):
    """gdn_3d.py:78-85 / 143-150 reparametrisation."""
    b = lower_bound(beta, BETA_BOUND) ** 2 - pedestal
    g = lower_bound(gamma, GAMMA_BOUND) ** 2 - pedestal
    return b, g


def gdn_apply(x: torch.Tensor, beta_eff: torch.Tensor, gamma_eff: torch.Tensor, inverse: bool):
    """gdn_3d.py:88-92 (divide) / 153-157 (multiply); gamma indexed [out,in]."""
    ch = x.shape[1]
    norm = torch.sqrt(F.conv3d(x * x, gamma_eff.view(ch, ch, 1, 1, 1), beta_eff))
    return x * norm if inverse else x / norm


def gaussian_bits(x: torch.Tensor, sigma: torch.Tensor, mu: torch.Tensor, step: float):
    """utils/network.py:139-161 GaussianModel.forward: bits of x under
    N(mu,sigma) with bin `step`, likelihood floored at 1e-8 by LowerBound,
    summed."""
    nd = torch.distributions.normal.Normal(0.0, 1.0)
    half = 0.5 * step
    up = nd.cdf((x - mu + half) / sigma)
    lo = nd.cdf((x - mu - half) / sigma)
    lik = lower_bound(up - lo, 1e-8)
    return (-1 * torch.log(lik) / np.log(2)).sum()


def effective_decoder_weights(sd: Dict[str, torch.Tensor], q: int, noises: Dict[str, torch.Tensor] | None = None):
    """Effective tensors CompDecoder.forward convolves with
    (utils/network.py:4758-4768 via the wrappers at :606-621, :669-687, :735-741)."""
    R = "reconstructor."
    w = {}
    for name in Q_KERNEL_LAYERS:
        nz = None if noises is None else noises.get(name)
        w[name + "_w"] = q_kernel(sd[R + name + ".kernel"], sd[R + name + ".kernel_init"], q, nz)
        w[name + "_b"] = sd[R + name + ".b"] + sd[R + name + ".b_init"]
    for name in I_KERNEL_LAYERS:
        if R + name + ".kernel" in sd:
            w[name + "_w"] = sd[R + name + ".kernel"] + sd[R + name + ".kernel_init"]
            w[name + "_b"] = sd[R + name + ".b"] + sd[R + name + ".b_init"]
    w["igdn_beta"], w["igdn_gamma"] = gdn_effective(
        sd[R + "activation.beta"], sd[R + "activation.gamma"], sd[R + "activation.pedestal"])
    return w


def decoder_forward(latent: torch.Tensor, w: Dict[str, torch.Tensor], aux: bool = True, keep: bool = False):
    """utils/network.py:4758-4768 on effective tensors.  Returns
    (out, cls0, cls1, logits, intermediates)."""
    x0 = F.conv_transpose3d(latent, w["up0_w"], w["up0_b"], 2, 2, 1)
    a0 = gdn_apply(x0, w["igdn_beta"], w["igdn_gamma"], inverse=True)
    a1 = F.relu(F.conv_transpose3d(a0, w["conv0_w"], w["conv0_b"], 2, 2, 1))
    cls0 = torch.sigmoid(F.conv3d(a1, w["conv0_cls_w"], w["conv0_cls_b"], 1, 1)) if aux else None
    a2 = F.relu(F.conv_transpose3d(a1, w["up1_w"], w["up1_b"], 2))
    a3 = F.relu(F.conv3d(a2, w["conv1_w"], w["conv1_b"]))
    cls1 = torch.sigmoid(F.conv3d(a3, w["conv1_cls_w"], w["conv1_cls_b"], 1, 1)) if aux else None
    a4 = F.relu(F.conv_transpose3d(a3, w["up2_w"], w["up2_b"], 2))
    a5 = F.relu(F.conv3d(a4, w["conv2_w"], w["conv2_b"]))
    logit = F.conv3d(a5, w["conv2_cls_w"], w["conv2_cls_b"], 1, 1)
    out = torch.sigmoid(logit)
    inter = dict(x0=x0, a0=a0, a1=a1, a2=a2, a3=a3, a4=a4, a5=a5) if keep else None
    return out, cls0, cls1, logit, inter


def net_bits(sd: Dict[str, torch.Tensor]):
    """utils/network.py:4777-4778 + 301-305: bits of round16(kernel) (the
    DELTA, not kernel+init) under the scalar Gaussian, bin 1/16, per layer."""
    R = "reconstructor."
    sig = torch.abs(sd[R + "likelihood_model.sigma"])
    mu = sd[R + "likelihood_model.mu"]
    return torch.stack([
        gaussian_bits(round16_ste(sd[R + n + ".kernel"]).reshape(-1, 1), sig, mu, 1 / 16)
        for n in Q_KERNEL_LAYERS])


def latent_gen(emb: torch.Tensor, sd: Dict[str, torch.Tensor]):
    """utils/network.py:4610-4612: 1x1x1 IConv3d + GDN."""
    L = "latent_gen."
    k = sd[L + "h_analysis_2.kernel"] + sd[L + "h_analysis_2.kernel_init"]
    b = sd[L + "h_analysis_2.b"] + sd[L + "h_analysis_2.b_init"]
    h = F.conv3d(emb, k, b)
    be, ga = gdn_effective(sd[L + "gdn_2.beta"], sd[L + "gdn_2.gamma"], sd[L + "gdn_2.pedestal"])
    return gdn_apply(h, be, ga, inverse=False)


def entropy_coder(latent: torch.Tensor, sd: Dict[str, torch.Tensor], mode: str, noise: torch.Tensor | None = None):
    """utils/network.py:4514-4539: returns (round_ste(latent), bits) where the
    rate is evaluated on latent+U(-.5,.5) in 'train' and on the rounded latent
    in 'eval'."""
    if noise is None:
        noise = torch.rand_like(latent)
    x_noisy = latent + (noise - 0.5)
    x_rounded = round_ste(latent)
    x_form = x_noisy if mode == "train" else x_rounded
    bits = gaussian_bits(x_form, torch.abs(sd["entropy_coder.sigma"]), sd["entropy_coder.mu"], 1.0)
    return x_rounded, bits


def net_forward(emb, sd, mode: str, q: int, latent_noise=None, kernel_noises=None, keep=False):
    """NVFPCC.py:41-45 Net.forward."""
    latent = latent_gen(emb, sd)
    latent_rounded, latent_bits = entropy_coder(latent, sd, mode, latent_noise)
    w = effective_decoder_weights(sd, q, kernel_noises)
    out, cls0, cls1, logit, inter = decoder_forward(latent_rounded, w, aux=True, keep=keep)
    return dict(out=out, cls=[cls0, cls1, out], net_bits=net_bits(sd), latent_bits=latent_bits,
                latent_rounded=latent_rounded, logit=logit, inter=inter, w=w)


def reconstruct(latent, sd, q: int = 2):
    """NVFPCC.py:47-49 Net.reconstruct (decode path; aux heads/net_bits dropped)."""
    w = effective_decoder_weights(sd, q)
    return decoder_forward(latent, w, aux=False)[0]


# --------------------------------------------------------------------------
# losses / metrics
# --------------------------------------------------------------------------


def focal_dense(p, gt, alpha=0.97, gamma=2):
    """utils/loss.py:61-72 get_focal_dense (sum, natural log, clamp 1e-9)."""
    mask = gt.bool()
    imask = ~mask
    Fv = p * ((-1) * imask + 1 * mask) + 1 * imask
    alphas = alpha * ((-1) * imask + 1 * mask) + 1 * imask
    Fv = torch.clamp(Fv, min=1e-9)
    return ((-1) * alphas * ((1 - Fv) ** gamma) * torch.log(Fv)).sum()


def surf_focal_dense(p, gt, dist, beta=1, alpha=0.97, gamma=2):
    """utils/loss.py:94-111 get_surf_focal_dense."""
    mask = gt.bool()
    imask = ~mask
    dist_w = dist + mask * beta
    Fv = p * ((-1) * imask + 1 * mask) + 1 * imask
    alphas = alpha * ((-1) * imask + 1 * mask) + 1 * imask
    Fv = torch.clamp(Fv, min=1e-9)
    return ((-1) * alphas * ((1 - Fv) ** gamma) * dist_w * torch.log(Fv)).sum()


def sse1(p, dist, thh):
    """utils/loss.py:113-121 get_sse1 -> (sse, denom)."""
    pred = (p > thh).float()
    return torch.square(pred * dist).sum(), pred.sum()


def acc_dense(p, gt, thh=0.5):
    """utils/loss.py:74-84 get_acc_dense -> (tp/ap, tn/an); also returns raw counts."""
    mask = gt.bool()
    imask = ~mask
    tp = ((p > thh) * mask).sum()
    ap = mask.sum()
    tn = ((p <= thh) * imask).sum()
    an = imask.sum()
    return tp / ap, tn / an, (tp, ap, tn, an)


def gt_pyramid(gt):
    """NVFPCC.py:76-88 MultiscaleProcessor: [8^3, 16^3, 32^3] by two MaxPool3d(2,2)."""
    x1 = F.max_pool3d(gt, 2, 2)
    x2 = F.max_pool3d(x1, 2, 2)
    return [x2, x1, gt]


def train_loss(res, gt, dist, n_pts, n_total, lmbda, w1, w2, focal_alpha=0.9):
    """NVFPCC.py:154-196: total loss of one weight-loop step."""
    gts = gt_pyramid(gt)
    b_latent = res["latent_bits"].sum() / n_pts
    b_net = res["net_bits"].sum() / n_total
    bpp_loss = b_latent * w1 + b_net * w2
    ms0 = focal_dense(res["cls"][0], gts[0], alpha=0.85)
    ms1 = focal_dense(res["cls"][1], gts[1], alpha=0.85)
    bce = surf_focal_dense(res["out"], gt, dist, beta=1, alpha=focal_alpha)
    loss = bce + ms0 + ms1 + lmbda * bpp_loss
    return dict(loss=loss, bce=bce, ms0=ms0, ms1=ms1, b_latent=b_latent, b_net=b_net)


# --------------------------------------------------------------------------
# threshold -> points
# --------------------------------------------------------------------------


def threshold_points(p: torch.Tensor, origins: np.ndarray, thh: float) -> Tuple[np.ndarray, np.ndarray]:
    """NVFPCC.py:631-637: strict `p > thh`; point = (i,j,k) index of
    out_dense[b,0,i,j,k] + origin[b]; block order, then row-major in-block
    order (the dense->sparse order assumed in SURVEY.md section 8c(ii))."""
    pts: List[np.ndarray] = []
    counts = []
    for b in range(p.shape[0]):
        idx = torch.nonzero(p[b, 0] > thh).numpy().astype(np.int32)
        pts.append(idx + np.asarray(origins[b], dtype=np.int32)[None, :])
        counts.append(idx.shape[0])
    allp = np.concatenate(pts, 0) if pts else np.zeros((0, 3), np.int32)
    return allp, np.asarray(counts, dtype=np.int32)


def quantize_state_16(sd: Dict[str, torch.Tensor], iqp: int = 16) -> Dict[str, torch.Tensor]:
    """manipulate_weights.py:36-51: round(w*iqp)/iqp on the 7 reconstructor
    kernels; aux-head keys and reconstructor *_init buffers are dropped
    (rebuilt from SEED3 at load time)."""
    keep_prefix = ("latent_gen.", "entropy_coder.", "reconstructor.activation.", "reconstructor.likelihood_model.")
    nd = {}
    for k, v in sd.items():
        if k.startswith("_"):
            continue
        if k.startswith(keep_prefix):
            nd[k] = v.clone()
    for n in Q_KERNEL_LAYERS:
        nd[f"reconstructor.{n}.kernel"] = torch.round(sd[f"reconstructor.{n}.kernel"] * iqp) / iqp
        nd[f"reconstructor.{n}.b"] = sd[f"reconstructor.{n}.b"].clone()
    return nd

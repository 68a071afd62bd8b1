"""ORACLE — test infrastructure only. Never imported by the product path (image-diffusion_b200/), only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

A functional fp32 restatement, in plain torch ops driven by a state_dict, of the reference hot path
(jklimmek/image-diffusion): UNet forward, CFG DDPM sampling loop, KL/VQ VAE encode/decode, VQ codebook lookup and
the UNet training-step loss. Every function cites the reference file:line it follows (paths relative to the
reference tree). It runs on CPU or, as a same-device fp32 checker, on CUDA.

Parity pin: the reference ships no tests or golden vectors, so the pin is the reference ITSELF run in
the build container: tests/golden/make_golden.py imports /root/reference/modules, runs it on seeded weights and
inputs, and stores its outputs under tests/golden/*.pt; tests/test_oracle_golden.py checks this file against those
vectors (CPU, no GPU, no /root/reference needed).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by the golden generator, the tests and bench.py)
# ---------------------------------------------------------------------------------------------
def seeded_state_dict(shapes: dict, seed: int, dtype=torch.float32) -> dict:
    """Fills a {name: shape} spec with reproducible values, independent of module construction order.

    1-D "weight" tensors are treated as norm gains (1 + 0.1 n), biases as 0.1 n, matrices/filters as
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like torch's default init. Special buffers keep their defining formula.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("time_embedding.factor"):
            half = shape[0]
            sd[name] = 10000 ** (torch.arange(0, half, dtype=torch.float32) / half)  # components.py:432
        elif name.endswith("ema_cluster_size"):
            sd[name] = torch.zeros(shape)
        elif name.endswith("codebook.embeddings.weight") or name.endswith("codebook.ema_w"):
            size = shape[0]
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / size  # components.py:254,263
        elif name.endswith("class_embedding.weight"):
            sd[name] = torch.randn(shape, generator=g)
        elif len(shape) == 1 and name.endswith("weight"):
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = sd[name].to(dtype)
    return sd


# ---------------------------------------------------------------------------------------------
# Scheduler  (components.py:364-424)
# ---------------------------------------------------------------------------------------------
class SchedulerTables:
    """components.py:366-397 — note "linear" is linear in sqrt(beta) (scaled-linear)."""

    def __init__(self, num_steps, beta_start=1e-4, beta_end=0.02, type="linear", device="cpu"):
        self.num_steps = num_steps
        if type == "cosine":  # components.py:380-387
            offset = 8e-3
            ts = torch.arange(num_steps + 1, dtype=torch.float32) / num_steps
            f = torch.cos((ts + offset) / (1 + offset) * math.pi / 2).pow(2)
            ah = f / f[0]
            betas = torch.clip(1 - ah[1:] / ah[:-1], min=0, max=0.999)
        elif type == "linear":  # components.py:389-392
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_steps) ** 2
        else:
            raise ValueError(type)
        self.betas = betas.to(device)
        self.alphas = 1.0 - self.betas
        self.alpha_cum_prod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alpha_cum_prod = torch.sqrt(self.alpha_cum_prod)
        self.sqrt_one_minus_alpha_cum_prod = torch.sqrt(1 - self.alpha_cum_prod)


def _b4(v):
    return v.view(-1, 1, 1, 1)


def add_noise(s: SchedulerTables, x, noise, t):
    """components.py:399-403"""
    return _b4(s.sqrt_alpha_cum_prod[t]) * x + _b4(s.sqrt_one_minus_alpha_cum_prod[t]) * noise


def posterior_step(s: SchedulerTables, xt, eps, t, z):
    """components.py:405-424 with the noise `z` injected instead of drawn (z is ignored when t[0] == 0)."""
    so = _b4(s.sqrt_one_minus_alpha_cum_prod[t])
    x0 = torch.clamp((xt - so * eps) / _b4(s.sqrt_alpha_cum_prod[t]), -1.0, 1.0)
    beta = _b4(s.betas[t])
    mean = (xt - (beta * eps) / so) / torch.sqrt(_b4(s.alphas[t]))
    if t[0] == 0:
        return mean, x0
    var = (1 - _b4(s.alpha_cum_prod[t - 1])) / (1.0 - _b4(s.alpha_cum_prod[t])) * beta
    return mean + var ** 0.5 * z, x0


def ddim_step(s: SchedulerTables, xt, eps, t: int, t_prev: int, eta: float, z=None, clamp_x0: bool = False):
    """Strided step x_t -> x_{t_prev} (t_prev < 0: final step, abar_prev = 1). NOT in the reference (its only sampler
    is the 1-step ancestral update above, components.py:405-424): SURVEY.md §8 row f4. Restates the published
    algorithm, Song, Meng & Ermon, "Denoising Diffusion Implicit Models", ICLR 2021, eq. 12 / 16, on the reference's
    tables (components.py:366-397). Pin: with eta = 1 and t_prev = t - 1 it must reproduce posterior_step (which is
    pinned on the reference's golden vectors) up to fp32 round-off - tests/test_oracle_golden.py."""
    a_t = s.alpha_cum_prod[t]
    a_p = s.alpha_cum_prod[t_prev] if t_prev >= 0 else torch.ones_like(a_t)
    x0 = (xt - torch.sqrt(1 - a_t) * eps) / torch.sqrt(a_t)
    if clamp_x0:
        x0 = x0.clamp(-1.0, 1.0)
    sigma = eta * torch.sqrt((1 - a_p) / (1 - a_t)) * torch.sqrt(torch.clamp(1 - a_t / a_p, min=0))
    out = torch.sqrt(a_p) * x0 + torch.sqrt(torch.clamp(1 - a_p - sigma ** 2, min=0)) * eps
    if float(sigma) > 0:
        out = out + sigma * z
    return out, x0


def cfg_sample_strided(unet_sd, unet_arch, sched: SchedulerTables, x_T, labels, cfg_scales, steps, eta=0.0, noises=None):
    """diffusion.py:46-56 with the ancestral update replaced by ddim_step over a decreasing subset of timesteps."""
    xt = x_T
    N = xt.shape[0]
    s4 = cfg_scales.view(-1, 1, 1, 1)
    steps = list(steps)
    for k, i in enumerate(steps):
        t = torch.full((N,), i, dtype=torch.long, device=xt.device)
        ec = unet_forward(unet_sd, unet_arch, xt, t, labels)
        eu = unet_forward(unet_sd, unet_arch, xt, t)
        eps = eu + s4 * (ec - eu)
        xt, _ = ddim_step(sched, xt, eps, i, steps[k + 1] if k + 1 < len(steps) else -1, eta,
                          None if noises is None else noises[k])
    return xt


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def _gn(sd, p, x, groups, silu):
    y = F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], 1e-5)
    return F.silu(y) if silu else y


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def mha(sd, p, x, heads, groups):
    """components.py:64-103: GroupNorm -> tokens -> q,k,v Linear of the SAME normalised tokens -> head-major split
    -> softmax(QK^T / sqrt(hd)) V -> merge -> out_proj -> + input."""
    B, C, H, W = x.shape
    hd = C // heads
    tok = _gn(sd, p + ".groupnorm", x, groups, False).flatten(2).transpose(1, 2)  # b (h w) c
    q = _lin(sd, p + ".to_q", tok).view(B, H * W, heads, hd).transpose(1, 2)
    k = _lin(sd, p + ".to_k", tok).view(B, H * W, heads, hd).transpose(1, 2)
    v = _lin(sd, p + ".to_v", tok).view(B, H * W, heads, hd).transpose(1, 2)
    w = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (w @ v).transpose(1, 2).reshape(B, H * W, C)
    o = _lin(sd, p + ".out_proj", o).transpose(1, 2).reshape(B, C, H, W)
    return o + x


def downsample(sd, p, x):
    """components.py:110-117: stride-2 pad-0 conv, then zero-pad the OUTPUT on the right/bottom."""
    return F.pad(_conv(sd, p + ".down", x, stride=2, padding=0), (0, 1, 0, 1), value=0.0)


def upsample(sd, p, x):
    """components.py:124-130: nearest 2x then conv3x3."""
    return _conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


def time_embedding(sd, p, t):
    """components.py:441-445 (sin first, then cos)."""
    a = t[:, None] / sd[p + ".factor"]
    e = torch.cat([torch.sin(a), torch.cos(a)], dim=-1)
    return _lin(sd, p + ".embeddings.2", F.silu(_lin(sd, p + ".embeddings.0", e)))


def diffusion_block(sd, p, x, temb, layers, heads, groups, skip=None):
    """components.py:513-538. The skip projection is ALWAYS a 1x1 conv (components.py:499-504)."""
    if skip is not None:
        x = torch.cat((x, skip), dim=1)
    for l in range(layers):
        resid = x
        x = _conv(sd, f"{p}.first_halfs.{l}.layers.2", _gn(sd, f"{p}.first_halfs.{l}.layers.0", x, groups, True))
        x = x + _lin(sd, f"{p}.time_projs.{l}.1", F.silu(temb))[:, :, None, None]
        x = _conv(sd, f"{p}.second_halfs.{l}.layers.2", _gn(sd, f"{p}.second_halfs.{l}.layers.0", x, groups, True))
        x = x + _conv(sd, f"{p}.residuals.{l}", resid, padding=0)
        x = mha(sd, f"{p}.self_attns.{l}", x, heads, groups)
    return x


def unet_forward(sd, arch, x, t, context=None, context_mask=None, taps=None):
    """unet.py:103-136. `taps` (optional dict) collects named intermediate activations for layer-level checks."""
    ch, layers, heads, groups = arch["channels"], arch["num_res_layers"], arch["num_heads"], arch["num_groups"]
    temb = time_embedding(sd, "time_embedding", t)
    if context is not None:  # unet.py:109-114
        c = F.one_hot(context, arch["num_classes"]).float() @ sd["class_embedding.weight"]
        if context_mask is not None:
            c = c * context_mask
        temb = temb + c
    x = _conv(sd, "in_conv", x)
    if taps is not None:
        taps["temb"], taps["in_conv"] = temb, x
    skips = []
    for i in range(len(ch) - 1):
        x = diffusion_block(sd, f"down_blocks.{i}", x, temb, layers, heads, groups)
        skips.append(x)
        if taps is not None:
            taps[f"down_blocks.{i}"] = x
        x = downsample(sd, f"downsamples.{i}", x)
    for i in range(len(arch["mid_channels"]) - 1):
        x = diffusion_block(sd, f"mid_blocks.{i}", x, temb, layers, heads, groups)
    if taps is not None:
        taps["mid"] = x
    for i in range(len(ch) - 1):
        x = upsample(sd, f"upsamples.{i}", x)
        x = diffusion_block(sd, f"ups.{i}", x, temb, layers, heads, groups, skip=skips.pop())
        if taps is not None:
            taps[f"ups.{i}"] = x
    return _conv(sd, "out_conv.2", _gn(sd, "out_conv.0", x, groups, True))  # unet.py:97-101,135


# ---------------------------------------------------------------------------------------------
# VAE  (vae.py, components.py:26-49, 133-315)
# ---------------------------------------------------------------------------------------------
def residual(sd, p, x, groups):
    """components.py:26-49: identity skip when Cin == Cout, else 1x1 conv."""
    h = _conv(sd, p + ".branch.2", _gn(sd, p + ".branch.0", x, groups, True))
    h = _conv(sd, p + ".branch.5", _gn(sd, p + ".branch.3", h, groups, True))
    if (p + ".residual_wrapper.weight") in sd:
        x = _conv(sd, p + ".residual_wrapper", x, padding=0)
    return h + x


def decoder_program(arch):
    """Layer list of Decoder.up (components.py:205-242) as (kind, index, cin, cout); channels arrive reversed
    (vae.py:66) and curr_res starts at init_resolution // 2**len(channels) (vae.py:71)."""
    ch = list(arch["channels"])[::-1]
    nres, attn_res = arch["dec_num_res_blocks"], arch["attn_resolutions"]
    res = arch["init_resolution"] // 2 ** len(arch["channels"])
    prog, i = [("conv1x1", 0, arch["z_dim"], arch["z_dim"]), ("conv3x3", 1, arch["z_dim"], ch[0])], 2
    for _ in range(nres):
        prog.append(("res", i, ch[0], ch[0])); i += 1
    prog.append(("attn", i, ch[0], ch[0])); i += 1
    for _ in range(nres):
        prog.append(("res", i, ch[0], ch[0])); i += 1
    for s in range(len(ch) - 1):
        cin = ch[s]
        for _ in range(nres):
            prog.append(("res", i, cin, ch[s + 1])); i += 1
            cin = ch[s + 1]
        if res in attn_res:
            prog.append(("attn", i, ch[s + 1], ch[s + 1])); i += 1
        prog.append(("up", i, ch[s + 1], ch[s + 1])); i += 1
        res *= 2
    for _ in range(nres):
        prog.append(("res", i, ch[-1], ch[-1])); i += 1
    prog.append(("gn_silu", i, ch[-1], ch[-1])); i += 2  # GroupNorm at i, SiLU at i+1
    prog.append(("conv3x3", i, ch[-1], arch["in_channels"]))
    return prog


def encoder_program(arch):
    """Layer list of Encoder.down (components.py:148-181)."""
    ch = list(arch["channels"])
    nres, attn_res = arch["enc_num_res_blocks"], arch["attn_resolutions"]
    zc = arch["z_dim"] if arch["bottleneck"] == "vq" else 2 * arch["z_dim"]
    res = arch["init_resolution"]
    prog, i = [("conv3x3", 0, arch["in_channels"], ch[0])], 1
    for s in range(len(ch) - 1):
        cin = ch[s]
        for _ in range(nres):
            prog.append(("res", i, cin, ch[s + 1])); i += 1
            cin = ch[s + 1]
        if res in attn_res:
            prog.append(("attn", i, ch[s + 1], ch[s + 1])); i += 1
        prog.append(("down", i, ch[s + 1], ch[s + 1])); i += 1
        res /= 2
    for _ in range(nres):
        prog.append(("res", i, ch[-1], ch[-1])); i += 1
    prog.append(("attn", i, ch[-1], ch[-1])); i += 1
    for _ in range(nres):
        prog.append(("res", i, ch[-1], ch[-1])); i += 1
    prog.append(("gn_silu", i, ch[-1], ch[-1])); i += 2
    prog.append(("conv3x3", i, ch[-1], zc)); i += 1
    prog.append(("conv1x1", i, zc, zc))
    return prog


def _run_program(sd, prefix, prog, x, arch):
    heads, groups = arch["num_heads"], arch["num_groups"]
    for kind, i, _, _ in prog:
        p = f"{prefix}.{i}"
        if kind == "conv1x1":
            x = _conv(sd, p, x, padding=0)
        elif kind == "conv3x3":
            x = _conv(sd, p, x)
        elif kind == "res":
            x = residual(sd, p, x, groups)
        elif kind == "attn":
            x = mha(sd, p, x, heads, groups)
        elif kind == "up":
            x = upsample(sd, p, x)
        elif kind == "down":
            x = downsample(sd, p, x)
        elif kind == "gn_silu":
            x = _gn(sd, p, x, groups, True)
    return x


def codebook_forward(sd, p, z, beta):
    """components.py:265-315 in eval mode: returns (z_q, quant_loss, perplexity, indices)."""
    B, C, H, W = z.shape
    e = sd[p + ".embeddings.weight"]
    x = z.permute(0, 2, 3, 1).reshape(B, H * W, C)
    d = torch.cdist(x, e[None, :].repeat(B, 1, 1))  # components.py:272
    idx = d.argmin(dim=-1).view(-1)  # components.py:275 (first minimal index)
    q = e[idx]
    flat = x.reshape(B * H * W, C)
    loss = beta * F.mse_loss(q, flat)  # components.py:301-302
    zq = (flat + (q - flat)).view(B, H, W, C).permute(0, 3, 1, 2)  # components.py:305-308
    probs = F.one_hot(idx, e.shape[0]).float().mean(dim=0)
    perplexity = torch.exp(-torch.sum(probs * torch.log(probs + 1e-6)))  # components.py:311-313
    return zq, loss, perplexity, idx


def vae_decode(sd, arch, z, quantize=False):
    """vae.py:115-121"""
    if arch["bottleneck"] == "kl" and quantize:
        raise ValueError("Cannot quantize in the KL model!")
    if quantize:
        z = codebook_forward(sd, "codebook", z, arch["codebook_beta"])[0]
    return _run_program(sd, "decoder.up", decoder_program(arch), z, arch)


def vae_encode(sd, arch, x, noise=None):
    """vae.py:92-113. `noise` is the injected reparametrisation draw (sample=True) or None (sample=False)."""
    z = _run_program(sd, "encoder.down", encoder_program(arch), x, arch)
    if arch["bottleneck"] == "vq":
        zq, loss, perp, _ = codebook_forward(sd, "codebook", z, arch["codebook_beta"])
        return zq, loss, perp
    mean, log_var = torch.chunk(z, 2, dim=1)
    log_var = torch.clamp(log_var, -30.0, 20.0)
    kl = -0.5 * torch.sum(1 + log_var - mean.pow(2) - log_var.exp(), dim=[1, 2, 3])
    if noise is not None:
        z = mean + noise * torch.exp(0.5 * log_var)
    return z, kl.mean(), 0.0


# ---------------------------------------------------------------------------------------------
# sampling loop and training step
# ---------------------------------------------------------------------------------------------
def cfg_sample(unet_sd, unet_arch, sched: SchedulerTables, x_T, labels, cfg_scales, noises, steps=None, trace=None):
    """diffusion.py:46-56 re-driven with injected noise. `cfg_scales` is an int64 (N,) tensor as in diffusion.py:44;
    `steps` is the list of timesteps to run (default: num_steps-1 .. 0); noises[k] is the draw of the k-th step."""
    xt = x_T
    N = xt.shape[0]
    s4 = cfg_scales.view(-1, 1, 1, 1)
    steps = list(reversed(range(sched.num_steps))) if steps is None else list(steps)
    for k, i in enumerate(steps):
        t = torch.full((N,), i, dtype=torch.long, device=xt.device)
        ec = unet_forward(unet_sd, unet_arch, xt, t, labels)
        eu = unet_forward(unet_sd, unet_arch, xt, t)
        eps = eu + s4 * (ec - eu)
        xt, _ = posterior_step(sched, xt, eps, t, noises[k] if i > 0 else None)
        if trace is not None:
            trace.append(xt)
    return xt


def train_step_loss(unet_sd, unet_arch, sched: SchedulerTables, latents, labels, noise, t, context_mask,
                    reparam_noise=None):
    """trainers/diffusion_trainer.py:141-170 with every random draw injected: KL reparametrisation of the stored
    (mean || logvar) latents, add_noise, masked-class forward, MSE(mean) against the noise."""
    x = latents.float()
    if reparam_noise is not None:
        mean, log_var = torch.chunk(x, 2, dim=1)
        x = mean + reparam_noise * torch.exp(0.5 * torch.clamp(log_var, -30.0, 20.0))
    x_noise = add_noise(sched, x, noise, t)
    pred = unet_forward(unet_sd, unet_arch, x_noise, t, labels, context_mask)
    return F.mse_loss(pred, noise)


# ---------------------------------------------------------------------------------------------
# parameter shape specs (what the reference classes register; verified by tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------
def unet_param_shapes(arch) -> dict:
    ch, mid, D = arch["channels"], arch["mid_channels"], arch["time_dim"]
    L = arch["num_res_layers"]
    s = {"class_embedding.weight": (arch["num_classes"], D), "time_embedding.factor": (D // 2,),
         "time_embedding.embeddings.0.weight": (4 * D, D), "time_embedding.embeddings.0.bias": (4 * D,),
         "time_embedding.embeddings.2.weight": (D, 4 * D), "time_embedding.embeddings.2.bias": (D,),
         "in_conv.weight": (ch[0], arch["z_dim"], 3, 3), "in_conv.bias": (ch[0],),
         "out_conv.0.weight": (ch[0],), "out_conv.0.bias": (ch[0],),
         "out_conv.2.weight": (arch["z_dim"], ch[0], 3, 3), "out_conv.2.bias": (arch["z_dim"],)}

    def block(p, cin, cout):
        for l in range(L):
            ci = cin if l == 0 else cout
            s[f"{p}.first_halfs.{l}.layers.0.weight"] = (ci,); s[f"{p}.first_halfs.{l}.layers.0.bias"] = (ci,)
            s[f"{p}.first_halfs.{l}.layers.2.weight"] = (cout, ci, 3, 3); s[f"{p}.first_halfs.{l}.layers.2.bias"] = (cout,)
            s[f"{p}.time_projs.{l}.1.weight"] = (cout, D); s[f"{p}.time_projs.{l}.1.bias"] = (cout,)
            s[f"{p}.second_halfs.{l}.layers.0.weight"] = (cout,); s[f"{p}.second_halfs.{l}.layers.0.bias"] = (cout,)
            s[f"{p}.second_halfs.{l}.layers.2.weight"] = (cout, cout, 3, 3); s[f"{p}.second_halfs.{l}.layers.2.bias"] = (cout,)
            s[f"{p}.residuals.{l}.weight"] = (cout, ci, 1, 1); s[f"{p}.residuals.{l}.bias"] = (cout,)
            a = f"{p}.self_attns.{l}"
            s[a + ".groupnorm.weight"] = (cout,); s[a + ".groupnorm.bias"] = (cout,)
            for n in ("to_q", "to_k", "to_v", "out_proj"):
                s[f"{a}.{n}.weight"] = (cout, cout); s[f"{a}.{n}.bias"] = (cout,)

    for i in range(len(ch) - 1):
        block(f"down_blocks.{i}", ch[i], ch[i + 1])
        s[f"downsamples.{i}.down.weight"] = (ch[i + 1], ch[i + 1], 3, 3); s[f"downsamples.{i}.down.bias"] = (ch[i + 1],)
    for i in range(len(mid) - 1):
        block(f"mid_blocks.{i}", mid[i], mid[i + 1])
    rev = ch[::-1]
    for i in range(len(ch) - 1):
        block(f"ups.{i}", rev[i] * 2, rev[i + 1])
        s[f"upsamples.{i}.conv.weight"] = (rev[i], rev[i], 3, 3); s[f"upsamples.{i}.conv.bias"] = (rev[i],)
    return s


def vae_param_shapes(arch) -> dict:
    s = {}

    def add(prefix, prog):
        for kind, i, cin, cout in prog:
            p = f"{prefix}.{i}"
            if kind == "conv1x1":
                s[p + ".weight"] = (cout, cin, 1, 1); s[p + ".bias"] = (cout,)
            elif kind == "conv3x3":
                s[p + ".weight"] = (cout, cin, 3, 3); s[p + ".bias"] = (cout,)
            elif kind == "res":
                s[p + ".branch.0.weight"] = (cin,); s[p + ".branch.0.bias"] = (cin,)
                s[p + ".branch.2.weight"] = (cout, cin, 3, 3); s[p + ".branch.2.bias"] = (cout,)
                s[p + ".branch.3.weight"] = (cout,); s[p + ".branch.3.bias"] = (cout,)
                s[p + ".branch.5.weight"] = (cout, cout, 3, 3); s[p + ".branch.5.bias"] = (cout,)
                if cin != cout:
                    s[p + ".residual_wrapper.weight"] = (cout, cin, 1, 1); s[p + ".residual_wrapper.bias"] = (cout,)
            elif kind == "attn":
                s[p + ".groupnorm.weight"] = (cin,); s[p + ".groupnorm.bias"] = (cin,)
                for n in ("to_q", "to_k", "to_v", "out_proj"):
                    s[f"{p}.{n}.weight"] = (cin, cin); s[f"{p}.{n}.bias"] = (cin,)
            elif kind == "up":
                s[p + ".conv.weight"] = (cout, cin, 3, 3); s[p + ".conv.bias"] = (cout,)
            elif kind == "down":
                s[p + ".down.weight"] = (cout, cin, 3, 3); s[p + ".down.bias"] = (cout,)
            elif kind == "gn_silu":
                s[p + ".weight"] = (cin,); s[p + ".bias"] = (cin,)

    add("encoder.down", encoder_program(arch))
    add("decoder.up", decoder_program(arch))
    if arch["bottleneck"] == "vq":
        s["codebook.ema_cluster_size"] = (arch["codebook_size"],)
        s["codebook.ema_w"] = (arch["codebook_size"], arch["z_dim"])
        s["codebook.embeddings.weight"] = (arch["codebook_size"], arch["z_dim"])
    return s


UNET_ARCH = dict(z_dim=3, channels=[128, 256, 384, 512], mid_channels=[512, 512], time_dim=512, num_res_layers=2,
                 num_heads=8, num_groups=32, num_classes=3)  # configs/diff-kl-lin-32x32.yaml:2-9
VAE_KL_ARCH = dict(in_channels=3, channels=[128, 256, 384], z_dim=3, bottleneck="kl", codebook_size=None,
                   codebook_beta=None, codebook_gamma=None, enc_num_res_blocks=2, dec_num_res_blocks=2,
                   attn_resolutions=[], num_heads=1, init_resolution=128, num_groups=32)  # configs/vae-kl-32x32.yaml
VAE_VQ_ARCH = dict(VAE_KL_ARCH, bottleneck="vq", codebook_size=1024, codebook_beta=0.25,
                   codebook_gamma=0.99)  # configs/vae-vq-32x32.yaml

"""Parameter trees of the drop-in modules: names, shapes and default initialisation, keyed exactly like the
reference's state_dict (SURVEY.md §8a "State-dict schema"), so reference checkpoints load unchanged."""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

# Architectures of the three shipped configurations (the values of the reference's YAML files).
UNET_ARCH = dict(z_dim=3, channels=[128, 256, 384, 512], mid_channels=[512, 512], time_dim=512, num_res_layers=2,
                 num_heads=8, num_groups=32, num_classes=3)  # configs/diff-kl-lin-32x32.yaml:2-9
VAE_KL_ARCH = dict(in_channels=3, channels=[128, 256, 384], z_dim=3, bottleneck="kl", codebook_size=None,
                   codebook_beta=None, codebook_gamma=None, enc_num_res_blocks=2, dec_num_res_blocks=2,
                   attn_resolutions=[], num_heads=1, init_resolution=128, num_groups=32)  # configs/vae-kl-32x32.yaml:2-15
VAE_VQ_ARCH = dict(VAE_KL_ARCH, bottleneck="vq", codebook_size=1024, codebook_beta=0.25,
                   codebook_gamma=0.99)  # configs/vae-vq-32x32.yaml:2-15

# kinds: conv/linear weight "w", its bias "b" (fan_in attached), norm gain "g", norm shift "z", embedding "e",
# codebook "c", buffers "factor" / "zeros"


def _conv(s, p, cout, cin, k):
    s[p + ".weight"] = ((cout, cin, k, k), "w", cin * k * k)
    s[p + ".bias"] = ((cout,), "b", cin * k * k)


def _lin(s, p, cout, cin):
    s[p + ".weight"] = ((cout, cin), "w", cin)
    s[p + ".bias"] = ((cout,), "b", cin)


def _gn(s, p, c):
    s[p + ".weight"] = ((c,), "g", 0)
    s[p + ".bias"] = ((c,), "z", 0)


def _attn(s, p, c):
    _gn(s, p + ".groupnorm", c)
    for n in ("to_q", "to_k", "to_v", "out_proj"):
        _lin(s, f"{p}.{n}", c, c)


def unet_blocks(arch):
    """[(prefix, cin, cout, resolution_divisor)] for every DiffusionBlock in execution order, plus the
    down/up-sample prefixes — the single source of truth for both parameters and the execution plan."""
    ch, mid = list(arch["channels"]), list(arch["mid_channels"])
    rev = ch[::-1]
    downs = [(f"down_blocks.{i}", ch[i], ch[i + 1]) for i in range(len(ch) - 1)]
    mids = [(f"mid_blocks.{i}", mid[i], mid[i + 1]) for i in range(len(mid) - 1)]
    ups = [(f"ups.{i}", rev[i] * 2, rev[i + 1]) for i in range(len(ch) - 1)]
    return downs, mids, ups


def unet_param_spec(arch) -> OrderedDict:
    ch, D, L = list(arch["channels"]), arch["time_dim"], arch["num_res_layers"]
    s: OrderedDict = OrderedDict()
    s["class_embedding.weight"] = ((arch["num_classes"], D), "e", 0)
    s["time_embedding.factor"] = ((D // 2,), "factor", 0)
    _lin(s, "time_embedding.embeddings.0", 4 * D, D)
    _lin(s, "time_embedding.embeddings.2", D, 4 * D)
    _conv(s, "in_conv", ch[0], arch["z_dim"], 3)
    downs, mids, ups = unet_blocks(arch)

    def block(p, cin, cout):
        for l in range(L):
            ci = cin if l == 0 else cout
            _gn(s, f"{p}.first_halfs.{l}.layers.0", ci)
            _conv(s, f"{p}.first_halfs.{l}.layers.2", cout, ci, 3)
        for l in range(L):
            _lin(s, f"{p}.time_projs.{l}.1", cout, D)
        for l in range(L):
            _gn(s, f"{p}.second_halfs.{l}.layers.0", cout)
            _conv(s, f"{p}.second_halfs.{l}.layers.2", cout, cout, 3)
        for l in range(L):
            _conv(s, f"{p}.residuals.{l}", cout, cin if l == 0 else cout, 1)
        for l in range(L):
            _attn(s, f"{p}.self_attns.{l}", cout)

    for p, cin, cout in downs:
        block(p, cin, cout)
    for i in range(len(ch) - 1):
        _conv(s, f"downsamples.{i}.down", ch[i + 1], ch[i + 1], 3)
    for p, cin, cout in mids:
        block(p, cin, cout)
    for p, cin, cout in ups:
        block(p, cin, cout)
    rev = ch[::-1]
    for i in range(len(ch) - 1):
        _conv(s, f"upsamples.{i}.conv", rev[i], rev[i], 3)
    _gn(s, "out_conv.0", ch[0])
    _conv(s, "out_conv.2", arch["z_dim"], ch[0], 3)
    return s


def vae_program(arch, part: str):
    """Layer program of Encoder.down / Decoder.up: (kind, sequential index, cin, cout)."""
    nres_e, nres_d, attn_res = arch["enc_num_res_blocks"], arch["dec_num_res_blocks"], arch["attn_resolutions"]
    prog = []
    if part == "encoder":
        ch = list(arch["channels"])
        zc = arch["z_dim"] if arch["bottleneck"] == "vq" else 2 * arch["z_dim"]
        res = arch["init_resolution"]
        prog.append(("conv3x3", 0, arch["in_channels"], ch[0]))
        i = 1
        for s_ in range(len(ch) - 1):
            cin = ch[s_]
            for _ in range(nres_e):
                prog.append(("res", i, cin, ch[s_ + 1])); i += 1
                cin = ch[s_ + 1]
            if res in attn_res:
                prog.append(("attn", i, ch[s_ + 1], ch[s_ + 1])); i += 1
            prog.append(("down", i, ch[s_ + 1], ch[s_ + 1])); i += 1
            res /= 2
        for _ in range(nres_e):
            prog.append(("res", i, ch[-1], ch[-1])); i += 1
        prog.append(("attn", i, ch[-1], ch[-1])); i += 1
        for _ in range(nres_e):
            prog.append(("res", i, ch[-1], ch[-1])); i += 1
        prog.append(("gn_silu", i, ch[-1], ch[-1])); i += 2
        prog.append(("conv3x3", i, ch[-1], zc)); i += 1
        prog.append(("conv1x1", i, zc, zc))
    else:
        ch = list(arch["channels"])[::-1]
        res = arch["init_resolution"] // 2 ** len(arch["channels"])
        prog += [("conv1x1", 0, arch["z_dim"], arch["z_dim"]), ("conv3x3", 1, arch["z_dim"], ch[0])]
        i = 2
        for _ in range(nres_d):
            prog.append(("res", i, ch[0], ch[0])); i += 1
        prog.append(("attn", i, ch[0], ch[0])); i += 1
        for _ in range(nres_d):
            prog.append(("res", i, ch[0], ch[0])); i += 1
        for s_ in range(len(ch) - 1):
            cin = ch[s_]
            for _ in range(nres_d):
                prog.append(("res", i, cin, ch[s_ + 1])); i += 1
                cin = ch[s_ + 1]
            if res in attn_res:
                prog.append(("attn", i, ch[s_ + 1], ch[s_ + 1])); i += 1
            prog.append(("up", i, ch[s_ + 1], ch[s_ + 1])); i += 1
            res *= 2
        for _ in range(nres_d):
            prog.append(("res", i, ch[-1], ch[-1])); i += 1
        prog.append(("gn_silu", i, ch[-1], ch[-1])); i += 2
        prog.append(("conv3x3", i, ch[-1], arch["in_channels"]))
    return prog


def vae_param_spec(arch) -> OrderedDict:
    s: OrderedDict = OrderedDict()
    for part, prefix in (("encoder", "encoder.down"), ("decoder", "decoder.up")):
        for kind, i, cin, cout in vae_program(arch, part):
            p = f"{prefix}.{i}"
            if kind == "conv1x1":
                _conv(s, p, cout, cin, 1)
            elif kind == "conv3x3":
                _conv(s, p, cout, cin, 3)
            elif kind == "res":
                _gn(s, p + ".branch.0", cin)
                _conv(s, p + ".branch.2", cout, cin, 3)
                _gn(s, p + ".branch.3", cout)
                _conv(s, p + ".branch.5", cout, cout, 3)
                if cin != cout:
                    _conv(s, p + ".residual_wrapper", cout, cin, 1)
            elif kind == "attn":
                _attn(s, p, cin)
            elif kind == "up":
                _conv(s, p + ".conv", cout, cin, 3)
            elif kind == "down":
                _conv(s, p + ".down", cout, cin, 3)
            elif kind == "gn_silu":
                _gn(s, p, cin)
    if arch["bottleneck"] == "vq":
        s["codebook.embeddings.weight"] = ((arch["codebook_size"], arch["z_dim"]), "c", 0)
        s["codebook.ema_cluster_size"] = ((arch["codebook_size"],), "zeros", 0)
        s["codebook.ema_w"] = ((arch["codebook_size"], arch["z_dim"]), "c", 0)
    return s


def _init_tensor(shape, kind, fan_in) -> torch.Tensor:
    if kind in ("w", "b"):
        bound = 1.0 / math.sqrt(fan_in)  # torch's default Conv2d/Linear init (kaiming_uniform with a = sqrt(5))
        return torch.empty(shape).uniform_(-bound, bound)
    if kind == "g":
        return torch.ones(shape)
    if kind in ("z", "zeros"):
        return torch.zeros(shape)
    if kind == "e":
        return torch.randn(shape)
    if kind == "c":
        return torch.empty(shape).uniform_(-1.0 / shape[0], 1.0 / shape[0])
    if kind == "factor":
        half = shape[0]
        return 10000 ** (torch.arange(0, half, dtype=torch.float32) / half)
    raise ValueError(kind)


def register_tree(root: nn.Module, spec: OrderedDict, prefix: str = "") -> None:
    """Creates nested container modules so that root.state_dict() has exactly the keys of `spec`."""
    for name, (shape, kind, fan_in) in spec.items():
        parts = (prefix + name).split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, nn.Module())
            node = node._modules[p]
        t = _init_tensor(tuple(shape), kind, fan_in)
        if kind in ("factor", "zeros"):
            node.register_buffer(parts[-1], t)
        else:
            node.register_parameter(parts[-1], nn.Parameter(t))

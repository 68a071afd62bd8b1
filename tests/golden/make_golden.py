"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, read-only) on seeded weights
and inputs. Run in the build container only:  python tests/golden/make_golden.py

The weights come from oracle.ref_path.seeded_state_dict (a recipe that needs no reference code), loaded into the
reference classes with load_state_dict(strict=True) — which also proves that the oracle's parameter-shape specs
equal the reference's state_dict schema. Only inputs that cannot be regenerated from a seed and the reference
OUTPUTS are stored, so the fixtures stay small.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from modules.components import Codebook, Scheduler  # noqa: E402  (reference)
from modules.unet import Unet  # noqa: E402  (reference)
from modules.vae import VAE  # noqa: E402  (reference)

from oracle import ref_path as O  # noqa: E402

torch.manual_seed(0)
torch.set_grad_enabled(False)

TINY_UNET = dict(z_dim=3, channels=[32, 64, 96], mid_channels=[96, 96], time_dim=64, num_res_layers=1, num_heads=2,
                 num_groups=8, num_classes=3)
TINY_VAE_VQ = dict(in_channels=3, channels=[32, 64], z_dim=3, bottleneck="vq", codebook_size=64, codebook_beta=0.25,
                   codebook_gamma=0.99, enc_num_res_blocks=1, dec_num_res_blocks=1, attn_resolutions=[16],
                   num_heads=2, init_resolution=32, num_groups=8)
TINY_VAE_KL = dict(TINY_VAE_VQ, bottleneck="kl", codebook_size=None, codebook_beta=None, codebook_gamma=None)


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def build_unet(arch, seed):
    m = Unet(**arch).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {k: tuple(v) for k, v in O.unet_param_shapes(arch).items()}, "UNet state_dict schema mismatch"
    m.load_state_dict(O.seeded_state_dict(shapes, seed), strict=True)
    return m


def build_vae(arch, seed):
    m = VAE(**arch).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {k: tuple(v) for k, v in O.vae_param_shapes(arch).items()}, "VAE state_dict schema mismatch"
    m.load_state_dict(O.seeded_state_dict(shapes, seed), strict=True)
    return m


def save(name, obj):
    path = os.path.join(HERE, name)
    torch.save(obj, path)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def unet_cases():
    out = {}
    for tag, arch, res, seed in (("tiny", TINY_UNET, 16, 11), ("full", O.UNET_ARCH, 32, 2018)):
        m = build_unet(arch, seed)
        B = 3 if tag == "tiny" else 2
        x = gen(100 + B, B, 3, res, res)
        t = torch.tensor([0, 500, 999][:B])
        ctx = torch.tensor([2, 0, 1][:B])
        mask = torch.tensor([[1.0], [0.0], [1.0]][:B])
        out[tag] = dict(arch=arch, seed=seed, x_seed=100 + B, shape=(B, 3, res, res), t=t, ctx=ctx, mask=mask,
                        cond=m(x, t, ctx), uncond=m(x, t), masked=m(x, t, ctx, mask.clone()))
    save("unet.pt", out)


def scheduler_cases():
    out = {}
    for typ in ("linear", "cosine"):
        s = Scheduler(1000, 1e-4, 0.02, typ)
        o = O.SchedulerTables(1000, 1e-4, 0.02, typ)
        for name in ("betas", "alphas", "alpha_cum_prod", "sqrt_alpha_cum_prod", "sqrt_one_minus_alpha_cum_prod"):
            assert torch.equal(getattr(s, name), getattr(o, name)), (typ, name)
        out[typ] = dict(betas=s.betas.clone(), alpha_cum_prod=s.alpha_cum_prod.clone())
    s = Scheduler(1000)
    xt, eps, z = gen(1, 4, 3, 8, 8), gen(2, 4, 3, 8, 8), gen(3, 4, 3, 8, 8)
    steps = {}
    orig = torch.randn_like
    torch.randn_like = lambda *_a, **_k: z  # the reference draws inside sample_prev_timestep (components.py:423)
    try:
        for i in (999, 500, 1, 0):
            t = torch.full((4,), i, dtype=torch.long)
            xp, x0 = s.sample_prev_timestep(xt, eps, t)
            steps[i] = dict(x_prev=xp, x0=x0)
    finally:
        torch.randn_like = orig
    tn = torch.tensor([0, 10, 500, 999])
    out["posterior"] = dict(seeds=(1, 2, 3), shape=(4, 3, 8, 8), steps=steps)
    out["add_noise"] = dict(t=tn, out=s.add_noise(xt, eps, tn))
    save("scheduler.pt", out)


def sample_loop_case():
    """diffusion.py:46-56 re-driven on CPU with the reference Unet + Scheduler and injected noise."""
    m = build_unet(TINY_UNET, 11)
    s = Scheduler(1000)
    N, res = 6, 16
    xt = gen(40, N, 3, res, res)
    labels = torch.tensor([0, 1, 2] * 2)
    cfg = torch.tensor([1, 3, 7] * 2)[:, None, None, None]
    steps = [999, 998, 500, 2, 1, 0]
    noises = [gen(50 + k, N, 3, res, res) for k in range(len(steps))]
    trace = []
    orig = torch.randn_like
    try:
        for k, i in enumerate(steps):
            torch.randn_like = lambda *_a, _z=noises[k], **_k: _z
            t = torch.full((N,), i, dtype=torch.long)
            ec, eu = m(xt, t, labels), m(xt, t)
            eps = eu + cfg * (ec - eu)
            xt, _ = s.sample_prev_timestep(xt, eps, t)
            trace.append(xt)
    finally:
        torch.randn_like = orig
    save("sample_loop.pt", dict(arch=TINY_UNET, seed=11, x_seed=40, noise_seed0=50, shape=(N, 3, res, res),
                                labels=labels, cfg=cfg.view(-1), steps=steps, trace=torch.stack(trace)))


def vae_cases():
    out = {}
    # full-size KL decoder / encoder, one image
    m = build_vae(O.VAE_KL_ARCH, 2018)
    z = gen(60, 1, 3, 32, 32)
    img = gen(61, 1, 3, 128, 128).clamp(-1, 1)
    zenc, kl, _ = m.encode(img, sample=False)
    out["kl_full"] = dict(arch=O.VAE_KL_ARCH, seed=2018, z_seed=60, img_seed=61, decode=m.decode(z), encode=zenc,
                          kl=kl)
    # tiny VQ / KL with an attention layer in the up/down path too
    mv = build_vae(TINY_VAE_VQ, 5)
    img = gen(62, 2, 3, 32, 32).clamp(-1, 1)
    zq, loss, perp = mv.encode(img)
    out["vq_tiny"] = dict(arch=TINY_VAE_VQ, seed=5, img_seed=62, zq=zq, loss=loss, perplexity=perp,
                          decode=mv.decode(zq), decode_requant=mv.decode(gen(63, 2, 3, 16, 16) * 0.01, quantize=True),
                          forward=mv(img))
    mk = build_vae(TINY_VAE_KL, 6)
    zk, klk, _ = mk.encode(img, sample=False)
    out["kl_tiny"] = dict(arch=TINY_VAE_KL, seed=6, img_seed=62, encode=zk, kl=klk, decode=mk.decode(zk[:, :3]))
    # full-size VQ: quantiser on a realistic and on an adversarial (default-init) codebook
    cbk = Codebook(1024, 3, 0.25, 0.99).eval()
    zz = gen(64, 2, 3, 32, 32)
    for tag, w in (("default", (torch.rand(1024, 3, generator=torch.Generator().manual_seed(65)) * 2 - 1) / 1024),
                   ("normal", gen(66, 1024, 3))):
        cbk.embeddings.weight.data.copy_(w)
        q, loss, perp = cbk(zz)
        x = zz.permute(0, 2, 3, 1).reshape(2, 1024, 3)
        idx = torch.cdist(x, w[None].repeat(2, 1, 1)).argmin(-1).view(-1)
        out["codebook_" + tag] = dict(z_seed=64, w=w, zq=q, loss=loss, perplexity=perp, idx=idx.to(torch.int16))
    save("vae.pt", out)


def train_step_case():
    """trainers/diffusion_trainer.py:141-173 restated around the reference Unet/Scheduler (the trainer module itself
    needs mlflow/matplotlib and cannot be imported); fp32, no autocast, loss + a few gradients."""
    torch.set_grad_enabled(True)
    m = build_unet(TINY_UNET, 11).train()
    s = Scheduler(1000)
    B, res = 4, 16
    lat = gen(70, B, 6, res, res)
    c = torch.tensor([0, 2, 1, 1])
    mean, log_var = torch.chunk(lat, 2, dim=1)
    x = mean + gen(71, B, 3, res, res) * torch.exp(0.5 * torch.clamp(log_var, -30.0, 20.0))
    noise = gen(72, B, 3, res, res)
    t = torch.tensor([3, 250, 700, 999])
    mask = torch.tensor([[True], [False], [True], [True]])
    x_noise = s.add_noise(x, noise, t)
    pred = m(x_noise, t, context=c, context_mask=mask)
    loss = torch.nn.MSELoss()(pred, noise)
    loss.backward()
    names = ["in_conv.weight", "out_conv.2.bias", "mid_blocks.0.self_attns.0.to_q.weight",
             "down_blocks.0.time_projs.0.1.weight", "class_embedding.weight", "ups.1.residuals.0.weight"]
    grads = {n: dict(m.named_parameters())[n].grad.clone() for n in names}
    torch.set_grad_enabled(False)
    save("train_step.pt", dict(arch=TINY_UNET, seed=11, seeds=(70, 71, 72), shape=(B, 6, res, res), labels=c, t=t,
                               mask=mask, loss=loss.detach(), grads=grads))


if __name__ == "__main__":
    unet_cases()
    scheduler_cases()
    sample_loop_case()
    vae_cases()
    train_step_case()

"""Round-2 behaviours on the B200: the strided (DDIM-style) sampler of SURVEY §8 row f4, the out-of-range timestep
wrap of the posterior kernel, and weight-staleness handling between training and sampling (cached samplers, captured
graphs and packed bf16 weights must follow optimizer steps and load_state_dict)."""
import pytest
import torch

from oracle import ref_path as O

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"
SMALL = dict(z_dim=3, channels=[128, 256], mid_channels=[256, 256], time_dim=128, num_res_layers=1, num_heads=4,
             num_groups=32, num_classes=3)


def gen(seed, *shape):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def rel_rms(a, b):
    return ((a - b).norm() / b.norm()).item()


def make_unet(arch, seed):
    from modules.unet import Unet
    sd = O.seeded_state_dict(O.unet_param_shapes(arch), seed)
    m = Unet(**arch)
    m.load_state_dict(sd)
    return m.to(DEV), {k: v.to(DEV) for k, v in sd.items()}


@torch.no_grad()
def test_ddim_step_kernel_matches_published_formula():
    from modules.components import Scheduler
    s = Scheduler(1000, device=DEV)
    os_ = O.SchedulerTables(1000, device=DEV)
    xt, eps, z = (gen(k, 5, 3, 32, 32).to(DEV) for k in (1, 2, 3))
    for t, tp, eta in ((999, 949, 0.0), (500, 450, 0.0), (50, 0, 0.0), (0, -1, 0.0), (999, 998, 1.0), (300, 100, 0.5),
                       (7, -1, 1.0)):
        got, x0 = s.ddim_step(xt, eps, t, tp, eta=eta, noise=z)
        ref, x0_ref = O.ddim_step(os_, xt, eps, t, tp, eta, z)
        tol = 2e-5 * ref.abs().max().item()
        assert (got - ref).abs().max().item() <= tol, (t, tp, eta)
        assert (x0 - x0_ref).abs().max().item() <= 2e-5 * x0_ref.abs().max().item()
    # eta = 1, stride 1 == the reference's ancestral step (components.py:405-424)
    t = torch.full((5,), 640, device=DEV)
    ref = O.posterior_step(os_, xt, eps, t, z)[0]
    got, _ = s.ddim_step(xt, eps, 640, 639, eta=1.0, noise=z)
    assert (got - ref).abs().max().item() <= 5e-5 * ref.abs().max().item()


@torch.no_grad()
def test_strided_cfg_sampler_matches_oracle():
    """50-step deterministic DDIM over the 1000-step schedule (every 20th timestep) and a 10-step eta = 1 run."""
    from idf_b200.sampler import CfgSampler
    from modules.components import Scheduler
    m, sd = make_unet(O.UNET_ARCH, 2018)
    m.eval()
    N = 6
    labels = torch.tensor([0, 1, 2] * 2, device=DEV)
    cfg = torch.tensor([3, 3, 3, 7, 7, 7], device=DEV)
    osched = O.SchedulerTables(1000, device=DEV)
    x_T = gen(11, N, 3, 32, 32).to(DEV)
    steps = list(range(999, -1, -20))
    smp = CfgSampler(m, Scheduler(1000, device=DEV), labels, cfg, (3, 32, 32), kind="ddim", eta=0.0)
    got = smp.run(x_T, steps=steps).clone()
    ref = O.cfg_sample_strided(sd, O.UNET_ARCH, osched, x_T, labels, cfg, steps, eta=0.0)
    r = rel_rms(got, ref)
    print(f"DDIM 50 steps: rel-RMS {r:.3e}")
    assert r <= 2e-2, r
    steps = list(range(999, 0, -111)) + [0]   # ... 111, 0: the step INTO timestep 0 still takes noise (sigma > 0)
    noises = [gen(20 + k, N, 3, 32, 32).to(DEV) for k in range(len(steps))]
    smp = CfgSampler(m, Scheduler(1000, device=DEV), labels, cfg, (3, 32, 32), kind="ddim", eta=1.0)
    got = smp.run(x_T, steps=steps, noises=noises).clone()
    ref = O.cfg_sample_strided(sd, O.UNET_ARCH, osched, x_T, labels, cfg, steps, eta=1.0, noises=noises)
    r = rel_rms(got, ref)
    print(f"strided ancestral (eta=1) {len(steps)} steps: rel-RMS {r:.3e}")
    assert r <= 2e-2, r


@torch.no_grad()
def test_posterior_per_sample_timestep_zero_wraps_like_the_reference():
    """Per-sample t with t[0] != 0 and some t[n] == 0: the reference indexes alpha_cum_prod[-1] (components.py:419)."""
    from modules.components import Scheduler
    s = Scheduler(1000, device=DEV)
    xt, eps = gen(1, 4, 3, 32, 32).to(DEV), gen(2, 4, 3, 32, 32).to(DEV)
    t = torch.tensor([5, 0, 999, 0], device=DEV)
    state = torch.cuda.get_rng_state()
    got, _ = s.sample_prev_timestep(xt, eps, t)
    torch.cuda.set_rng_state(state)
    z = torch.randn_like(xt)
    ref = O.posterior_step(O.SchedulerTables(1000, device=DEV), xt, eps, t, z)[0]
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_sampling_follows_training_and_load_state_dict():
    """train -> sample -> train -> sample: the fused Adam kernel updates parameters through a raw pointer, so the
    inference engine's packed weights, the cached sampler's embedding table and its captured graph must be refreshed
    (round-1 advisor finding). Checked against the oracle evaluated on the CURRENT parameters each time."""
    from idf_b200.sampler import CfgSampler
    from idf_b200.trainer import DiffusionTrainStep
    from modules.components import Scheduler
    m, _ = make_unet(SMALL, 3)
    sched = Scheduler(1000, device=DEV)
    osched = O.SchedulerTables(1000, device=DEV)
    N = 3
    labels = torch.tensor([0, 1, 2], device=DEV)
    cfg = torch.full((N,), 3, device=DEV)
    x_T = gen(5, N, 3, 16, 16).to(DEV)
    noises = [gen(6, N, 3, 16, 16).to(DEV), gen(7, N, 3, 16, 16).to(DEV)]
    steps = [700, 699]
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(4, 6, 16, 16, generator=g).to(DEV)
    lab = torch.randint(0, 3, (4,), generator=g).to(DEV)

    def check(tag, sampler):
        with torch.no_grad():
            sd_now = {k: v.detach().clone() for k, v in m.state_dict().items()}
            got = sampler.run(x_T, steps=steps, noises=noises).clone()
            ref = O.cfg_sample(sd_now, SMALL, osched, x_T, labels, cfg, noises, steps=steps)
            r = rel_rms(got, ref)
            print(f"{tag}: rel-RMS {r:.3e}")
            assert r <= 1.5e-2, (tag, r)
            return got

    with torch.no_grad():
        sampler = CfgSampler(m.eval(), sched, labels, cfg, (3, 16, 16))
        a = check("before training", sampler)
    ts = DiffusionTrainStep(m.train(), sched, 4, (3, 16, 16), clip_grad=1.0)
    for _ in range(4):
        ts.step(lat, lab, 1e-2)   # large lr: the weights move far beyond the parity tolerance
    m.eval()
    b = check("after 4 optimizer steps (same cached sampler and graph)", sampler)
    assert rel_rms(a, b) > 3e-2  # the update moved the output beyond the parity gate: a stale engine fails the check above
    with torch.no_grad():
        out = m(x_T, torch.full((N,), 700, device=DEV), labels)   # plain Unet.forward sees the new weights too
        ref = O.unet_forward({k: v.detach() for k, v in m.state_dict().items()}, SMALL, x_T,
                             torch.full((N,), 700, device=DEV), labels)
        assert rel_rms(out, ref) <= 3e-2
    ts.step(lat, lab, 1e-2)
    check("after one more step", sampler)
    # load_state_dict into the same module (in-place copies: version counters bump)
    sd2 = O.seeded_state_dict(O.unet_param_shapes(SMALL), 77)
    m.load_state_dict(sd2)
    c = check("after load_state_dict", sampler)
    assert rel_rms(b, c) > 2e-2

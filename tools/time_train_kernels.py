"""Per-launch CUDA-event timing of one eager training step at batch 48 (every C-ABI call with its shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from bench import build_models, BATCH
from idf_b200 import native, ops
from idf_b200.trainer import DiffusionTrainStep
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
unet, _, sched = build_models("cuda")
unet.train()
ts = DiffusionTrainStep(unet, sched, BATCH, (3, 32, 32), use_graph=False)
lat = torch.randn(BATCH, 6, 32, 32, device="cuda"); lab = torch.randint(0, 3, (BATCH,), device="cuda")
ts.step(lat, lab, 0.0); ts.step(lat, lab, 0.0)
torch.cuda.synchronize()
orig = native.call
recs = []
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    desc = ""
    if name == "idf_conv2d_igemm":
        g = a[0]
        m = (g.s2_batch if g.s2_batch else g.a[0].n) * g.a[0].h * g.a[0].w
        k = g.taps[0] * g.a[0].c + (g.taps[1] * g.a[1].c if g.a[1].ptr else 0)
        desc = f"M={m} N={g.N} K={k} taps={g.taps[0]} res={bool(g.res)} vt={bool(g.vt)} f={2.0*m*g.N*k/1e9:.1f}GF"
    elif name == "idf_conv2d_wgrad":
        g = a[0]
        m = (g.s2_batch if g.s2_batch else g.x.n) * g.x.h * g.x.w
        desc = f"M={m} Cout={g.cout} Cin={g.x.c} taps={g.taps} f={2.0*m*g.cout*g.x.c*g.taps/1e9:.1f}GF"
    elif name in ("idf_attention_fwd_train",):
        desc = f"M={a[6]} T={a[7]} heads={a[8]} hd={a[9]}"
    elif name == "idf_attention_bwd":
        desc = f"M={a[-5]} T={a[-4]} heads={a[-3]} hd={a[-2]}"
    elif name in ("idf_groupnorm_silu_train",):
        desc = f"B={a[6]} HW={a[7]} C={a[8]} silu={a[11]}"
    elif name == "idf_groupnorm_silu_bwd":
        desc = f"B={a[13]} HW={a[14]} C={a[15]} silu={a[17]}"
    elif name == "idf_colsum_bf16":
        desc = f"B={a[2]} HW={a[3]} C={a[4]}"
    recs.append((name, desc, e0, e1))
native.call = timed; ops.call = timed
acc = {}
for rep in range(reps):
    recs.clear()
    torch.cuda._sleep(int(8e7))
    ts._whole_step()
    torch.cuda.synchronize()
    for i, (n, d, e0, e1) in enumerate(recs):
        acc.setdefault(i, [n, d, []])[2].append(e0.elapsed_time(e1) * 1e3)
tot = 0; agg = {}
for i in sorted(acc):
    n, d, tsl = acc[i]
    t = sorted(tsl)[len(tsl) // 2]
    tot += t
    agg[n] = agg.get(n, 0) + t
    extra = ""
    if "GF" in d:
        gf = float(d.split("f=")[1][:-2]); extra = f" {gf / t * 1e3:8.0f} TF/s"
    print(f"{i:3d} {n[4:]:24s} {t:8.1f} us  {d}{extra}")
print("total us", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"  {k:28s} {v:9.1f} us {100 * v / tot:5.1f}%")

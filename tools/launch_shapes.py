"""Shapes of every C-ABI call of one CFG UNet forward, recorded on the CPU (the library call is replaced by a
recorder, nothing is computed), joined with an ncu launch list of one eager step to give per-shape time, FLOP/s and
bytes/s:  python tools/launch_shapes.py [batch] [profiles/r02_sample_step_launches.csv]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "image-diffusion_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from idf_b200 import ops, spec  # noqa: E402

calls = []


def recorder(name, *args):
    if name == "idf_conv2d_igemm":
        a = args[0]
        M = a.a[0].n * a.a[0].h * a.a[0].w
        K = a.a[0].c * a.taps[0] + (a.a[1].c * a.taps[1] if a.a[1].ptr else 0)
        calls.append(dict(kernel="igemm", M=M, N=a.N, K=K, taps=a.taps[0], seg2=bool(a.a[1].ptr), s2=a.s2_direct,
                          up2=a.out_up2, splits=a.force_splits, f32=a.out_f32, hw=a.a[0].h))
    elif name == "idf_attention_fwd_qkv":
        calls.append(dict(kernel="attention", M=args[4], T=args[5], heads=args[6], hd=args[7]))
    elif name == "idf_groupnorm_silu":
        calls.append(dict(kernel="groupnorm", B=args[6], HW=args[7], C=args[8], silu=args[11]))
    else:
        calls.append(dict(kernel=name))
    return 0


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    ops.call = recorder
    from idf_b200 import engine as E
    from modules.unet import Unet
    m = Unet(**spec.UNET_ARCH).eval()
    eng = E.UnetEngine(m, spec.UNET_ARCH, "cpu")
    x = torch.zeros(B, 3, 32, 32)
    out = torch.zeros(B, 3, 32, 32)
    try:
        table = torch.zeros(4, eng.P, dtype=torch.bfloat16)
        calls.clear()
        eng.run(x, None, None, None, torch.zeros(B, dtype=torch.int32), out, table=table)
    except Exception as e:  # noqa: BLE001
        print("run stopped:", repr(e)[:300], file=sys.stderr)
    print(len(calls), "calls", file=sys.stderr)
    json.dump(calls, open("/tmp/launch_shapes.json", "w"))
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_sample_step_launches.csv")
    rows = [r for r in csv.DictReader(open(path))]
    fam = {"igemm": "igemm_persist", "attention": "attention", "groupnorm": "groupnorm_kernel"}
    it = {k: iter([r for r in rows if v in r["kernel"]]) for k, v in fam.items()}
    fin = iter([r for r in rows if "splitk_finish" in r["kernel"]])
    agg = {}
    for c in calls:
        k = c["kernel"]
        if k not in fam:
            continue
        r = next(it[k], None)
        if r is None:
            continue
        us = float(r["duration_us"])
        dram = float(r["dram_read_MB"]) + float(r["dram_write_MB"])
        if k == "igemm":
            if c["splits"] > 1:
                us += float(next(fin)["duration_us"])
            flop = 2.0 * c["M"] * c["N"] * c["K"] * (4 if c["up2"] == 2 else 1) / (4 if c["s2"] else 1)
            mb = (c["M"] * (c["K"] // c["taps"] if not c["seg2"] else 0) + c["M"] * c["N"]) * 2 / 1e6
            key = ("igemm", c["hw"] if c["taps"] > 1 else "1x1", c["M"], c["N"], c["K"], c["up2"], c["s2"], c["splits"], r["kernel"][21:36])
        elif k == "attention":
            C = c["heads"] * c["hd"]
            flop = 4.0 * c["M"] * c["T"] * C
            mb = c["M"] * C * 4 * 2 / 1e6
            key = ("attention", c["T"], c["M"], c["heads"], c["hd"], r["kernel"][:24])
        else:
            flop = 0.0
            mb = c["B"] * c["HW"] * c["C"] * 4 / 1e6
            key = ("groupnorm", c["B"], c["HW"], c["C"], c["silu"], r["kernel"][:22])
        a = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += flop; a[3] += mb; a[4] += dram
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot:.1f} us")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:8.1f} us {100 * a[1] / tot:5.1f}% n={a[0]:2d} {a[2] / a[1] / 1e6:7.1f} TF/s  min-bytes {a[3] / a[1] * 1e3:6.0f} GB/s  dram {a[4] / a[1] * 1e3:6.0f} GB/s  {key}")

"""Does a low-resolution conv run faster when its weights are already in L2? (cold: a 512 MB write between runs evicts
them; warm: back-to-back). Decides whether prefetching the next layer's weights into L2 is worth building."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
from idf_b200 import ops
dev = "cuda"
flush = torch.empty(512 * 1024 * 1024, device=dev, dtype=torch.uint8)
for (B, H, Cin, Cout, splits) in ((96, 4, 512, 512, 3), (96, 8, 512, 512, 0), (96, 8, 384, 384, 0), (96, 16, 384, 384, 0)):
    x = torch.randn(B * H * H, Cin, device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin))
    b = torch.randn(Cout, device=dev)
    out = torch.empty(B * H * H, Cout, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(4 * B * H * H * Cout, device=dev, dtype=torch.float32) if splits else None
    kw = dict(ws=ws, splits=splits) if splits else {}
    run = lambda: ops.igemm([(x, (B, H, H), Cin, 9)], w, Cout, out, bias=b, **kw)
    for _ in range(3): run()
    torch.cuda.synchronize()
    res = {}
    for mode in ("warm", "cold"):
        ts = []
        for _ in range(10):
            if mode == "cold":
                flush.fill_(1)
                x.add_(0)  # activations back into L2 (the producer just wrote them in the real step)
            else:
                run()      # weights and activations L2-resident
            torch.cuda._sleep(int(4e7))  # keeps the GPU busy while the host enqueues: events see GPU time only
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[mode] = sorted(ts)[len(ts) // 2]
    print(f"B={B} {H}x{H} {Cin}->{Cout} splits={splits}: warm {res['warm']:.1f} us, cold weights {res['cold']:.1f} us ({w.numel() * 2 / 1e6:.1f} MB of weights)")

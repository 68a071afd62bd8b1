"""Data-parallel training-step check (run under torchrun, >= 2 ranks, NCCL): gradients after the bucketed all-reduce
must equal (rel-RMS <= 1e-2) the single-GPU gradients of the concatenated batch, and all ranks must hold
identical parameters after the Adam step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "image-diffusion_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device(dev))
from idf_b200.trainer import DiffusionTrainStep
from modules.components import Scheduler
from modules.unet import Unet
from idf_b200.spec import UNET_ARCH

b = 6
g = torch.Generator().manual_seed(7)
lat = torch.randn(world * b, 6, 32, 32, generator=g).to(dev)
rn = torch.randn(world * b, 3, 32, 32, generator=g).to(dev)
noise = torch.randn(world * b, 3, 32, 32, generator=g).to(dev)
t = torch.randint(0, 1000, (world * b,), generator=g).to(dev)
lab = torch.randint(0, 3, (world * b,), generator=g).to(dev)
mask = (torch.rand(world * b, generator=g) > 0.15).float().to(dev)
sched = Scheduler(1000, device=dev)


def make(batch, data_parallel=True):
    torch.manual_seed(2018 + (rank if data_parallel else 0))  # rank-dependent init: the constructor must broadcast rank 0's
    m = Unet(**UNET_ARCH).to(dev).train()
    return m, DiffusionTrainStep(m, sched, batch, (3, 32, 32), clip_grad=1.0, data_parallel=data_parallel,
                                 use_graph=(os.environ.get("DDP_CHECK_GRAPH", "1") == "1"))


def load(ts, sl):
    ts.reparam_noise.copy_(rn[sl]); ts.noise.copy_(noise[sl]); ts.t.copy_(t[sl]); ts.mask.copy_(mask[sl])


m, ts = make(b)
sl = slice(rank * b, (rank + 1) * b)
load(ts, sl)
loss = ts.step(lat[sl], lab[sl], 1e-3, draw=False)
torch.cuda.synchronize()
g_ddp = ts.eng.flat_grad.clone() / world
ok = True
if rank == 0:
    # single-GPU reference on the concatenated batch (no process group use: world forced to 1)
    m1, ts1 = make(world * b, data_parallel=False)
    load(ts1, slice(0, world * b))
    ts1.step(lat, lab, 1e-3, draw=False)
    torch.cuda.synchronize()
    g1 = ts1.eng.flat_grad
    rel = ((g_ddp - g1).norm() / g1.norm()).item()
    dp = ((ts.flat_param - ts1.flat_param).norm() / (ts1.flat_param.norm())).item()
    print(f"ddp_check: world {world}, grad rel-RMS vs single-GPU concatenated batch {rel:.3e}, "
          f"grad norm {ts.grad_norm.item():.4f} vs {ts1.grad_norm.item():.4f}, param diff after Adam {dp:.3e}")
    ok = rel <= 1e-2
# all ranks hold the same parameters
chk = ts.flat_param.double().sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool((lo == hi).item())
if rank == 0:
    print(f"ddp_check: parameters identical on all ranks after the step: {same}")
flag = torch.tensor([1.0 if (ok and same) else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)

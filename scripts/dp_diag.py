"""2-rank diagnostic of the data-parallel path with progress prints and a stack dump on a stall (gpurun --gpus 2)."""
import faulthandler, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
faulthandler.dump_traceback_later(45, exit=True)
def say(*a):
    print(f"[r{rank} {time.time() % 1000:.1f}]", *a, flush=True)
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
say("init ok")
t = torch.ones(4, device=dev); dist.all_reduce(t); torch.cuda.synchronize(); say("allreduce ok", t[0].item())
import mmvae_b200 as M
from mmvae_b200 import parallel as PAR, data as D
torch.manual_seed(0)
m = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
          sigma_decoder=0.1, input_image_size=64, precision=os.environ.get("PREC", "bf16")).to(dev).train()
PAR.data_parallel(m); say("data_parallel ok")
x = D.prepare_input(D.synthetic_labels(16, 64, seed=5 + rank).to(dev))
ns = types.SimpleNamespace(data_ratio_of_labels=None)
m.mmd_diagnostic = os.environ.get("MMD", "1") == "1"
for it in range(int(os.environ.get("EAGER", "2"))):
    mu, lv, enc, rec = m(x); say("fwd enqueued")
    loss, *_ = m.loss(x, mu, lv, enc, rec, dev, ns); say("loss ok")
    m.zero_grad(set_to_none=True)
    loss.backward(); say("bwd enqueued")
    torch.cuda.synchronize(); say("eager step", it, float(loss))
m.defer_metrics = True
g = M.GraphedTrainStep(m, 16, args=ns, warmup=2); say("graph captured")
for it in range(3):
    g(x)
torch.cuda.synchronize(); say("graph replays ok", float(g.loss))
opt = M.FusedAdam(m)
g2 = M.GraphedTrainStep(m, 16, args=ns, warmup=1, optimizer=opt); say("graph2 captured")
for it in range(3):
    g2(x)
torch.cuda.synchronize(); say("graph2 replays ok", float(g2.loss))
g.graph = g2.graph = None
del g, g2
import gc; gc.collect()
dist.barrier(); dist.destroy_process_group(); say("done")

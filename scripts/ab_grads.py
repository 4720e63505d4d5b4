"""One seeded train step at N frames through the public module API; dumps loss + the gradient arena so that two builds /
environment toggles can be compared (python scripts/ab_grads.py out.pt; python scripts/ab_grads.py a.pt b.pt = compare)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
if len(sys.argv) == 3:
    a, b = torch.load(sys.argv[1]), torch.load(sys.argv[2])
    ga, gb = a["grads"].double(), b["grads"].double()
    print("loss", a["loss"], b["loss"], "rel grad diff", ((ga - gb).norm() / ga.norm()).item(), "max abs", (ga - gb).abs().max().item())
    sys.exit(0)
import mmvae_b200 as M
from mmvae_b200 import data as D
n = int(os.environ.get("N", "256"))
torch.manual_seed(0)
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64, precision="bf16").cuda().train()
x = D.prepare_input(D.synthetic_labels(n, 64).cuda())
eps = torch.randn(n, 64, 1, 1, generator=torch.Generator().manual_seed(4321)).cuda()
largs = types.SimpleNamespace(data_ratio_of_labels=None)
mu, lv, enc, rec = model(x, eps=eps)
loss, *_ = model.loss(x, mu, lv, enc, rec, x.device, largs)
model.zero_grad(set_to_none=True)
loss.backward()
torch.cuda.synchronize()
grads = torch.cat([p.grad.flatten() for p in model.parameters()]).cpu()
torch.save({"loss": float(loss), "grads": grads}, sys.argv[1])
print("loss", float(loss), "grad norm", grads.norm().item())

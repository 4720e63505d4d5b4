"""Run W warm-up steps and S profiled steps of the bench workload (for ncu launch lists / captures)."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmvae_b200 as M
from mmvae_b200 import data as D

n = int(os.environ.get("N", "256")); W = int(os.environ.get("W", "2")); S = int(os.environ.get("S", "1"))
prec = os.environ.get("PREC", "bf16")
torch.manual_seed(0)
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False,
              only_pixelcnn=False, sigma_decoder=0.1, input_image_size=64, precision=prec).cuda().train()
model.defer_metrics = True
x = D.prepare_input(D.synthetic_labels(n, 64).cuda())
largs = types.SimpleNamespace(data_ratio_of_labels=None)
for i in range(W + S):
    if i == W:
        torch.cuda.synchronize()
        print("launches before profiled step:", M._lib.lib.mmvae_launch_count())
    mu, lv, enc, rec = model(x)
    loss, *_ = model.loss(x, mu, lv, enc, rec, x.device, largs)
    model.zero_grad(set_to_none=True)
    loss.backward()
torch.cuda.synchronize()
print("total launches:", M._lib.lib.mmvae_launch_count(), "loss", float(loss.detach()))

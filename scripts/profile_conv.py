"""One launch of the production kernel of chosen (conv, direction) pairs through mmvae_bench_conv, for `ncu --set full`.
    ncu --set full --clock-control none --import-source on -k regex:'gconv_tc|slab_' -s <skip> -c <n> -o out python scripts/profile_conv.py
The step runs first (so the workspace holds real tensors); MARK kernels (a tiny torch fill of 12345 elements) bracket the
measured launches so that they are easy to find in a launch list."""
import ctypes, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmvae_b200 as M
from mmvae_b200 import data as D

PAIRS = [("encoder.layer4.0.conv2", 0), ("encoder.layer4.0.conv2", 1), ("decoder.uplayer5.0.conv2", 0),
         ("decoder.uplayer5.0.conv2", 1), ("decoder.uplayer5.0.conv2", 2), ("encoder.layer1.0.conv2", 0)]
n = int(os.environ.get("N", "256"))
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64, precision="bf16").cuda().train()
model.defer_metrics = True
model.mmd_diagnostic = False
x = D.prepare_input(D.synthetic_labels(n, 64).cuda())
largs = types.SimpleNamespace(data_ratio_of_labels=None)
for _ in range(2):
    mu, lv, enc, rec = model(x)
    loss, *_ = model.loss(x, mu, lv, enc, rec, x.device, largs)
    model.zero_grad(set_to_none=True)
    loss.backward()
torch.cuda.synchronize()
desc, ws, _ = model._workspace(n, True)
names = [c[0] for c in M._lib.conv_table(desc)]
scratch = torch.zeros(model._n_params, dtype=torch.float32, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ab, af = ctypes.c_int64(), ctypes.c_int64()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
mark = torch.empty(12345, dtype=torch.float32, device="cuda")
for name, d in PAIRS:
    flush.zero_()
    mark.fill_(1.0)
    rc = M._lib.lib.mmvae_bench_conv(ctypes.byref(desc), names.index(name), d, ctypes.c_void_p(model._arena.data_ptr()),
                                     ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(scratch.data_ptr()),
                                     ctypes.byref(ab), ctypes.byref(af), stream)
    torch.cuda.synchronize()
    print(name, d, "rc", rc, "algo MB", ab.value / 1e6, "MFLOP", af.value / 1e6)

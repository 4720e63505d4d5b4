"""Debug: weight gradient of one transposed conv through mmvae_bench_conv (dir 2) vs torch autograd on the same workspace
tensors; prints the relative error per (ky, kx) and per channel half."""
import ctypes, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmvae_b200 as M
from mmvae_b200 import data as D

name = sys.argv[1] if len(sys.argv) > 1 else "decoder.uplayer5.0.conv2"
n = int(os.environ.get("N", "8"))
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64, precision="bf16").cuda().train()
model.defer_metrics = True
x = D.prepare_input(D.synthetic_labels(n, 64).cuda())
largs = types.SimpleNamespace(data_ratio_of_labels=None)
mu, lv, enc, rec = model(x)
loss, *_ = model.loss(x, mu, lv, enc, rec, x.device, largs)
loss.backward()
torch.cuda.synchronize()
desc, ws, _ = model._workspace(n, True)
table = M._lib.conv_table(desc)
names = [c[0] for c in table]
ci = names.index(name)
def wt(nm):
    off, dims = M._lib.workspace_tensor(desc, nm)
    numel = dims[0] * dims[1] * dims[2] * dims[3]
    return ws[off:off + 2 * numel].view(torch.bfloat16).view(dims).permute(0, 3, 1, 2).float()
blk = name.rsplit(".", 1)[0]
inp = wt(blk + ".relu1") if name.endswith("conv2") else None
if inp is None:
    prev = {"decoder.uplayer5.0": "decoder.uplayer4.0", "decoder.uplayer4.0": "decoder.uplayer3.0"}[blk]
    inp = wt(prev)
dy = wt(name + ".grad")
w = torch.zeros(inp.shape[1], dy.shape[1], 4, 4, device="cuda", requires_grad=True)
y = torch.nn.functional.conv_transpose2d(inp, w, stride=2, padding=1)
(y * dy).sum().backward()
ref = w.grad
scratch = torch.zeros(model._n_params, dtype=torch.float32, device="cuda")
ab, af = ctypes.c_int64(), ctypes.c_int64()
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rc = M._lib.lib.mmvae_bench_conv(ctypes.byref(desc), ci, 2, ctypes.c_void_p(model._arena.data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                                 ws.numel(), ctypes.c_void_p(scratch.data_ptr()), ctypes.byref(ab), ctypes.byref(af), stream)
torch.cuda.synchronize()
assert rc == 0, M._lib.lib.mmvae_last_error()
ptab = {p[0]: p for p in M._lib.param_table(desc)}
off = ptab[name + ".weight"][1]
got = scratch[off:off + ref.numel()].view_as(ref)
print("total rel", ((got - ref).norm() / ref.norm()).item())
for ky in range(4):
    print("ky", ky, " ".join(f"{((got[:, :, ky, kx] - ref[:, :, ky, kx]).norm() / ref[:, :, ky, kx].norm()).item():.3f}" for kx in range(4)),
          "  | got/ref norms", " ".join(f"{(got[:, :, ky, kx].norm() / ref[:, :, ky, kx].norm()).item():.2f}" for kx in range(4)))
h = ref.shape[0] // 2
print("ci halves", ((got[:h] - ref[:h]).norm() / ref[:h].norm()).item(), ((got[h:] - ref[h:]).norm() / ref[h:].norm()).item())
print("co halves", ((got[:, :8] - ref[:, :8]).norm() / ref[:, :8].norm()).item(), ((got[:, 8:] - ref[:, 8:]).norm() / ref[:, 8:].norm()).item())

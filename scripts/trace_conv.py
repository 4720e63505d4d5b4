"""Timeline of ONE launch of the tcgen05 conv kernel per layer (debugging): %globaltimer stamps of the pipeline
milestones of every CTA, printed relative to the earliest CTA's entry.  Usage (on a B200):
    python scripts/trace_conv.py [layer-name ...]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmvae_b200 as M
from mmvae_b200 import data as D
import types

SLAB_SLOTS = ["entry", "prologue", "pdl_wait", "fill0_issued", "fill0_done", "fill1_done", "m_full0", "m_tile0", "m_slab0", "m_all", "e_tfull0", "e_tile0", "e_all", "dealloc", "bn_fin"]
SLOTS = ["entry", "prologue", "pdl_wait", "-", "tma_all", "-", "mma_tile0", "mma_all", "tfull0", "epi0", "epi_all",
         "stats", "dealloc", "bn_fin", "-", "-", "p_decoded"]
n = int(os.environ.get("N", "256"))
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64, pixelcnn=False, only_pixelcnn=False,
              sigma_decoder=0.1, input_image_size=64, precision="bf16").cuda().train()
model.defer_metrics = True
x = D.prepare_input(D.synthetic_labels(n, 64).cuda())
largs = types.SimpleNamespace(data_ratio_of_labels=None)
for _ in range(2):
    mu, lv, enc, rec = model(x)
    loss, *_ = model.loss(x, mu, lv, enc, rec, x.device, largs)
    model.zero_grad(set_to_none=True)
    loss.backward()
torch.cuda.synchronize()
desc, ws, _info = model._workspace(n, True)
table = M._lib.conv_table(desc)
names = [c[0] for c in table]
want = sys.argv[1:] or ["decoder.uplayer1.0.conv1", "encoder.layer4.0.conv2", "encoder.layer1.0.conv2", "decoder.uplayer5.0.conv2"]
buf = torch.zeros(444 * 32, dtype=torch.int64, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ab, af = ctypes.c_int64(), ctypes.c_int64()
for name in want:
    ci = names.index(name)
    for d in (0, 1):
        for rep in range(3):
            buf.zero_()
            torch.cuda.synchronize()
            M._lib.lib.mmvae_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
            rc = M._lib.lib.mmvae_bench_conv(ctypes.byref(desc), ci, d, ctypes.c_void_p(model._arena.data_ptr()),
                                             ctypes.c_void_p(ws.data_ptr()), ws.numel(), None, ctypes.byref(ab), ctypes.byref(af), stream)
            M._lib.lib.mmvae_debug_set_trace(ctypes.c_void_p(0))
            torch.cuda.synchronize()
            if rc != 0:
                break
        if rc != 0:
            print(name, "dir", d, "skipped:", M._lib.lib.mmvae_last_error().decode())
            continue
        t = buf.view(444, 32).cpu()
        used = (t[:, 0] > 0).nonzero().flatten()
        t0 = int(t[used, 0].min())
        print(f"== {name} dir {d}: {len(used)} CTAs, algo {ab.value / 1e6:.2f} MB {af.value / 1e6:.0f} MFLOP; "
              f"kernel span {(int(t[used].max()) - t0) / 1e3:.2f} us")
        for label, sel in (("first CTA", int(used[0])), ("last-ending CTA", int(used[t[used].max(dim=1).values.argmax()]))):
            row = t[sel]
            slots = SLAB_SLOTS if (os.environ.get("SLAB") == "1") else SLOTS
            print(f"   {label:16s} (cta {sel}): " + "  ".join(f"{slots[s]}={(int(row[s]) - t0) / 1e3:.2f}" for s in range(len(slots)) if row[s] > 0))

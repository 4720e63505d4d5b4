"""GPU timeline of one CUDA-graph replay of the train step (CUPTI through torch.profiler): per kernel start, duration,
stream and the gap to the previous kernel on the same stream.  A diagnostic, never a bench number.
    python scripts/timeline.py [out.csv]"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import mmvae_b200 as M
from mmvae_b200 import data as D

n = int(os.environ.get("N", "256"))
width = int(os.environ.get("WIDTH", "1"))          # WIDTH=2: the widened model (BASELINE configs[3], z = 256)
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "timeline.csv")
torch.manual_seed(0)
model = M.VAE(1, 32, decoder_out_channels=1, pixelcnn_out_channels=0, z_dimension=64 if width == 1 else 256, pixelcnn=False,
              only_pixelcnn=False, sigma_decoder=0.1, input_image_size=64, precision="bf16", width=width).cuda().train()
model.defer_metrics = True
model.mmd_diagnostic = os.environ.get("MMD", "1") == "1"     # A/B: the MMD diagnostic's side branch
largs = types.SimpleNamespace(data_ratio_of_labels=None)
g = M.GraphedTrainStep(model, n, args=largs, warmup=2)
g.x.copy_(D.prepare_input(D.synthetic_labels(n, 64).cuda()))
for _ in range(5):
    g(None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        g(None)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays at pack_weights_kernel
starts = [i for i, e in enumerate(evs) if "pack_weights" in e.name]
if len(starts) >= 2:
    evs = evs[starts[-2]:starts[-1]]
t0 = evs[0].time_range.start
last_end = {}
rows = []
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    stream = getattr(e, "stream", None)
    if stream is None:
        stream = e.device_index
    gap = s - last_end.get(stream, s)
    last_end[stream] = s + d
    name = e.name.replace("mmvae::(anonymous namespace)::", "").replace("void ", "")
    rows.append((s, d, gap, stream, name[:70]))
with open(out, "w") as f:
    f.write("start_us,dur_us,gap_us,stream,name\n")
    for r in rows:
        f.write(f"{r[0]:.2f},{r[1]:.2f},{r[2]:.2f},{r[3]},{r[4]}\n")
tot = rows[-1][0] + rows[-1][1]
print(f"replay span {tot:.1f} us, {len(rows)} activities")
by = {}
for s, d, gap, st, name in rows:
    k = (st, name.split("(")[0][:50])
    a = by.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += d; a[2] += max(gap, 0.0)
for (st, name), (c, d, gp) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"stream {st} {name:52s} x{c:3d} dur {d:8.1f} us  gaps-before {gp:7.1f} us")

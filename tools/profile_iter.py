"""Run a few native iterations at one size; the profiled window (cudaProfilerStart/Stop) is the last K iterations.
Usage under ncu:  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
                      --log-file gpurun_out/launches.csv python tools/profile_iter.py 2048 1"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from oracle import st_oracle as O  # noqa: E402  (weights/images fixture only)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device('cuda:0')
wts = O.make_vgg_weights(1234)
st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
m = st.model
m.ensure_workspace([(size, size)])
cimg = O.to_tensor(O.synth_image(1, 16, size, size)).to(dev)
simg = O.to_tensor(O.synth_image(2, 32, size, size)).to(dev)
ct = m.content_features(cimg)
means, srms = m.style_stats(simg)
m.set_targets(size, size, ct, 0.015, means, srms, st.style_weights, 2.0)
st.image = cimg.clone()
st.average = stb.style_transfer.EMA(st.image, 0.99)
ea, eas = torch.zeros_like(st.image), torch.zeros_like(st.image)
for i in range(3):
    st._iterate(ea, eas, i + 1, 0.02, 0.99, True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(k):
    st._iterate(ea, eas, i + 4, 0.02, 0.99, True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('loss', float(st._loss_host[0]))

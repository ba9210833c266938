"""Timeline of the W2 chain kernel's rounds as seen by CTA 0 (largest C = 512 tile of every round).
  STB_W2_TRACE=1 python tools/w2_trace.py [size]"""
import ctypes
import os
import sys
from pathlib import Path

import numpy as np
import torch

os.environ['STB_W2_TRACE'] = '1'
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from style_transfer_b200 import _lib  # noqa: E402
from oracle import st_oracle as O  # noqa: E402  (weights/images fixture only)

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
for i in range(8):
    st._iterate(ea, eas, i + 1, 0.02, 0.99, True)
torch.cuda.synchronize()
buf = np.zeros(8 * 128, dtype=np.uint64)
n = ctypes.c_int()
_lib.check(m.lib.stb_debug_w2_trace(m.ctx, buf.ctypes.data_as(ctypes.c_void_p), buf.size, ctypes.byref(n)))
t = buf.reshape(128, 8).astype(np.int64)
names = ['start->data', 'data->acc', 'acc->staged', 'staged->stored', 'stored->arrive', 'arrive->passed']
rows = []
for r in range(n.value):
    if t[r, 0] == 0:
        continue
    d = [t[r, k + 1] - t[r, k] if t[r, k + 1] and t[r, k] else -1 for k in range(6)]
    nxt = t[r + 1, 0] - t[r, 6] if r + 1 < n.value and t[r + 1, 0] and t[r, 6] else -1
    rows.append((r, d, nxt, t[r, 6] - t[r, 0] if t[r, 6] else -1))
print('round  ' + '  '.join(f'{x:>15s}' for x in names) + '   passed->next  total(ns)')
for r, d, nxt, tot in rows:
    print(f'{r:5d}  ' + '  '.join(f'{x:15d}' for x in d) + f'  {nxt:12d}  {tot:9d}')
arr = np.array([d for _, d, _, _ in rows if min(d) >= 0])
if len(arr):
    print('median ', '  '.join(f'{int(x):15d}' for x in np.median(arr, axis=0)), '  total', int(np.median([tt for *_, tt in rows if tt > 0])))

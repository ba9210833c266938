"""BASELINE.json configs[4] under torchrun (one rank per GPU): 2896 x 2172 content, the default coarse-to-fine pyramid
(128 ... 2896, ten scales), two style images with style_weights=[3, 1] (`-sw 3 1`), tiled over all ranks; the same job
is then run untiled on every rank's own GPU and compared (loss trace per scale, final image).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
      tools/config5_check.py [initial_iterations iterations]      -> gpurun_out/config5_N<world>.json (rank 0)

Iteration counts default to 20 + 9 x 10 (SURVEY.md section 8d allows the CI-size cut of 1000 + 9 x 500).
"""
import contextlib
import io
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import style_transfer_b200 as stb  # noqa: E402
from oracle import st_oracle as O  # noqa: E402  (fixture generator only)

INIT_ITS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ITS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
wts = O.make_vgg_weights(1234)
content = O.synth_image(1, 16, 2896, 2172)
styles = [O.synth_image(2, 32, 2048, 1536), O.synth_image(3, 32, 1500, 2000)]
kw = dict(style_weights=[3, 1], end_scale=2896, min_scale=128, initial_iterations=INIT_ITS, iterations=ITS)


def run(distributed):
    st = stb.StyleTransfer(devices=[f'cuda:{local}'], pooling='max', vgg_weights=wts, distributed=distributed)
    trace = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        img = st.stylize(content, styles, callback=lambda it: trace.append((it.w, it.h, it.i, it.loss, it.time)), **kw)
    torch.cuda.synchronize()
    return trace, np.asarray(img, dtype=np.float32), time.perf_counter() - t0, (st._comm_mode, st._tile_mode, st._halo_now), st.model.graph_status()


tr_t, img_t, sec_t, mode, gstat = run(None)
dist.barrier()
tr_s, img_s, sec_s, _, _ = run(False)
lt, ls = np.array([t[3] for t in tr_t]), np.array([t[3] for t in tr_s])
rel = np.abs(lt - ls) / np.abs(ls)
scales = []
for (w, h) in dict.fromkeys((t[0], t[1]) for t in tr_t):
    idx = [k for k, t in enumerate(tr_t) if (t[0], t[1]) == (w, h)]
    ts = [tr_t[k][4] for k in idx]
    its = (len(ts) - 1) / (ts[-1] - ts[0]) if len(ts) > 1 and ts[-1] > ts[0] else None
    from style_transfer_b200 import distributed as D
    scales.append(dict(w=w, h=h, tiled=D.make_band(h, 0, world) is not None, first_loss=float(lt[idx[0]]),
                       last_loss=float(lt[idx[-1]]), max_rel_vs_untiled=float(rel[idx].max()), it_per_s=its))
d = np.abs(img_t - img_s)
ok = bool(rel.max() < 3e-3 and d.mean() < 1.5)
out = dict(config='BASELINE.json configs[4]: 2896x2172 content, default pyramid, 2 styles with -sw 3 1', n_gpus=world,
           iterations=f'{INIT_ITS} + 9 x {ITS}', comm_mode=mode, graph_status=gstat, seconds_tiled=sec_t,
           seconds_untiled_one_gpu=sec_s, scales=scales, max_rel_loss_diff_vs_untiled=float(rel.max()),
           final_image_mean_abs_diff_255=float(d.mean()), final_image_max_abs_diff_255=float(d.max()), ok=ok)
if rank == 0:
    Path(ROOT / 'gpurun_out').mkdir(exist_ok=True)
    (ROOT / 'gpurun_out' / f'config5_N{world}.json').write_text(json.dumps(out, indent=1))
    print(json.dumps(out))
flag = torch.tensor([int(ok)], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)

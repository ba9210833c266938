#!/usr/bin/env python
"""Benchmark of the stylize() hot path: iterations/sec at end_scale=2048 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--size 2048] [--impl native|reference]

A "step" is one stylize iteration (VGG-19 forward, W2/content/TV losses, backward to the image, Adam+clamp+EMA)
on a synthetic 2048x2048 content+style pair with the seeded synthetic VGG-19 (no network for images/checkpoints).

native arm   value   = it/s of K stb_iterate calls, state resident in HBM, CUDA events on the launching stream,
                       barrier + synchronize on both sides, max over ranks.
             e2e     = it/s of the public API: StyleTransfer.stylize(PIL content, [PIL style], single scale, K
                       iterations, callback reading the loss every iteration) -- host PIL inputs, H2D of the
                       resized images, per-iteration D2H of the loss, final D2H of the result, all inside the timed
                       region (wall clock around the call, since it spans host work).
             roofline= conv tensor-pipe: algorithmic conv FLOPs per iteration / summed duration of the tcgen05 conv
                       launches (CUDA events around every launch, second instrumented pass), vs the measured
                       sustained bf16 peak of MEASURED_PEAKS.json.
             cpu_baseline = the oracle port (same torch-CPU primitives the reference's CPU path uses) timed on the
                       host cores on a bounded sample, extrapolated with an affine cost model in pixels.
reference arm (--impl reference): the CPU baseline alone, printed in the same JSON shape.
N > 1: ONE 2048x2048 job tiled spatially over the N GPUs (strong scaling): horizontal bands with 80-row halo aprons,
one NCCL all-reduce of the 2.4 MB statistics block, a seam exchange of the image gradient and a halo refresh of the
image per iteration (style-transfer-pytorch_b200/distributed.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CONV_FLOP_PER_PIXEL = 1444608.0  # fwd + dgrad, 2 flop/MAC
CONV0_FLOP_PER_PIXEL = 4 * 27 * 64.0  # conv0 fwd + dgrad run in their own kernels (conv0_tc.cu), not in the timed class
PROF_CLASSES = ['conv0_fwd_tv', 'conv_fwd', 'pool_fwd', 'gram', 'sse', 'w2', 'conv_bwd', 'pool_bwd',
                'conv0_bwd_adam', 'finalize']


def measured_peaks():
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'],
                    bf16_tflops_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']), source='measured')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax = float(parts[2])
            except ValueError:
                continue
            for nm, val in zip(names, parts[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        sm.sort()
        hi = [v for v in sm if smax and v > 0.3 * smax] or sm
        med = hi[len(hi) // 2] if hi else None
        return dict(sm_mhz=med, sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_port_iteration_time(size, iters, warm):
    """Seconds per iteration of the oracle port at size x size on the host cores (all threads torch uses)."""
    import torch
    from oracle import st_oracle as O
    wts = O.make_vgg_weights(1234)
    content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
    tg, _ = O.make_targets(content, [style], [1.0], size, wts, 'max', 0.015, 2.0)
    st = O.IterState.fresh(O.to_tensor(content))
    times = []
    for i in range(warm + iters):
        t0 = time.perf_counter()
        O.iterate(st, wts, tg, 'max')
        times.append(time.perf_counter() - t0)
    times = sorted(times[warm:])
    return times[len(times) // 2], torch.get_num_threads()


def cpu_reference_iteration_time(size, iters):
    """Seconds per iteration of the UNMODIFIED reference's stylize() on the host cores (baseline/_ref install)."""
    import torch
    from oracle import reference_harness as RH
    from oracle import st_oracle as O
    wts = O.make_vgg_weights(1234)
    content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
    t, _ = RH.time_reference(size, iters, wts, content, style, devices=('cpu',))
    return t, torch.get_num_threads()


def cpu_baseline(size, budget_s=25.0):
    """Affine model t(px) = a + b*px from two bounded samples (the W2/sqrtm part, ~65 GFLOP, does not scale with
    pixels; the conv part does), evaluated at size^2.  Times the unmodified reference (kind "reference") when its
    install travelled with the repo (baseline/_ref), else the oracle port of the same loop body (kind "port")."""
    from oracle import reference_harness as RH
    if RH.reference_available():
        kind, what = 'reference', 'unmodified reference StyleTransfer.stylize(devices=[cpu]), median of STIterate.time diffs'
        timer, its = cpu_reference_iteration_time, 3
    else:
        kind, what = 'port', 'oracle port (torch-CPU explicit schedule), median'
        timer, its = (lambda size_, n: cpu_port_iteration_time(size_, n, 1)), 2
    # two bounded samples; the larger one as close to the target size as the time budget allows (cache effects make
    # the small sizes optimistic)
    small = 256
    t256, threads = timer(small, its)
    big = 512
    for cand in (1024, 768):
        if t256 * (cand * cand) / (small * small) * (its + 3) < budget_s:
            big = cand
            break
    tbig, _ = timer(big, its)
    b = (tbig - t256) / (big * big - small * small)
    a = max(t256 - b * small * small, 0.0)
    t_full = a + b * size * size
    return dict(value=1.0 / t_full, unit='it/s', cores=threads, kind=kind,
                sample=f'{what}: {small}^2 ({t256:.3f} s/it) and {big}^2 ({tbig:.3f} s/it), affine-in-pixels '
                       f'extrapolation to {size}^2 ({t_full:.2f} s/it)')


# ------------------------------------------------------------------------------------------------ arms
def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = cpu_baseline(args.size, budget_s=60.0)
    line = dict(metric='stylize iterations/sec at end_scale=2048', value=cb['value'], unit='it/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup, ms_per_step=1000.0 / cb['value'], higher_is_better=True,
                scaling='strong', vs_baseline=None, dtype='f32', data='synthetic', impl='reference',
                config=dict(workload=f'{args.size}x{args.size} single scale, pooling=max, content+1 style, '
                                     'CPU reference path (' + cb['kind'] + ')'),
                cpu_baseline=cb, e2e=dict(value=cb['value'], unit='it/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import ctypes
    import style_transfer_b200 as stb
    from style_transfer_b200 import _lib
    from oracle import st_oracle as O  # fixture generator (weights / synthetic images) + cpu_baseline leg only

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    size = args.size
    wts = O.make_vgg_weights(1234)
    content, style = O.synth_image(1, 16, size, size), O.synth_image(2, 32, size, size)
    from style_transfer_b200 import distributed as D
    st = stb.StyleTransfer(devices=[str(dev)], pooling='max', vgg_weights=wts)
    m = st.model
    band = D.make_band(size, rank, world) if world > 1 else None
    h_loc = band.h_local if band is not None else size
    m.ensure_workspace([(h_loc, size), (size, size)])
    cimg = O.to_tensor(content).to(dev)
    simg = O.to_tensor(style).to(dev)
    means, srms = m.style_stats(simg)
    if band is not None:
        cimg = D.local_slice(cimg, band)
        m.set_band(True, size, band.own0, band.own_rows)
    ct = m.content_features(cimg)
    m.set_targets(h_loc, size, ct, 0.015, means, srms, st.style_weights, 2.0)
    st.image = cimg.clone()
    st.average = stb.style_transfer.EMA(st.image, 0.99)
    ea, eas = torch.zeros_like(st.image), torch.zeros_like(st.image)
    step = 0
    if band is not None:
        stats, grad = m.stats_view(h_loc, size), torch.empty_like(st.image)

    def one_iteration():
        nonlocal step
        step += 1
        if band is not None:
            st._iterate_banded(band, stats, grad, ea, eas, step, 0.02, 0.99)
        else:
            st._iterate(ea, eas, step, 0.02, 0.99, True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    side = torch.cuda.Stream(device=dev)  # non-legacy stream: lets the library replay iterations as a CUDA graph
    side.wait_stream(torch.cuda.current_stream(dev))
    torch.cuda.set_stream(side)
    for _ in range(max(args.warmup, 3)):
        one_iteration()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_iteration()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(st._loss_host[0])
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    tiled = band is not None
    # tiled: all ranks advanced ONE job by `steps` iterations; otherwise (N == 1) the single job
    value = args.steps / (ms_total / 1000.0)

    # ---- instrumented pass: per-kernel-class device time (events around every launch)
    prof = None
    if rank == 0:
        _lib.check(m.lib.stb_profile_enable(m.ctx, 1))
        n_prof = min(args.steps, 10)
        for _ in range(n_prof):
            if band is None:
                one_iteration()
            else:  # instrumented compute only (no collectives inside the event spans' critical path on rank 0)
                step += 1
                m.iterate_fwd(st.image)
                m.iterate_bwd(st.image, grad, st._loss_host)
        torch.cuda.synchronize()
        ms = (ctypes.c_float * 10)()
        cnt = (ctypes.c_int * 10)()
        _lib.check(m.lib.stb_profile_read(m.ctx, ms, cnt, 10))
        _lib.check(m.lib.stb_profile_enable(m.ctx, 0))
        prof = {name: dict(ms_per_iter=ms[i] / n_prof, launches_per_iter=cnt[i] / n_prof)
                for i, name in enumerate(PROF_CLASSES)}
    barrier()

    # ---- e2e through the public API (host PIL in, loss read back every iteration, result image on the host)
    e2e = None
    losses = []
    st2 = stb.StyleTransfer(devices=[str(dev)], pooling='max', vgg_weights=wts)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):  # stylize() prints progress like the reference does
        st2.stylize(content, [style], min_scale=size, end_scale=size, initial_iterations=2, callback=lambda it: None)
    barrier()
    # BASELINE.json's config runs 500 iterations at this scale; fewer would mostly time the per-call setup
    e2e_its = max(args.steps, 500)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = st2.stylize(content, [style], min_scale=size, end_scale=size, initial_iterations=e2e_its,
                          callback=lambda it: losses.append(it.loss))
    _ = out.size
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    h2d = 2 * 3 * size * size                # content + style uint8 pixels (converted to fp32 on the device)
    d2h = 32 * e2e_its + 3 * size * size     # loss terms every step + final uint8 image
    e2e = dict(value=e2e_its / t_e2e, unit='it/s', h2d_bytes_per_step=h2d / e2e_its,
               d2h_bytes_per_step=d2h / e2e_its, iterations=e2e_its,
               note='StyleTransfer.stylize(PIL inputs, single scale, per-iteration loss callback) wall clock over '
                    f'{e2e_its} iterations (BASELINE.json config: 500 at this scale), incl. PIL resize, target '
                    'extraction, H2D of the inputs, per-iteration loss D2H, final image D2H')

    if rank == 0:
        peaks = measured_peaks()
        # the twelve 3x3 convs of pixel_gemm_kernel on this rank's rows (band + aprons when tiled); the tap-gradient
        # GEMMs folded into the same launches (0.15 TFLOP at 2048^2) are NOT credited
        conv_flops = (CONV_FLOP_PER_PIXEL - CONV0_FLOP_PER_PIXEL) * h_loc * size
        conv_ms = prof['conv_fwd']['ms_per_iter'] + prof['conv_bwd']['ms_per_iter']
        achieved = conv_flops / (conv_ms / 1000.0) / 1e12
        peak = peaks['bf16_tflops_sustained']
        # kernels per iteration: counted from the ncu launch list of one eager iteration
        # (profiles/r1_launches_2048_v11.csv: 24 conv + 52 W2 rounds + 5 W2 helpers + 5 Gram + 5 reduce + conv0 fwd/bwd
        # + border + TV + 4 pool bwd + SSE + 3 scalar/finalize kernels = 104); the tiled path adds adam_rows
        launches = 104 + (1 if tiled else 0)
        traffic = None
        tpath = ROOT / 'profiles' / 'r1_conv_traffic.json'
        if tpath.exists() and size == 2048 and world == 1:
            traffic = json.loads(tpath.read_text())['dram_total_bytes']  # dram read+write of the conv launches, ncu
        roofline = dict(bound='tensor', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                        traffic=traffic, peak_source=f"{peaks['source']} (bf16_tflops_sustained)",
                        kernel='pixel_gemm_kernel (tcgen05 conv fwd+dgrad incl. tap-gradient GEMMs, 25 launches/iter)',
                        note='algorithmic FLOPs/iter of the twelve 3x3 convs (1437696/pixel) / summed CUDA-event duration of the conv '
                             'launches in an instrumented pass; traffic = DRAM bytes (read+write) of the same launches per '
                             'iteration from one ncu --set full capture (profiles/r1_ncu_conv_v11.csv; algorithmic ~8.5 GB)',
                        step_fraction=conv_ms / sum(v['ms_per_iter'] for v in prof.values()),
                        classes_ms_per_iter={k: round(v['ms_per_iter'], 4) for k, v in prof.items()},
                        whole_step_tensor_frac=CONV_FLOP_PER_PIXEL * size * size * (1000.0 / ms_per_step) / 1e12 /
                        (peak * world))
        cb = None
        if not args.no_cpu_baseline and world == 1:  # reported baseline: rank 0, N = 1 only
            cb = cpu_baseline(size)
        line = dict(metric='stylize iterations/sec at end_scale=2048', value=value, unit='it/s', n_gpus=world,
                    steps=args.steps, warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True,
                    scaling='strong', vs_baseline=None, dtype='bf16', data='synthetic', impl='native',
                    config=dict(workload=f'{size}x{size} single scale (BASELINE.json configs[2]), pooling=max, '
                                         'content+1 style, bf16 operands / fp32 accumulate, fp32 sqrtm+Adam',
                                parallelism='single GPU' if world == 1 else
                                f'{world}-way spatial tiling: bands of {band.own_rows}+{band.top_apron + band.bottom_apron} '
                                'halo rows (rank 0), 1 stats all-reduce + grad seam exchange + halo refresh per iteration',
                                l2='working set per iteration (>= 2.4 GB of activations) far exceeds the 126 MB L2',
                                final_loss=final_loss),
                    clocks=clocks, e2e=e2e, gpu_launches=int(round(launches * args.steps)), roofline=roofline,
                    cpu_baseline=cb)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--size', type=int, default=2048)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_native(args, rank, local_rank, world)


if __name__ == '__main__':
    main()

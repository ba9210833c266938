"""Spatial tiling of the stylize() hot path across the GPUs of one box.

One process per GPU.  The image is cut into horizontal bands whose edges are multiples of 16 rows (four floor-mode
2x2 pools never straddle a seam).  Every rank works on its band plus a halo "apron" of APRON rows on each interior
side -- the receptive-field radius of relu5_1 rounded up to 80 = 5*16 -- so that every activation and every gradient
path that touches the band's own rows is exact without any per-layer exchange.  Per iteration there are three
exchanges, all tiny next to the compute:
  1. all-reduce(sum) of the stats block (5 Gram matrices + channel sums + content SSE + TV sum, 2.4 MB),
  2. seam reduce of the image gradient: the contributions a rank computed for its halo rows are added to the
     neighbour's own rows (<= 2 x APRON x W x 3 floats),
  3. halo refresh of the updated image rows.
Default ("peer" mode): all three run INSIDE the library, as kernels that read the peers' mailboxes over NVLink (CUDA
IPC peer memory, csrc/comm.cu), so an iteration is one CUDA graph per rank -- `stb_iterate_banded`.  Fallback ("nccl"
mode, `STB_COMM=nccl` or when the peer mapping cannot be set up): the host drives `stb_iterate_fwd` -> NCCL all-reduce
-> `stb_iterate_bwd` -> NCCL send/recv -> `stb_adam_update` (exchange_add_grad / exchange_halo below).
The reference's own multi-device mode (a 2-GPU layer split, ST:326-333) is superseded by this.

Collectives that are per SCALE, not per iteration (mailbox handle exchange, barriers, tiled style statistics, the
final gather of the bands) go through a small group interface: `TorchGroup` (torch.distributed, one process per GPU)
or `ThreadGroup` (the ranks are threads of one process sharing one GPU -- how the test-suite runs the very same tiled
path on a single-GPU box).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass

import torch
import torch.distributed as dist

APRON = 80
ALIGN = 16
MIN_BAND_ROWS = 96  # >= APRON so that a neighbour's apron never reaches past the adjacent band


@dataclass(frozen=True)
class Band:
    rank: int
    world: int
    H: int          # global height
    own_begin: int  # global rows [own_begin, own_end) are updated by this rank
    own_end: int
    loc_begin: int  # global rows [loc_begin, loc_end) are held locally (own rows + aprons)
    loc_end: int

    @property
    def own0(self):  # first own row in local coordinates (= height of the top apron)
        return self.own_begin - self.loc_begin

    @property
    def own_rows(self):
        return self.own_end - self.own_begin

    @property
    def h_local(self):
        return self.loc_end - self.loc_begin

    @property
    def top_apron(self):
        return self.own_begin - self.loc_begin

    @property
    def bottom_apron(self):
        return self.loc_end - self.own_end


def band_edges(H: int, world: int):
    """Band boundaries R_0=0 <= R_1 <= ... <= R_world=H, interior ones on multiples of ALIGN."""
    edges = [0]
    for r in range(1, world):
        edges.append(int(round(r * H / world / ALIGN)) * ALIGN)
    edges.append(H)
    return edges


def make_band(H: int, rank: int, world: int, apron: int = APRON):
    """Geometry of `rank`'s band, or None when the image is too small to tile (every band needs MIN_BAND_ROWS rows):
    the caller then runs the whole image on every rank (replicated, no communication)."""
    if world <= 1:
        return None
    edges = band_edges(H, world)
    if min(b - a for a, b in zip(edges[:-1], edges[1:])) < max(MIN_BAND_ROWS, apron):
        return None
    a, b = edges[rank], edges[rank + 1]
    return Band(rank, world, H, a, b, max(a - apron, 0), min(b + apron, H))


def all_bands(H: int, world: int):
    return [make_band(H, r, world) for r in range(world)]


# ------------------------------------------------------------------------------------------------ groups
class TorchGroup:
    """torch.distributed default group: one process per GPU (NCCL on the GPU box, gloo in the CPU tests)."""
    peer_kind = 'ipc'

    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def barrier(self):
        dist.barrier()

    def all_gather_object(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def all_reduce_sum(self, t):
        dist.all_reduce(t)
        return t

    def all_gather_tensor(self, t):
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t)
        return parts

    def broadcast(self, t, src=0):
        dist.broadcast(t, src)
        return t


class ThreadGroup:
    """The ranks are THREADS of one process (each with its own context and stream, usually on the same GPU): the
    per-scale collectives go through shared host state; tensors handed over are synchronised on the producer's
    stream first.  Build one `ThreadGroup.Shared(world)` and give `ThreadGroup(shared, rank)` to every thread."""
    peer_kind = 'local'

    class Shared:
        def __init__(self, world):
            self.world = world
            self.slots = [None] * world
            self.bar = threading.Barrier(world)

    def __init__(self, shared, rank):
        self.shared, self.rank, self.world = shared, rank, shared.world

    def barrier(self):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        self.shared.bar.wait()

    def all_gather_object(self, obj):
        self.shared.slots[self.rank] = obj
        self.barrier()
        out = list(self.shared.slots)
        self.barrier()
        return out

    def all_gather_tensor(self, t):
        return [p.to(t.device).clone() for p in self.all_gather_object(t.detach().clone())]

    def all_reduce_sum(self, t):
        parts = self.all_gather_tensor(t)
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        t.copy_(acc)
        return t

    def broadcast(self, t, src=0):
        t.copy_(self.all_gather_tensor(t)[src])
        return t


def default_group():
    """The torch.distributed group when it is initialised with more than one rank, else None (single GPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchGroup()
    return None


# ------------------------------------------------------------------------------------------------ nccl-mode exchanges
def _p2p(ops, group):
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def exchange_add_grad(grad: torch.Tensor, band: Band, group=None):
    """grad: [1,3,h_local,W] d loss/d(local image).  Adds the neighbours' halo contributions to this rank's own rows
    (in place) and ships this rank's halo contributions to the neighbours."""
    ops, recv = [], []
    up, down = band.rank - 1, band.rank + 1
    if band.top_apron > 0:  # my top apron = bottom own rows of `up`; `up`'s bottom apron = my first own rows
        send = grad[:, :, :band.top_apron].contiguous()
        n = min(APRON, band.own_rows)
        buf = torch.empty_like(grad[:, :, band.own0:band.own0 + n])
        ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, buf, up, group)]
        recv.append((buf, band.own0, band.own0 + n))
    if band.bottom_apron > 0:
        end = band.own0 + band.own_rows
        send = grad[:, :, end:end + band.bottom_apron].contiguous()
        n = min(APRON, band.own_rows)
        buf = torch.empty_like(grad[:, :, end - n:end])
        ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, buf, down, group)]
        recv.append((buf, end - n, end))
    _p2p(ops, group)
    for buf, lo, hi in recv:
        grad[:, :, lo:hi] += buf
    return grad


def exchange_halo(x: torch.Tensor, band: Band, group=None):
    """x: [1,C,h_local,W] whose own rows are current: refresh the apron rows from the neighbours' own rows (in place)."""
    ops, recv = [], []
    up, down = band.rank - 1, band.rank + 1
    if band.top_apron > 0:
        send = x[:, :, band.own0:band.own0 + APRON].contiguous()          # what `up` keeps as its bottom apron
        buf = torch.empty_like(x[:, :, :band.top_apron])
        ops += [dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, buf, up, group)]
        recv.append((buf, 0, band.top_apron))
    if band.bottom_apron > 0:
        end = band.own0 + band.own_rows
        send = x[:, :, end - APRON:end].contiguous()
        buf = torch.empty_like(x[:, :, end:end + band.bottom_apron])
        ops += [dist.P2POp(dist.isend, send, down, group), dist.P2POp(dist.irecv, buf, down, group)]
        recv.append((buf, end, end + band.bottom_apron))
    _p2p(ops, group)
    for buf, lo, hi in recv:
        x[:, :, lo:hi] = buf
    return x


# ------------------------------------------------------------------------------------------------ per-scale helpers
def gather_rows(x: torch.Tensor, band: Band, group=None):
    """All-gather the own rows of every rank into the full [1,C,H,W] tensor (identical on all ranks)."""
    group = group or TorchGroup()
    edges = band_edges(band.H, band.world)
    max_rows = max(b - a for a, b in zip(edges[:-1], edges[1:]))
    own = x[:, :, band.own0:band.own0 + band.own_rows]
    pad = torch.zeros(own.shape[0], own.shape[1], max_rows, own.shape[3], dtype=x.dtype, device=x.device)
    pad[:, :, :band.own_rows] = own
    parts = group.all_gather_tensor(pad)
    return torch.cat([p[:, :, :b - a] for p, a, b in zip(parts, edges[:-1], edges[1:])], dim=2)


def local_slice(full: torch.Tensor, band: Band):
    return full[:, :, band.loc_begin:band.loc_end].contiguous()


def tap_pixel_counts(h: int, w: int):
    """Pixels of the five style taps (relu1_1 ... relu5_1) of an h x w image: four floor-mode 2x2 pools."""
    out = []
    for _ in range(5):
        out.append(h * w)
        h, w = h // 2, w // 2
    return out

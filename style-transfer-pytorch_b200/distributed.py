"""Spatial tiling of the stylize() hot path across the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink).  The image is cut into horizontal bands whose edges are
multiples of 16 rows (four floor-mode 2x2 pools never straddle a seam).  Every rank works on its band plus a halo
"apron" of APRON rows on each interior side -- the receptive-field radius of relu5_1 (156 px -> 78, rounded up to
80 = 5*16) -- so that every activation and every gradient path that touches the band's own rows is exact without any
per-layer exchange.  Per iteration there are three collectives, all tiny next to the compute:
  1. all-reduce(sum) of the stats block (5 Gram matrices + channel sums + content SSE + TV sum, 2.4 MB),
  2. seam exchange of the image gradient: the contributions a rank computed for its halo rows are added to the
     neighbour's own rows (<= 2 x APRON x W x 3 floats),
  3. halo refresh of the updated image rows.
The reference's own multi-device mode (a 2-GPU layer split, ST:326-333) is superseded by this.
"""
from __future__ import annotations

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


def gather_rows(x: torch.Tensor, band: Band, group=None):
    """All-gather the own rows of every rank into the full [1,C,H,W] tensor (identical on all ranks)."""
    edges = band_edges(band.H, band.world)
    max_rows = max(b - a for a, b in zip(edges[:-1], edges[1:]))
    own = x[:, :, band.own0:band.own0 + band.own_rows]
    pad = torch.zeros(own.shape[0], own.shape[1], max_rows, own.shape[3], dtype=x.dtype, device=x.device)
    pad[:, :, :band.own_rows] = own
    parts = [torch.empty_like(pad) for _ in range(band.world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:, :, :b - a] for p, a, b in zip(parts, edges[:-1], edges[1:])], dim=2)


def local_slice(full: torch.Tensor, band: Band):
    return full[:, :, band.loc_begin:band.loc_end].contiguous()

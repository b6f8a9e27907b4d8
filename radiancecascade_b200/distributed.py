"""Multi-GPU partitioning of the GI path: one process per GPU.

Two ways to shard, both with a replicated scene (every bundled scene is < 4 MB on device):

* multi-view batch — each rank renders whole frames of its own views; no data-path
  collective at all (bench.py's default for N > 1, "weak" scaling);
* screen-space tiles — one frame is cut into horizontal strips; each rank renders its
  strip with the upper-cascade halo recomputed locally (rc_config.tile_*; bit-identical to
  the single-GPU frame, tests/test_gpu_parity.py::test_tile_equals_full_frame_crop) and the
  finished strips are exchanged with ONE collective, an NCCL all-gather of the RGBA16F
  irradiance strips (66 MB in total at 4K) over NVLink/NVSwitch.

The collective is torch.distributed plumbing; rendering is librc_b200.so.  On CPU (gloo)
only the host-side logic below runs — there is no CPU rendering path.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

Tile = Tuple[int, int, int, int]  # x0, y0, w, h


def partition_strips(width: int, height: int, n: int, align: int = 4) -> List[Tile]:
    """`n` horizontal strips covering the frame exactly; heights differ by at most `align`
    rows (strip boundaries sit on multiples of `align`, the level-0 probe spacing, so that
    neighbouring strips share as few level-0 probes as possible)."""
    if n < 1 or height < n:
        raise ValueError("cannot cut %d rows into %d strips" % (height, n))
    units = -(-height // align)            # rows in units of `align`
    if units < n:
        align, units = 1, height
    base, extra = divmod(units, n)
    tiles, y = [], 0
    for r in range(n):
        rows = (base + (1 if r < extra else 0)) * align
        rows = min(rows, height - y)
        if r == n - 1:
            rows = height - y
        tiles.append((0, y, width, rows))
        y += rows
    assert y == height and all(t[3] > 0 for t in tiles)
    return tiles


def partition_grid(width: int, height: int, nx: int, ny: int, align: int = 4) -> List[Tile]:
    """nx x ny tiles (row-major rank order); the 2-D cut halves the cumulative halo of strips
    at 8 GPUs (4x2 tiles ~30 % redundant work vs ~56 % for 8 strips at 4K)."""
    cols = [(t[1], t[3]) for t in partition_strips(1, width, nx, align)]
    rows = [(t[1], t[3]) for t in partition_strips(1, height, ny, align)]
    return [(cx, ry, cw, rh) for (ry, rh) in rows for (cx, cw) in cols]


def strips_from_cuts(width: int, height: int, cuts: Sequence[int]) -> List[Tile]:
    """Full-width strips from the interior cut rows (ascending, exclusive of 0 and height)."""
    ys = [0] + [int(c) for c in cuts] + [height]
    assert all(b > a for a, b in zip(ys, ys[1:])), ys
    return [(0, a, width, b - a) for a, b in zip(ys, ys[1:])]


class StripBalancer:
    """Cost-balanced horizontal cuts from measured frame times (SURVEY §8e "cost-balanced cuts").

    Equal-height strips of one frame are badly balanced: on living_room 3840x2160 cut 8 ways the slowest strip
    takes 0.48 ms against a mean of 0.32 ms (profiles/r2_tile_timelines.md) — the cost of a ray differs 3x
    between the upper and the lower half of the screen, so neither rows nor ray counts predict it.  Each rank
    therefore reports the device time of its last frame; every rank runs this same deterministic update on
    the all-gathered times: per strip, cost density = (t - fixed) / rows (the fixed part — the short
    dependent kernels every tile runs regardless of its size — does not move with the cut); the new cuts
    equalise the integral of that piecewise-constant density, moved only `damping` of the way (the density
    inside a strip is not constant, so a full step overshoots), aligned to `align` rows."""

    def __init__(self, width: int, height: int, n: int, align: int = 4, min_rows: int = 16, fixed_ms: float = 0.07,
                 damping: float = 0.7):
        self.W, self.H, self.n = width, height, n
        self.align, self.min_rows, self.fixed_ms, self.damping = align, max(min_rows, align), fixed_ms, damping
        self.tiles = partition_strips(width, height, n, align)
        self.version = 0

    def cuts(self) -> List[int]:
        return [t[1] for t in self.tiles[1:]]

    def update(self, times_ms: Sequence[float]) -> bool:
        """New tiles from every rank's last frame time (rank order).  Returns True when a cut moved."""
        n = self.n
        assert len(times_ms) == n
        if n == 1:
            return False
        rows = [t[3] for t in self.tiles]
        dens = [max(float(t) - self.fixed_ms, 0.05 * max(float(t), 1e-6)) / h for t, h in zip(times_ms, rows)]
        total = sum(d * h for d, h in zip(dens, rows))
        target = total / n
        # walk the cumulative cost
        new_cuts, acc, k, y = [], 0.0, 0, 0
        want = target
        for r in range(n):
            h, d = rows[r], dens[r]
            while len(new_cuts) < n - 1 and acc + d * h >= want - 1e-12:
                frac = (want - acc) / (d * h) if d * h > 0 else 0.0
                new_cuts.append(y + frac * h)
                want += target
            acc += d * h
            y += h
        while len(new_cuts) < n - 1:
            new_cuts.append(float(self.H))
        old = self.cuts()
        moved = []
        for o, c in zip(old, new_cuts):
            v = o + self.damping * (c - o)
            moved.append(int(round(v / self.align)) * self.align)
        # keep every strip at least min_rows tall, front to back then back to front
        lo = 0
        for i in range(n - 1):
            moved[i] = max(moved[i], lo + self.min_rows)
            lo = moved[i]
        hi = self.H
        for i in range(n - 2, -1, -1):
            moved[i] = min(moved[i], hi - self.min_rows)
            hi = moved[i]
        if moved == old or any(b <= a for a, b in zip([0] + moved, moved + [self.H])):
            return False
        self.tiles = strips_from_cuts(self.W, self.H, moved)
        self.version += 1
        return True


def halo_overhead(width: int, height: int, tiles: Sequence[Tile], p0: int = 4, levels: int = 6) -> float:
    """Redundant work of halo recomputation: (probe-directions summed over tiles) / (full frame) - 1,
    from the same footprint recursion librc_b200 uses (rc_spec.h S1)."""
    from math import ceil

    def rects(tile):
        x0, y0, w, h = tile
        out = []
        xr = yr = None
        for i in range(levels):
            P = p0 << i
            gw, gh = ceil(width / P), ceil(height / P)
            if i == 0:
                xr = (max(0, min((x0 - p0 // 2) // p0, gw - 1)), max(0, min((x0 + w - 1 - p0 // 2) // p0 + 1, gw - 1)))
                yr = (max(0, min((y0 - p0 // 2) // p0, gh - 1)), max(0, min((y0 + h - 1 - p0 // 2) // p0 + 1, gh - 1)))
            else:
                lo = lambda q: q // 2 - 1 if q % 2 == 0 else (q - 1) // 2
                xr = (max(0, min(lo(xr[0]), gw - 1)), max(0, min(lo(xr[1]) + 1, gw - 1)))
                yr = (max(0, min(lo(yr[0]), gh - 1)), max(0, min(lo(yr[1]) + 1, gh - 1)))
            out.append((xr[1] - xr[0] + 1) * (yr[1] - yr[0] + 1) * (4 ** i))
        return out

    full = sum(rects((0, 0, width, height)))
    return sum(sum(rects(t)) for t in tiles) / full - 1.0


class _DevicePtr:
    """Zero-copy view of a device buffer owned by librc_b200 (consumed by torch.as_tensor)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def irradiance_tensor(renderer):
    """torch view (float16 [h][w][4]) of the renderer's irradiance tile in HBM — no copy."""
    import torch
    ptr, nbytes = renderer.irradiance_device_ptr()
    w, h = renderer.tile_size()
    assert nbytes == w * h * 8
    return torch.as_tensor(_DevicePtr(ptr, (h, w, 4), "<f2"), device="cuda")


def assemble_strips(parts: Sequence[np.ndarray], tiles: Sequence[Tile], width: int, height: int) -> np.ndarray:
    """Host-side assembly of per-rank tiles into the full frame (also the checker of the collective)."""
    ch = parts[0].shape[-1]
    out = np.zeros((height, width, ch), dtype=parts[0].dtype)
    for p, (x0, y0, w, h) in zip(parts, tiles):
        out[y0:y0 + h, x0:x0 + w] = p[:h, :w]
    return out


def all_gather_tiles(local, tiles: Sequence[Tile], width: int, height: int, group=None):
    """The one collective of tiled mode: all-gather of the ranks' irradiance tiles.  `local` is a
    torch tensor [h][w][4] on the backend's device (cuda for nccl, cpu for gloo).  Tiles are padded
    to the largest tile so a single all_gather_into_tensor moves everything; returns the assembled
    [height][width][4] tensor on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    assert world == len(tiles)
    if all(t[0] == 0 and t[2] == width and t[3] == tiles[0][3] for t in tiles) and local.is_contiguous():
        # equal full-width strips are exactly the row blocks of the frame: gather straight into it, no staging copies
        full = torch.empty((height, width, local.shape[-1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, local, group=group)
        return full
    mh = max(t[3] for t in tiles)
    mw = max(t[2] for t in tiles)
    pad = torch.zeros((mh, mw, local.shape[-1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0], :local.shape[1]] = local
    # concatenated along dim 0 (the layout both nccl and gloo accept), viewed per rank below
    gathered = torch.empty((world * mh, mw, local.shape[-1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    gathered = gathered.view(world, mh, mw, local.shape[-1])
    full = torch.zeros((height, width, local.shape[-1]), dtype=local.dtype, device=local.device)
    for r, (x0, y0, w, h) in enumerate(tiles):
        full[y0:y0 + h, x0:x0 + w] = gathered[r, :h, :w]
    return full


def _device_view(ptr: int, shape, typestr: str):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, shape, typestr), device="cuda")


class HaloExchanger:
    """The transport of RC_CFG_HALO_EXCHANGE for full-width strips: NCCL send / recv (torch.distributed P2P, grouped) of
    (a) request masks — every rank's requests of probes it does not own, to their owners, once per frame — and
    (b) child averages — each finished level's owned probe rows, to every rank whose sub-grid holds them, once per level.
    Strips make both contiguous row ranges of the library's row-major sub-grid buffers (rc_exchange_level_info), so the
    sends and receives work on zero-copy views of librc_b200's own device memory.  Every rank derives the same plan from the
    all-gathered level geometry; call refresh() (collectively) after the tiles moved."""

    def __init__(self, renderer, rank: int, world: int, group=None):
        self.r, self.rank, self.world, self.group = renderer, rank, world, group
        self.refresh()

    def refresh(self) -> None:
        import torch.distributed as dist
        n = len(self.r.levels())
        mine = []
        for L in range(n):
            e = self.r.exchange_level_info(L)
            mine.append(dict(exchanged=int(e.exchanged), px0=e.px0, py0=e.py0, sw=e.sub_w, sh=e.sub_h, oy0=e.own_y0, oy1=e.own_y1,
                             ox0=e.own_x0, ox1=e.own_x1, words=e.need_words_per_probe, f4=e.avg_float4_per_probe))
        allinfo = [None] * self.world
        dist.all_gather_object(allinfo, mine, group=self.group)
        self.levels, self.views = [], []
        for L in range(n):
            me = allinfo[self.rank][L]
            plan = dict(exchanged=bool(me["exchanged"]), recv_avg=[], send_avg=[])
            if me["exchanged"]:
                assert me["px0"] == 0 and me["ox0"] == 0 and me["ox1"] == me["sw"] - 1, "halo exchange transports full-width strips only"
                e = self.r.exchange_level_info(L)
                need = _device_view(e.need_ptr, (me["sh"], me["sw"] * me["words"]), "<i4")
                avg = _device_view(e.avg_ptr, (me["sh"], me["sw"] * me["f4"] * 4), "<f4")
                my_sub = (me["py0"], me["py0"] + me["sh"])                                # global probe rows [a, b)
                my_own = (me["py0"] + me["oy0"], me["py0"] + me["oy1"] + 1) if me["oy1"] >= me["oy0"] else (0, 0)
                for q in range(self.world):
                    if q == self.rank:
                        continue
                    o = allinfo[q][L]
                    q_sub = (o["py0"], o["py0"] + o["sh"])
                    q_own = (o["py0"] + o["oy0"], o["py0"] + o["oy1"] + 1) if o["oy1"] >= o["oy0"] else (0, 0)
                    a, b = max(q_own[0], my_sub[0]), min(q_own[1], my_sub[1])        # rows q owns that I hold: I receive their averages
                    if b > a:
                        plan["recv_avg"].append((q, a - me["py0"], b - me["py0"]))
                    a, b = max(my_own[0], q_sub[0]), min(my_own[1], q_sub[1])        # rows I own that q holds: I send their averages
                    if b > a:
                        plan["send_avg"].append((q, a - me["py0"], b - me["py0"]))
                self.views.append((need, avg))
            else:
                self.views.append((None, None))
            self.levels.append(plan)

    def exchange_masks(self) -> None:
        """My requests of probes owned by q go to q; q's requests of my probes are OR-ed into my masks (all levels, one batch)."""
        import torch
        import torch.distributed as dist
        ops, pending = [], []
        for L, plan in enumerate(self.levels):
            if not plan["exchanged"]:
                continue
            need = self.views[L][0]
            for q, a, b in plan["recv_avg"]:          # rows q owns -> q wants my requests of them
                ops.append(dist.P2POp(dist.isend, need[a:b], q, group=self.group))
            for q, a, b in plan["send_avg"]:          # rows I own -> I want q's requests of them
                tmp = torch.empty_like(need[a:b])
                ops.append(dist.P2POp(dist.irecv, tmp, q, group=self.group))
                pending.append((need[a:b], tmp))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for dst, tmp in pending:
            dst.bitwise_or_(tmp)

    def exchange_avg(self, L: int) -> None:
        """Level L has just been finalised for the probes each rank owns: fill everybody's halo rows."""
        import torch.distributed as dist
        plan = self.levels[L]
        if not plan["exchanged"]:
            return
        avg = self.views[L][1]
        ops = [dist.P2POp(dist.isend, avg[a:b], q, group=self.group) for q, a, b in plan["send_avg"]]
        ops += [dist.P2POp(dist.irecv, avg[a:b], q, group=self.group) for q, a, b in plan["recv_avg"]]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


class TiledRenderer:
    """One rank's share of a tiled frame: a DefaultRenderer restricted to this rank's tile plus the
    final all-gather.  Needs a CUDA device (no CPU fallback)."""

    def __init__(self, rank: int, world: int, device: int, size: Tuple[int, int], state, path: str,
                 cascade=None, grid: Optional[Tuple[int, int]] = None, balance: bool = False, halo_exchange: bool = False):
        from . import _ffi
        from .renderer import CascadeConfig, DefaultRenderer
        W, H = size
        self.balancer = StripBalancer(W, H, world) if (balance and not grid) else None
        self.tiles = partition_grid(W, H, *grid) if grid else partition_strips(W, H, world)
        self.rank, self.world, self.size = rank, world, (W, H)
        cc = cascade or CascadeConfig()
        cc.tile = self.tiles[rank]
        self.halo_exchange = bool(halo_exchange) and not grid and world > 1
        if self.halo_exchange:
            cc.flags |= _ffi.RC_CFG_HALO_EXCHANGE
        self.exchanger = None
        if self.balancer is not None and world > 1:
            # the cuts will move: size the (grow-only) device buffers once for a strip of three times the average height (balanced
            # 8-way cuts of the 4K living room reach 2.5x), so that re-tiling does not normally reach the allocator (a cudaFree +
            # cudaMalloc of the cascade stalls the rank for ~0.1 s; a taller strip still works, it just pays that once)
            rows = min(H, 3 * -(-H // world) + 64)
            y0 = min(self.tiles[rank][1], H - rows)
            cc.tile = (0, y0, W, rows)
        self.renderer = DefaultRenderer.new(device, (W, H), state, path, cc)
        if cc.tile != self.tiles[rank]:
            self.renderer.set_tile(self.tiles[rank])

    def render(self, state, stream: Optional[int] = None):
        self.renderer.update(state)
        if self.halo_exchange:
            return self._render_exchange(stream)
        self.renderer.render(stream)

    def _render_exchange(self, stream: Optional[int]):
        """RC_CFG_HALO_EXCHANGE frame: G-buffer / probes / request masks, mask exchange, ray lists of the owned probes, then the
        levels top-down with the child averages of each finished level sent to the neighbours before the level below merges.
        `stream` must be the handle of torch's CURRENT stream: the NCCL transfers are ordered on it."""
        import torch
        r = self.renderer
        cur = torch.cuda.current_stream().cuda_stream
        if stream is None:
            stream = cur
        assert stream == cur, "render the halo-exchange frame inside `with torch.cuda.stream(s)` and pass s.cuda_stream"
        if self.exchanger is None:
            self.exchanger = HaloExchanger(r, self.rank, self.world)
        r.render_begin(stream)
        self.exchanger.exchange_masks()
        r.render_lists(stream)
        for L in range(len(self.exchanger.levels) - 1, -1, -1):
            r.render_level(L, stream)
            self.exchanger.exchange_avg(L)
        r.render_end(stream)

    def rebalance(self, times_ms: Sequence[float]) -> bool:
        """Move the strip cuts from every rank's last frame time (identical input on every rank -> identical tiles);
        re-tiles this rank's context in place (rc_set_tile).  Returns True when the tiles changed."""
        if self.balancer is None or not self.balancer.update(times_ms):
            return False
        self.tiles = list(self.balancer.tiles)
        self.renderer.set_tile(self.tiles[self.rank])
        if self.exchanger is not None:
            self.exchanger.refresh()          # collective: every rank re-tiles in the same call
        return True

    def all_gather_times(self, my_ms: float, group=None) -> List[float]:
        """Host-side all-gather of one float per rank (control plane: a gloo group when given, else the default group)."""
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, float(my_ms), group=group)
        return [float(x) for x in out]

    def attach_peers(self, group=None):
        """Exchange CUDA IPC handles (host-side all-gather of 64-byte blobs) and map every rank's frame buffers:
        from the next render on, the gather kernel stores this rank's tile into all ranks' frames over NVLink."""
        import torch.distributed as dist
        mine = self.renderer.peer_export()
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self.renderer.peer_attach(handles, self.rank)
        dist.barrier(group=group)       # nobody renders before everybody has mapped everybody
        self.peers = True

    def gather_peer(self, stream: Optional[int] = None):
        """Fused path: enqueue the wait for all ranks' tiles of the last frame and return the assembled frame
        (torch view of this rank's peer buffer, float16 [H][W][4]); consume it on `stream` before the next render.
        `stream` defaults to torch's current stream — the stream the caller's consumer kernels run on; render on the
        same stream (the release handshake assumes render and consumption are stream-ordered)."""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream().cuda_stream
        self.renderer.peer_wait(stream)
        ptr, nbytes, _ = self.renderer.peer_frame()
        W, H = self.size
        assert nbytes == W * H * 8
        return torch.as_tensor(_DevicePtr(ptr, (H, W, 4), "<f2"), device="cuda")

    def gather(self, group=None):
        """NCCL all-gather of the finished tiles; returns the full frame (torch, cuda, float16)."""
        self.renderer.synchronize()
        return all_gather_tiles(irradiance_tensor(self.renderer), self.tiles, *self.size, group=group)

"""Multi-GPU partitioning of the stitching path -- one process per GPU, torch.distributed for the plumbing.

Only the three modes in which the path shards naturally:
  * independent streams (config 4)            : stream s -> rank s % world, no data-path collective
  * offline frame-pair sharding (config 3)    : contiguous frame chunks per rank (1-frame halo), per-pair RANSAC on the
                                                device, all_gather of the 3x3 relative homographies (72 B / pair), then
                                                the reference's sequential validate / smooth / prefix composition on
                                                every rank (tiny), failed pairs fixed up sequentially (main.py:722-731)
  * canvas row tiles (config 5)               : rank g owns canvas rows [g*Hc/G, (g+1)*Hc/G); it warps+blends every frame
                                                whose window touches its tile with H shifted by the tile origin; tiles are
                                                gathered with one collective at the end.
A single online stream on one canvas does not shard (three loop-carried dependencies): replicas only.

The host-side logic here is pure NumPy / torch and is covered by world_size-2 gloo tests on CPU (tests/test_sharding_cpu.py).
"""
from __future__ import annotations

import numpy as np

OK, SKIP_FEW, SKIP_NO_H = 0, 1, 2


# ---------------------------------------------------------------------------------------------------------------
# partitioning helpers
# ---------------------------------------------------------------------------------------------------------------
def shard_streams(n_streams: int, rank: int, world: int):
    """config 4: streams assigned round-robin."""
    return [s for s in range(n_streams) if s % world == rank]


def shard_pairs(n_frames: int, rank: int, world: int):
    """config 3 offline: pairs (t-1, t), t = 1..n_frames-1, split into contiguous chunks.  Returns (t_start, t_end): this
    rank estimates pairs t_start <= t < t_end and therefore needs frames t_start-1 .. t_end-1 (1-frame halo)."""
    n_pairs = max(n_frames - 1, 0)
    base, rem = divmod(n_pairs, world)
    start = 1 + rank * base + min(rank, rem)
    end = start + base + (1 if rank < rem else 0)
    return start, end


def tile_rows(canvas_h: int, rank: int, world: int):
    """config 5: contiguous row tiles, multiples of 16 rows (the block grid of the distance-transform tables)."""
    per = -(-canvas_h // world)
    per = -(-per // 16) * 16
    y0 = min(rank * per, canvas_h)
    y1 = min(y0 + per, canvas_h)
    return y0, y1


# ---------------------------------------------------------------------------------------------------------------
# the reference's host control flow on a list of per-pair results (main.py:734-746), used after the all_gather
# ---------------------------------------------------------------------------------------------------------------
def validate_homography(H, translation_threshold=50.0, scale_threshold=0.3):
    """main.py:761-801 without the prints: the library's one implementation (host-only entry point, no device needed)."""
    from . import _lib
    return _lib.validate_homography(H, translation_threshold, scale_threshold)[0] == _lib.BM_VAL_OK


def compose_chain(H0, rel, history_size=5):
    """Sequential part of the path on per-pair relative homographies: validate -> identity substitution, 5-tap weighted
    smoothing over the history (main.py:803-834), H_t = H_{t-1} @ H_s (main.py:746).  `rel` is a list of 3x3 arrays or None
    (None = the pair was skipped: state not advanced, no output for that frame).  Returns a list of absolute H or None."""
    H_old = np.array(H0, dtype=np.float64)
    hist, out = [], []
    for Hr in rel:
        if Hr is None:
            out.append(None)
            continue
        Hv = Hr if validate_homography(Hr) else np.eye(3)
        hist.append(np.array(Hv, dtype=np.float64))
        if len(hist) > history_size:
            hist.pop(0)
        if len(hist) < 2:
            Hs = Hv
        else:
            w = np.linspace(0.5, 1.0, len(hist))
            w = w / np.sum(w)
            Hs = np.zeros((3, 3))
            for wi, h in zip(w, hist):
                Hs += wi * h
        H_old = H_old @ Hs
        out.append(H_old.copy())
    return out


def pack_pairs(statuses, Hs):
    """(n,) int statuses + list of H -> float64 (n, 10) rows [status, h0..h8] for the all_gather"""
    a = np.zeros((len(statuses), 10), np.float64)
    for i, (s, H) in enumerate(zip(statuses, Hs)):
        a[i, 0] = s
        if H is not None:
            a[i, 1:] = np.asarray(H, np.float64).reshape(9)
    return a


def unpack_pairs(a):
    return [(None if int(r[0]) != OK else r[1:].reshape(3, 3).copy()) for r in a]


def all_gather_pairs(local_rows, n_frames, rank, world, dist=None, device="cpu"):
    """all_gather of ragged per-rank (n_local, 10) arrays -> (n_frames-1, 10) in pair order.  72 B per pair: latency-bound,
    one collective.  Works with gloo (CPU tests) and nccl (pass device='cuda')."""
    import torch
    counts = [shard_pairs(n_frames, r, world) for r in range(world)]
    mx = max(e - s for s, e in counts) if counts else 0
    buf = torch.zeros((max(mx, 1), 10), dtype=torch.float64, device=device)
    if len(local_rows):
        buf[:len(local_rows)] = torch.from_numpy(np.asarray(local_rows)).to(device)
    if dist is None or world == 1:
        gathered = [buf]
    else:
        gathered = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(gathered, buf)
    rows = []
    for r, (s, e) in enumerate(counts):
        rows.append(gathered[r][:e - s].cpu().numpy())
    return np.concatenate(rows, axis=0) if rows else np.zeros((0, 10))


# ---------------------------------------------------------------------------------------------------------------
# canvas row tiles
# ---------------------------------------------------------------------------------------------------------------
def tile_homography(H, y0):
    """homography into tile-local canvas coordinates: rows shifted by the tile origin"""
    T = np.eye(3)
    T[1, 2] = -float(y0)
    return T @ np.asarray(H, dtype=np.float64)


def window_rows(H, frame_w, frame_h):
    """row extent [ymin, ymax] of the warped frame quad on the canvas (conservative, +-4 px)"""
    c = np.array([[-1, -1, 1], [frame_w, -1, 1], [frame_w, frame_h, 1], [-1, frame_h, 1]], np.float64).T
    q = np.asarray(H, np.float64) @ c
    if np.any(q[2] <= 1e-9):
        return -np.inf, np.inf
    y = q[1] / q[2]
    return float(y.min()) - 4.0, float(y.max()) + 4.0


def touches_tile(H, frame_w, frame_h, y0, y1):
    lo, hi = window_rows(H, frame_w, frame_h)
    return hi >= y0 and lo < y1


def gather_tiles(tile, canvas_h, rank, world, dist=None, out=None):
    """tile: torch uint8 (rows_r, Wc, 3) on this rank's device; returns the full canvas on every rank (all_gather; NCCL over
    NVLink on the GPU box, gloo in the CPU tests).  `out`: a preallocated (canvas_h, Wc, 3) tensor to gather into -- with equal tiles
    the collective then writes the canvas in place (no padded staging copy, no concatenation, no allocation inside the call).
    Otherwise tiles are padded to the common size for the collective."""
    import torch
    spans = [tile_rows(canvas_h, r, world) for r in range(world)]
    per = max(e - s for s, e in spans)
    if out is not None and all(e - s == per for s, e in spans) and tile.is_contiguous() and out.is_contiguous():
        if dist is None or world == 1:
            out.copy_(tile)
        else:
            dist.all_gather_into_tensor(out.view(-1), tile.view(-1))
        return out
    pad = torch.zeros((per,) + tuple(tile.shape[1:]), dtype=tile.dtype, device=tile.device)
    pad[:tile.shape[0]] = tile
    if dist is None or world == 1:
        parts = [pad]
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
    full = torch.cat([parts[r][:spans[r][1] - spans[r][0]] for r in range(world)], dim=0)
    if out is not None:
        out.copy_(full)
        return out
    return full


# ---------------------------------------------------------------------------------------------------------------
# drivers on top of VideMosaic handles (GPU)
# ---------------------------------------------------------------------------------------------------------------
def estimate_pairs(frames, t_start, t_end, detector_type="sift", device=0, vm=None):
    """per-pair relative homographies for pairs t_start <= t < t_end; `frames` is indexable by absolute frame index.
    `vm`: an existing handle whose previous frame is frames[t_start - 1] (reused, not closed); otherwise one is created."""
    from .mosaic import VideMosaic
    if t_end <= t_start:
        return [], []
    own = vm is None
    if own:
        vm = VideMosaic(frames[t_start - 1], detector_type=detector_type, show_intermediate=False, visualize=False, device=device)
    st, Hs = [], []
    for t in range(t_start, t_end):
        s, H, _ = vm.estimate_frame(frames[t], frames[t + 1] if t + 1 < t_end else None,            # the next frames' uploads and detects overlap
                                    [frames[u] for u in (t + 2, t + 3) if u < t_end])
        st.append(s)
        Hs.append(H)
    if own:
        vm.close()
    return st, Hs


def fixup_skipped(frames, rel, detector_type="sift", device=0):
    """The reference does not advance its 'previous' frame when a pair is skipped (main.py:722-731): the next frame is matched
    against the last frame that was accepted.  Pairs after a failure are therefore re-estimated sequentially (rare)."""
    from .mosaic import VideMosaic
    rel = list(rel)
    t = 1
    n = len(rel) + 1
    while t < n:
        if rel[t - 1] is not None:
            t += 1
            continue
        prev = t - 1                      # last accepted frame
        vm = VideMosaic(frames[prev], detector_type=detector_type, show_intermediate=False, visualize=False, device=device)
        u = t + 1
        while u < n:
            # feature state must stay at `prev` until a pair succeeds: use the full process path's skip semantics
            vm2_status, H, _ = vm.estimate_frame(frames[u])
            if vm2_status == OK:
                rel[u - 1] = H
                break
            vm.close()
            vm = VideMosaic(frames[prev], detector_type=detector_type, show_intermediate=False, visualize=False, device=device)
            rel[u - 1] = None
            u += 1
        vm.close()
        t = u + 1
    return rel


def frame_window(H, frame_w, frame_h):
    """conservative integer bounding box (x0, y0, x1, y1), half open, of the warped frame quad on the full canvas (+-4 px slack);
    every process computes the same box from the same H, which is what keeps the tile protocol below in lock step"""
    c = np.array([[-1, -1, 1], [frame_w, -1, 1], [frame_w, frame_h, 1], [-1, frame_h, 1]], np.float64).T
    q = np.asarray(H, np.float64) @ c
    if np.any(q[2] <= 1e-9):
        raise ValueError("frame_window: homography sends a frame corner to infinity")
    x, y = q[0] / q[2], q[1] / q[2]
    return int(np.floor(x.min())) - 4, int(np.floor(y.min())) - 4, int(np.ceil(x.max())) + 5, int(np.ceil(y.max())) + 5


class TilePlanner:
    """Host-side protocol of the row tiles (pure logic, no device): which tiles blend a frame, which boundary sweep states have to be
    handed over first, which halo rectangles have to be copied afterwards.  Every process runs an identical planner on the same
    homographies, so all of them derive the same ordered list of operations and execute the ones they take part in."""

    def __init__(self, canvas_w, canvas_h, world, frame_w, frame_h, halo_rows):
        self.G, self.Wc, self.Hc = int(world), int(canvas_w), int(canvas_h)
        self.fw, self.fh, self.P = int(frame_w), int(frame_h), int(halo_rows)
        if self.P % 16:
            raise ValueError("halo_rows must be a multiple of 16 (the block grid of the distance-transform tables)")
        self.own = [tile_rows(self.Hc, g, self.G) for g in range(self.G)]
        T = min(y1 - y0 for y0, y1 in self.own[:-1]) if self.G > 1 else self.Hc
        if self.G > 1 and self.P > T - 16:
            raise ValueError(f"halo_rows {self.P} must be <= tile height - 16 = {T - 16}: a halo may only reach into the adjacent tile")
        self.ext = [(max(0, y0 - self.P), min(self.Hc, y1 + self.P)) if self.G > 1 else (0, self.Hc) for y0, y1 in self.own]
        # stale_down[k] = "ghost_top of tile k+1 does not reflect the current canvas above it"; stale_up[k] likewise for ghost_bot of
        # tile k-1.  Everything is stale at the start (an empty canvas above is all zero pixels, not a border).
        self.stale_down = [True] * self.G
        self.stale_up = [True] * self.G

    def _down(self, k, ops):
        if k < 0 or k >= self.G - 1 or not self.stale_down[k]:
            return
        self._down(k - 1, ops)                                  # tile k's own incoming state first
        ops.append(("carry", k, k + 1, 0, (self.ext[k + 1][0] - self.ext[k][0]) // 16 - 1))
        self.stale_down[k] = False

    def _up(self, k, ops):
        if k <= 0 or k > self.G - 1 or not self.stale_up[k]:
            return
        if self.ext[k - 1][1] >= self.Hc:                       # tile k-1's extended canvas reaches the canvas bottom: border
            self.stale_up[k] = False
            return
        self._up(k + 1, ops)
        ops.append(("carry", k, k - 1, 1, (self.ext[k - 1][1] - self.ext[k][0]) // 16))
        self.stale_up[k] = False

    def plan(self, H):
        """ordered operations for one frame: ("carry", src, dst, up, block) | ("blend", tile) | ("rect", src, dst, x0, y_abs, w, h)"""
        xa, wa, xb, wb = frame_window(H, self.fw, self.fh)
        wa, wb = max(wa, 0), min(wb, self.Hc)
        xa, xb = max(xa, 0), min(xb, self.Wc)
        ops = []
        if wa >= wb or xa >= xb:
            return ops
        S = [g for g in range(self.G) if self.own[g][0] < wb and self.own[g][1] > wa]
        for g in S:
            if self.ext[g][0] > wa or self.ext[g][1] < wb:
                raise ValueError(f"frame rows [{wa},{wb}) leave the extended canvas {self.ext[g]} of tile {g}: halo_rows = {self.P} is too small")
        for g in S:                                             # the boundary states the blending tiles need, refreshed if out of date
            self._down(g - 1, ops)
            self._up(g + 1, ops)
        ops += [("blend", g) for g in S]
        for g in S:                                             # halo rows of neighbours that did not blend this frame themselves
            for n in (g - 1, g + 1):
                if n < 0 or n >= self.G or n in S:
                    continue
                ya, yb = max(wa, self.own[g][0], self.ext[n][0]), min(wb, self.own[g][1], self.ext[n][1])
                if ya < yb:
                    ops.append(("rect", g, n, xa, ya, xb - xa, yb - ya))
        for k in range(self.G - 1):                             # rows [wa, wb) changed: which exported states they feed
            if self.ext[k + 1][0] - 1 >= wa:
                self.stale_down[k] = True
        for k in range(1, self.G):
            if self.ext[k - 1][1] < wb:
                self.stale_up[k] = True
        return ops


class TileGroup:
    """config 5: a (Wc, Hc) canvas cut into `world` row tiles that reproduce the UNTILED canvas bit for bit.

    Tile g owns rows [y0_g, y1_g) and keeps an extended local canvas [y0_g - P, y1_g + P) (P = halo_rows): a frame that touches a
    tile's own rows lies entirely inside the extended canvas, so warpPerspective, mask_new and cv2.distanceTransform(mask_new)
    (main.py:871-888) are computed locally exactly as on the full canvas.  cv2.distanceTransform(mask_old) (main.py:889) is global
    over the canvas; its sweeps are continued across tile boundaries: the downward sweep state (E1, E2, V per column, see dt.cu) at
    the row just above a tile's extended canvas comes from the tile above, the upward state from the tile below (bm_tile_export_carries
    -> bm_tile_set_ghost, 3 x Wc x 4 bytes per hop).  They are refreshed LAZILY: every process tracks, from the frames' windows alone,
    which boundary states are out of date, and a hop (or a chain of hops through unchanged tiles) only happens when a tile is about to
    blend and something above / below it has changed since.  Halo rows changed by a neighbour that did not blend the frame itself are
    copied over (bm_tile_export_rect -> bm_tile_import_rect).  A camera that stays inside one tile needs no communication at all.

    `local_tiles`: the tiles this process holds -- [rank] in a distributed run (one tile per rank, NCCL send / recv between
    neighbours), all of them in a single-process run (tests: the same protocol with device-to-device copies)."""

    exchange = True

    def __init__(self, frame0, canvas_w, canvas_h, world, local_tiles, halo_rows, dist=None, device=0):
        import torch
        from .mosaic import VideMosaic
        self.torch, self.dist = torch, dist
        self.planner = TilePlanner(canvas_w, canvas_h, world, frame0.shape[1], frame0.shape[0], halo_rows)
        self.G, self.Wc, self.Hc, self.P = self.planner.G, self.planner.Wc, self.planner.Hc, self.planner.P
        self.own, self.ext = self.planner.own, self.planner.ext
        self.local = {}
        for g in local_tiles:
            Y0, Y1 = self.ext[g]
            vm = VideMosaic(frame0, detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(Y1 - Y0, self.Wc), device=device)
            vm.clear_canvas()
            self.local[g] = vm
        self.hops = 0
        self.rect_bytes = 0
        if dist is not None and self.G > 1:                    # open the neighbour channels now (NCCL sets P2P connections up on first use)
            t = torch.zeros(8, dtype=torch.int32, device="cuda")
            ops = []
            for g in self.local:
                for n in (g - 1, g + 1):
                    if 0 <= n < self.G and n not in self.local:
                        ops.append(dist.P2POp(dist.isend, t, n))
                        ops.append(dist.P2POp(dist.irecv, torch.empty_like(t), n))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            torch.cuda.synchronize()

    # ---- transport: tile a -> tile b; every process walks the same sequence of calls ------------------------------------------
    def _xfer(self, a, b, make, consume, shape, dtype):
        torch = self.torch
        if a in self.local and b in self.local:
            consume(make())
        elif a in self.local:
            t = make()
            torch.cuda.current_stream().synchronize()
            self.dist.send(t, dst=b)
        elif b in self.local:
            t = torch.empty(shape, dtype=dtype, device="cuda")
            self.dist.recv(t, src=a)
            torch.cuda.current_stream().synchronize()
            consume(t)

    def _export_carries(self, g, up, block):
        import ctypes as C
        from . import _lib
        vm = self.local[g]
        t = self.torch.empty((3, self.Wc), dtype=self.torch.int32, device="cuda")
        _lib.check(vm._lib.bm_tile_export_carries(vm._h, int(up), int(block), C.c_void_p(t.data_ptr())), "bm_tile_export_carries")
        return t

    def _set_ghost(self, g, side, t):
        import ctypes as C
        from . import _lib
        vm = self.local[g]
        _lib.check(vm._lib.bm_tile_set_ghost(vm._h, int(side), C.c_void_p(t.data_ptr()) if t is not None else None), "bm_tile_set_ghost")

    def _copy_rect(self, g, n, x0, ya, w, h):
        import ctypes as C
        from . import _lib
        torch = self.torch

        def make():
            vm = self.local[g]
            t = torch.empty((h, w, 4), dtype=torch.uint8, device="cuda")
            _lib.check(vm._lib.bm_tile_export_rect(vm._h, x0, ya - self.ext[g][0], w, h, C.c_void_p(t.data_ptr())), "bm_tile_export_rect")
            return t

        def consume(t):
            vm = self.local[n]
            _lib.check(vm._lib.bm_tile_import_rect(vm._h, x0, ya - self.ext[n][0], w, h, C.c_void_p(t.data_ptr())), "bm_tile_import_rect")
            vm._canvas_cache = None
        self._xfer(g, n, make, consume, (h, w, 4), torch.uint8)
        self.rect_bytes += h * w * 4

    def put(self, frame, H):
        """VideMosaic.warp(frame, H) (main.py:861-927) on the tiled canvas; H is the absolute canvas homography.  Returns how many of
        this process's tiles blended the frame."""
        done = 0
        for op in self.planner.plan(H):
            if op[0] == "carry":
                _, a, b, up, block = op
                self._xfer(a, b, lambda: self._export_carries(a, up, block), lambda t: self._set_ghost(b, 1 if up else 0, t), (3, self.Wc), self.torch.int32)
                self.hops += 1
            elif op[0] == "blend":
                if op[1] in self.local:
                    self.local[op[1]].warp_nosync(frame, tile_homography(H, self.ext[op[1]][0]))
                    done += 1
            else:
                _, g, n, x0, ya, w, h = op
                self._copy_rect(g, n, x0, ya, w, h)
        return done

    def sync(self):
        for vm in self.local.values():
            vm.sync()

    def tile_tensor(self, g=None):
        """own rows of tile g (default: this process's only tile) as a packed-BGR uint8 cuda tensor"""
        torch = self.torch
        if g is None:
            (g,) = self.local.keys()
        vm = self.local[g]
        Y0, Y1 = self.ext[g]
        t = torch.empty((Y1 - Y0, self.Wc, 3), dtype=torch.uint8, device="cuda")
        vm.canvas_to_device(t.data_ptr())
        y0, y1 = self.own[g]
        return t[y0 - Y0:y1 - Y0].contiguous()

    def close(self):
        for vm in self.local.values():
            vm.close()
        self.local = {}


class TileStitcher:
    """config 5: this rank's row tile of a (Wc, Hc) canvas.  `put(frame, H)` warps + blends the frame into the tile if its window
    touches the tile's rows (H is the absolute canvas homography, shifted here by the tile origin); `tile_tensor()` returns the
    tile as a packed-BGR torch tensor for the final gather."""

    exchange = False        # boundary exchange between neighbouring tiles

    def __init__(self, frame0, canvas_w, canvas_h, rank, world, dist=None, device=0):
        from .mosaic import VideMosaic
        self.rank, self.world, self.dist = rank, world, dist
        self.Wc, self.Hc = int(canvas_w), int(canvas_h)
        self.fh, self.fw = frame0.shape[:2]
        self.y0, self.y1 = tile_rows(self.Hc, rank, world)
        self.vm = VideMosaic(frame0, detector_type="orb", show_intermediate=False, visualize=False,
                             canvas_size=(self.y1 - self.y0, self.Wc), device=device)
        self.vm.clear_canvas()

    def put(self, frame, H):
        if not touches_tile(H, self.fw, self.fh, self.y0, self.y1):
            return 0
        self.vm.warp_nosync(frame, tile_homography(H, self.y0))
        return 1

    def sync(self):
        self.vm.sync()

    def tile_tensor(self):
        import torch
        t = torch.empty((self.y1 - self.y0, self.Wc, 3), dtype=torch.uint8, device="cuda")
        self.vm.canvas_to_device(t.data_ptr())
        return t

    def close(self):
        self.vm.close()

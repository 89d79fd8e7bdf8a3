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


def gather_tiles(tile, canvas_h, rank, world, dist=None):
    """tile: torch uint8 (rows_r, Wc, 3) on this rank's device; returns the full canvas on every rank (all_gather; NCCL over
    NVLink on the GPU box, gloo in the CPU tests).  Tiles are padded to the common size for the collective."""
    import torch
    spans = [tile_rows(canvas_h, r, world) for r in range(world)]
    per = max(e - s for s, e in spans)
    pad = torch.zeros((per,) + tuple(tile.shape[1:]), dtype=tile.dtype, device=tile.device)
    pad[:tile.shape[0]] = tile
    if dist is None or world == 1:
        parts = [pad]
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
    return torch.cat([parts[r][:spans[r][1] - spans[r][0]] for r in range(world)], dim=0)


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
        s, H, _ = vm.estimate_frame(frames[t], frames[t + 1] if t + 1 < t_end else None)      # next frame's upload overlaps
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

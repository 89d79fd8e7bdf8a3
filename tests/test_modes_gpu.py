"""GPU tests of the sharded modes on one device: the split begin/end form used for concurrent streams, the
offline per-pair estimation + prefix composition, and canvas row tiles."""
import numpy as np
import cv2
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sweep_frames():
    from b200mosaic.synth import DroneSweep
    sw = DroneSweep(640, 360, seed=7, ground_size=2048, max_step=9.0)
    return sw.frames(7), sw


def test_interleaved_streams_equal_sequential(sweep_frames):
    import b200mosaic
    frames, _ = sweep_frames
    from b200mosaic.synth import DroneSweep
    frames_b = DroneSweep(640, 360, seed=8, ground_size=2048, max_step=9.0).frames(7)
    seq = []
    for fr in (frames, frames_b):
        vm = b200mosaic.VideMosaic(fr[0], detector_type="orb", show_intermediate=False, visualize=False)
        for t in range(1, 7):
            vm.process_frame(fr[t], t)
        seq.append((vm.output_img.copy(), vm.H.copy()))
    a = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    b = b200mosaic.VideMosaic(frames_b[0], detector_type="orb", show_intermediate=False, visualize=False)
    pin_a = torch.from_numpy(np.stack(frames)).pin_memory(); pin_b = torch.from_numpy(np.stack(frames_b)).pin_memory()
    fb = frames[0].nbytes
    for t in range(1, 7):
        a.begin_frame_ptr(pin_a.data_ptr() + t * fb)
        b.begin_frame_ptr(pin_b.data_ptr() + t * fb)
        assert a.end_frame() == 0 and b.end_frame() == 0
    assert np.array_equal(a.output_img, seq[0][0]) and np.array_equal(b.output_img, seq[1][0])
    assert np.array_equal(np.array(a.last_info.H).reshape(3, 3), seq[0][1])


def test_offline_pairs_and_prefix_composition(sweep_frames):
    import b200mosaic
    from b200mosaic import sharding as sh
    frames, _ = sweep_frames
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False)
    H0 = vm.H_old.copy()
    want_rel, want_abs = [], []
    for t in range(1, 7):
        vm.process_frame(frames[t], t)
        want_rel.append(np.array(vm.last_info.H_rel).reshape(3, 3)); want_abs.append(vm.H.copy())
    # two "ranks" on one device: pairs [1,4) and [4,7)
    rows = []
    for r in range(2):
        s, e = sh.shard_pairs(len(frames), r, 2)
        st, Hs = sh.estimate_pairs(frames, s, e, detector_type="orb")
        rows.append(sh.pack_pairs(st, Hs))
    rel = sh.unpack_pairs(np.concatenate(rows))
    for a, b in zip(rel, want_rel):
        assert np.array_equal(a, b)                       # same kernels, same inputs -> bit-identical H_rel
    got = sh.compose_chain(H0, rel)
    for a, b in zip(got, want_abs):
        assert np.abs(a - b).max() < 1e-9


def test_canvas_row_tiles_match_untiled(sweep_frames):
    """two row tiles on one device; the sweep stays inside the lower tile, so every tile must equal the same rows of the
    untiled canvas bit for bit; assembled with the same gather helper the NCCL path uses."""
    import b200mosaic
    from b200mosaic import sharding as sh
    frames, sw = sweep_frames
    fh, fw = frames[0].shape[:2]
    Hc, Wc = 3 * fh + 24, int(1.2 * fw)                 # frame 0 sits at the bottom; 6 steps of <= 9 px stay in the lower tile
    full = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(Hc, Wc))
    H = full.H_old.copy()
    Hs = [H.copy()]
    for t in range(1, 7):
        H = H @ sw.D_true[t - 1]
        Hs.append(H.copy())
        full.warp(frames[t], H)
    want = full.output_img
    tiles = []
    for r in range(2):
        y0, y1 = sh.tile_rows(Hc, r, 2)
        vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(y1 - y0, Wc))
        vm.clear_canvas()
        for t in range(0, 7):
            if sh.touches_tile(Hs[t], fw, fh, y0, y1):
                vm.warp(frames[t], sh.tile_homography(Hs[t], y0))
        dev = torch.empty((y1 - y0, Wc, 3), dtype=torch.uint8, device="cuda")
        vm.canvas_to_device(dev.data_ptr())
        tiles.append(dev)
    got = torch.cat(tiles, dim=0).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    y0, y1 = sh.tile_rows(Hc, 0, 2)
    assert got[:y1].max() == 0                            # the upper tile was never touched


@pytest.mark.parametrize("det", ["orb", "sift"])
def test_prefetch_and_overlap_do_not_change_results(sweep_frames, det):
    """double-buffered ingest (next_frame=) and the detect / chain stream overlap are scheduling only: the canvas and the
    trajectory must be bit-identical to the plain sequential calls, also when a prefetched frame is not the one processed next"""
    import b200mosaic
    frames, _ = sweep_frames
    ref = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
    ref.set_overlap(False)
    for t in range(1, 7):
        ref.process_frame(frames[t], t)
    want_canvas, want_H = ref.output_img.copy(), ref.H.copy()
    vm = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
    vm.warm_up()                                 # graphs captured up front instead of lazily: scheduling only, too
    for t in range(1, 7):
        nxt = frames[t + 1] if t + 1 < 7 else None
        if t == 3:
            nxt = frames[1]                      # a stale prefetch: frame 4 must be uploaded again, not taken from the staged copy
        vm.process_frame(frames[t], t, next_frame=nxt)
    assert np.array_equal(vm.output_img, want_canvas)
    assert np.array_equal(vm.H, want_H)
    pin = torch.from_numpy(np.stack(frames)).pin_memory()
    fb = frames[0].nbytes
    vp = b200mosaic.VideMosaic(frames[0], detector_type=det, show_intermediate=False, visualize=False)
    for t in range(1, 7):
        assert vp.process_frame_ptr(pin.data_ptr() + t * fb, pin.data_ptr() + (t + 1) * fb if t + 1 < 7 else None) == 0
    assert np.array_equal(vp.output_img, want_canvas)


@pytest.mark.parametrize("det", ["orb", "sift"])
def test_pipelined_calls_survive_skipped_frames(sweep_frames, det):
    """a featureless frame in the middle is skipped (main.py:722-724: state not advanced); with next_frame= pipelining (staged
    upload + early begin) the statuses, the trajectory and the canvas must equal the plain sequential run"""
    import b200mosaic
    frames, _ = sweep_frames
    seq = list(frames[:3]) + [np.full_like(frames[0], 7)] + list(frames[3:])
    def run(pipelined):
        vm = b200mosaic.VideMosaic(seq[0], detector_type=det, show_intermediate=False, visualize=False)
        st = []
        for t in range(1, len(seq)):
            nxt = seq[t + 1] if (pipelined and t + 1 < len(seq)) else None
            vm.process_frame(seq[t], t, next_frame=nxt)
            st.append(vm.last_info.status)
        return st, vm.H.copy(), vm.output_img.copy(), [(m.queryIdx, m.trainIdx) for m in vm.matches]
    a, b = run(False), run(True)
    assert a[0] == b[0] and a[0][2] != 0 and a[0].count(0) == len(seq) - 2      # exactly the blank frame is skipped
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[3] == b[3]                                                         # matches of the last finished frame


def test_long_pipelined_run_equals_serial(sweep_frames):
    """30 frames with detect-ahead, the estimate stream and both detector instances in flight (next_frame= on every call) against
    strictly serial calls with overlap off: same statuses, same trajectory, same canvas, same final features"""
    import b200mosaic
    from b200mosaic.synth import DroneSweep
    frames = DroneSweep(640, 360, seed=21, ground_size=2048, max_step=7.0).frames(31)
    def run(pipelined):
        vm = b200mosaic.VideMosaic(frames[0], detector_type="sift", show_intermediate=False, visualize=False)
        vm.set_overlap(pipelined)
        st, Hs = [], []
        for t in range(1, len(frames)):
            vm.process_frame(frames[t], t, next_frame=frames[t + 1] if (pipelined and t + 1 < len(frames)) else None)
            st.append(vm.last_info.status); Hs.append(vm.H.copy())
        kp = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in vm.kp_prev])
        return st, np.array(Hs), vm.output_img.copy(), kp, np.asarray(vm.des_prev).copy()
    a, b = run(False), run(True)
    assert a[0] == b[0] and a[0].count(0) >= 28
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2])
    assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])


def test_read_canvas_into_caller_buffers(sweep_frames):
    """bm_get_canvas into pinned and pageable caller-owned buffers (the pageable path is chunked through pinned staging)"""
    import b200mosaic
    frames, _ = sweep_frames
    vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(1500, 2100))
    for t in range(1, 4):
        vm.process_frame(frames[t], t)
    want = vm.output_img
    pinned = torch.empty(want.shape, dtype=torch.uint8).pin_memory().numpy()
    assert np.array_equal(vm.read_canvas(pinned), want)                  # 9.4 MB: three staging chunks on the pageable path above
    assert np.array_equal(vm.read_canvas(np.zeros_like(want)), want)
    with pytest.raises(ValueError):
        vm.read_canvas(np.zeros((4, 4, 3), np.uint8))


# ---- validate_homography rejections driven through bm_process_frame (main.py:734-737, 786-799) -----------------------------------
def _reject_case(kind):
    """(frame0, frame1): frame1 relates to frame0 by a homography the reference rejects for `kind`"""
    import cv2
    from b200mosaic.synth import make_ground
    g = make_ground(1536, seed=21)
    w, h = 640, 360
    x0, y0 = 400, 500
    f0 = g[y0:y0 + h, x0:x0 + w].copy()
    if kind == "translation":                              # 80 px > translation_threshold 50
        f1 = g[y0 - 80:y0 - 80 + h, x0:x0 + w].copy()
    elif kind == "scale":                                  # 1.5x zoom about the origin: no translation, sqrt(det) = 0.667 -> 0.33 > 0.3
        f1 = cv2.resize(f0[0:240, 0:427], (w, h), interpolation=cv2.INTER_LINEAR)
    elif kind == "perspective":                            # projective term about the origin: no translation, det = 1, |h31| = 1.5e-3 > 1e-3
        M = np.array([[1, 0, 0], [0, 1, 0], [1.5e-3, 0, 1.0]])
        f1 = cv2.warpPerspective(f0, M, (w, h), flags=cv2.INTER_LINEAR, borderValue=(1, 1, 1))
    else:
        f1 = g[y0 - 9:y0 - 9 + h, x0 + 2:x0 + 2 + w].copy()
    return f0, f1


@pytest.mark.parametrize("kind,reason", [("accept", 0), ("translation", 2), ("scale", 3), ("perspective", 4)])
def test_validate_rejections_through_the_abi(kind, reason, capsys):
    """the reject -> identity branch with its printed warnings, status BM_REJECTED_IDENTITY, state advance and history, against
    the oracle's RefMosaic on the same two frames"""
    import contextlib
    import io
    import b200mosaic
    from b200mosaic import _lib
    from oracle.mosaic_ref import RefMosaic
    f0, f1 = _reject_case(kind)
    ref = RefMosaic(f0, detector_type="orb")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.process_frame(f1)
    want_reject = "Невалидная гомография" in buf.getvalue()
    assert want_reject == (reason != 0), buf.getvalue()
    vm = b200mosaic.VideMosaic(f0, detector_type="orb", show_intermediate=False, visualize=False)
    capsys.readouterr()
    vm.process_frame(f1, 1)
    out = capsys.readouterr().out
    info = vm.last_info
    assert info.status == (_lib.BM_REJECTED_IDENTITY if reason else _lib.BM_OK)
    assert info.validate_reason == reason
    assert out.strip().splitlines() == buf.getvalue().strip().splitlines()          # the same warnings, word for word
    assert np.abs(vm.H - ref.H).max() < 1e-6 and np.abs(vm.H_old - ref.H_old).max() < 1e-6
    assert len(vm.homography_history) == len(ref.homography_history) == 1
    if reason:
        assert np.array_equal(vm.homography_history[0], np.eye(3))
    # the rejected frame still became "previous" (main.py:756-759) and was warped with the substituted identity
    kp = np.array([[k.pt[0], k.pt[1]] for k in vm.kp_prev], np.float32)
    kr = np.array([[k.pt[0], k.pt[1]] for k in ref.kp_prev], np.float32)
    assert np.array_equal(kp, kr)
    d = np.abs(vm.output_img.astype(np.int16) - ref.output_img.astype(np.int16))
    assert d.max() <= 1


def test_kp_cur_des_cur_and_output_img_setter():
    import b200mosaic
    f0, f1 = _reject_case("accept")
    vm = b200mosaic.VideMosaic(f0, detector_type="orb", show_intermediate=False, visualize=False)
    vm.process_frame(f1, 1)
    assert len(vm.kp_cur) == len(vm.des_cur) == vm.last_info.n_kp_cur
    assert np.array_equal(vm.des_cur, vm.des_prev)                       # accepted frame: cur became prev (main.py:757-758)
    img = np.full(vm.output_img.shape, 0, np.uint8); img[10:50, 20:90] = (5, 6, 7)
    vm.output_img = img.astype(np.float64)                               # reference callers hold a float64 canvas
    assert np.array_equal(vm.output_img, img)
    vm.stabilization_enabled = False
    vm.process_frame(f1, 2)
    assert len(vm.homography_history) == 1                               # smooth_homography returns before appending (:812-816)


def _tile_path(n_up=30, n_down=10, Wc=448, Hc=1024, fw=320, fh=180):
    """camera poses: from the canvas bottom up across every tile boundary with a sideways weave and +-3 degrees of rotation, then
    part of the way back down (revisits: boundary states go stale in both directions)"""
    ys = [Hc - fh - 6 - 25.0 * t for t in range(n_up)] + [Hc - fh - 6 - 25.0 * (n_up - 1) + 31.0 * t for t in range(1, n_down + 1)]
    Hs = []
    for t, y in enumerate(ys):
        ang = np.deg2rad(3.0 * np.sin(t / 4.0))
        c, s = np.cos(ang), np.sin(ang)
        R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        C = np.array([[1, 0, fw / 2], [0, 1, fh / 2], [0, 0, 1.0]])
        T = np.eye(3); T[0, 2] = (Wc - fw) / 2 + 40.0 * np.sin(t / 3.0); T[1, 2] = y
        Hs.append(T @ C @ R @ np.linalg.inv(C))
    return Hs


def test_row_tiles_with_boundary_exchange_equal_untiled_canvas():
    """four row tiles, frames that STRADDLE every tile boundary (and come back), frames with black holes.  With
    the sweep-state hand-over and the halo copies of sharding.TileGroup the assembled tiles equal the untiled canvas bit for bit
    ('a frame straddling a tile boundary blends differently from the untiled canvas = parity failure by design')."""
    import b200mosaic
    from b200mosaic import sharding as sh
    from b200mosaic.synth import DroneSweep
    fw, fh, Wc, Hc = 320, 180, 448, 1024
    Hs = _tile_path(Wc=Wc, Hc=Hc, fw=fw, fh=fh)
    base = DroneSweep(fw, fh, seed=5, ground_size=1024, max_step=6.0).frames(8)
    frames = []
    for t in range(len(Hs)):
        f = base[t % len(base)].copy()
        f[20 + 5 * (t % 7):44 + 5 * (t % 7), 30 + 11 * (t % 9):90 + 11 * (t % 9)] = 0      # holes in mask_new / mask_old
        if t % 5 == 0:
            f[:, 150:153] = 0                                                              # a black line from edge to edge
        frames.append(f)
    full = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(Hc, Wc))
    full.clear_canvas()
    for f, H in zip(frames, Hs):
        full.warp_nosync(f, H)
    want = full.output_img.copy()
    tg = sh.TileGroup(frames[0], Wc, Hc, world=4, local_tiles=[0, 1, 2, 3], halo_rows=208)
    touched = 0
    for f, H in zip(frames, Hs):
        touched += tg.put(f, H)
    tg.sync()
    got = torch.cat([tg.tile_tensor(g) for g in range(4)], dim=0).cpu().numpy()
    assert got.shape == want.shape
    assert touched > len(frames) + 8                       # many frames were blended by two tiles
    assert tg.hops >= 6 and tg.rect_bytes > 0              # sweep states crossed boundaries, halo rows were copied
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert np.array_equal(got, want), (int(d.max()), int((d > 0).sum()), sorted(set((np.nonzero(d.max(axis=(1, 2)))[0] // 256).tolist())))
    # and the exchange matters: tiles that ignore their neighbours (the round-1 mode) do differ on this path
    plain = []
    for g in range(4):
        y0, y1 = sh.tile_rows(Hc, g, 4)
        vm = b200mosaic.VideMosaic(frames[0], detector_type="orb", show_intermediate=False, visualize=False, canvas_size=(y1 - y0, Wc))
        vm.clear_canvas()
        for f, H in zip(frames, Hs):
            if sh.touches_tile(H, fw, fh, y0, y1):
                vm.warp_nosync(f, sh.tile_homography(H, y0))
        plain.append(vm.output_img.copy())
    assert not np.array_equal(np.concatenate(plain, axis=0), want)
    tg.close()


@pytest.mark.parametrize("det", ["orb", "sift"])
def test_multi_frame_lookahead_does_not_change_results(det):
    """next_frame= / next2_frame= / next3_frame= (up to three frames staged and detected ahead on the three detector instances) are
    scheduling only: statuses, trajectory, canvas and final features equal strictly serial calls -- also across a skipped (featureless)
    frame, stale staged frames and a caller that stops passing look-ahead frames in the middle of the run"""
    import b200mosaic
    from b200mosaic.synth import DroneSweep
    frames = DroneSweep(640, 360, seed=33, ground_size=2048, max_step=7.0).frames(26)
    seq = list(frames[:9]) + [np.full_like(frames[0], 9)] + list(frames[9:])           # a blank frame at index 9
    n = len(seq)

    def run(mode):
        vm = b200mosaic.VideMosaic(seq[0], detector_type=det, show_intermediate=False, visualize=False)
        vm.set_overlap(mode != "serial")
        st, Hs = [], []
        for t in range(1, n):
            n1 = seq[t + 1] if t + 1 < n else None
            n2 = seq[t + 2] if t + 2 < n else None
            n3 = seq[t + 3] if t + 3 < n else None
            if mode == "serial" or (mode == "mixed" and 14 <= t < 17):
                n1 = n2 = n3 = None                                                    # no look-ahead for a while
            elif mode == "mixed" and t == 5:
                n2 = seq[2]                                                            # a stale second look-ahead frame
            elif mode == "mixed" and t == 19:
                n1, n2 = seq[3], seq[t + 1]                                            # a stale first one
            elif mode == "mixed" and t == 11:
                n3 = seq[4]                                                            # a stale third one
            elif mode == "one":
                n2 = n3 = None
            elif mode == "two":
                n3 = None
            if n1 is None:
                n2 = None
            if n2 is None:
                n3 = None
            vm.process_frame(seq[t], t, next_frame=n1, next2_frame=n2, next3_frame=n3)
            st.append(vm.last_info.status)
            Hs.append(None if vm.H is None else vm.H.copy())
        kp = np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response] for k in vm.kp_prev])
        return st, Hs, vm.output_img.copy(), kp, np.asarray(vm.des_prev).copy(), [(m.queryIdx, m.trainIdx) for m in vm.matches]
    want = run("serial")
    assert want[0][8] != 0 and want[0].count(0) == n - 2                               # exactly the blank frame is skipped
    for mode in ("three", "two", "one", "mixed"):
        got = run(mode)
        assert got[0] == want[0], mode
        assert all((a is None) == (b is None) and (a is None or np.array_equal(a, b)) for a, b in zip(got[1], want[1])), mode
        assert np.array_equal(got[2], want[2]), mode
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[4], want[4]) and got[5] == want[5], mode

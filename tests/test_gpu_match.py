"""GPU (-m gpu): LINE-MOD parity of the CUDA path, called through the C ABI, against the C oracle and the golden fixtures.
Bit-exact for every stage: quantised images, spread images, linear memories, similarity maps, match lists."""
import os

import numpy as np
import pytest

import fealess_b200 as fb
import fl_oracle_py as F
from fealess_b200 import synth
from helpers import canonical, sha, tset_from_npz

pytestmark = pytest.mark.gpu


def _oracle(bgr, depth, T=(5, 8), masks=None):
    det = F.Detector(T)
    assert det.process(bgr, depth, masks) == 0
    return det


def _check_front_end(h, det, W, H, L=2, M=2):
    for l in range(L):
        for m in range(M):
            assert np.array_equal(h.debug_quantized(l, m, W, H), det.quantized(l, m)), "quantized L%d M%d" % (l, m)
            assert np.array_equal(h.debug_quantized(l, m, W, H, spread=True), det.spread(l, m)), "spread L%d M%d" % (l, m)
            for lab in range(8):
                assert np.array_equal(h.debug_lm(l, m, lab, W, H), det.lm(l, m, lab)), "LM L%d M%d label %d" % (l, m, lab)


@pytest.fixture(scope="module")
def vga():
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = _oracle(b, d, T)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(600, W, H, T, n_classes=4, seed=21, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    h.keep_spread(True)
    return W, H, b, d, det, ts, h


def test_front_end_bit_exact_vga(vga):
    W, H, b, d, det, ts, h = vga
    rc, m, q = h.match(b, d, 75.0, want_quantized=True)
    assert rc == 0
    for i in range(4):
        assert np.array_equal(q[i], det.quantized(i // 2, i % 2))
    _check_front_end(h, det, W, H)


@pytest.mark.parametrize("frame_idx", [1, 2, 3, 4])
def test_front_end_bit_exact_more_frames(vga, frame_idx):
    W, H, _, _, _, _, h = vga
    b, d = synth.make_frame(W, H, frame_idx)
    det = _oracle(b, d)
    assert h.match(b, d, 99.0)[0] == 0
    _check_front_end(h, det, W, H)


def test_edge_inputs(vga):
    W, H, _, _, _, _, h = vga
    rng = np.random.default_rng(0)
    cases = [(np.zeros((H, W, 3), np.uint8), np.zeros((H, W), np.uint16)),                         # empty frame
             (rng.integers(0, 256, (H, W, 3), dtype=np.uint8), rng.integers(0, 65536, (H, W)).astype(np.uint16)),   # noise incl. d >= 2000 and tiny depths
             (np.full((H, W, 3), 255, np.uint8), np.full((H, W), 65535, np.uint16))]
    step = np.zeros((H, W, 3), np.uint8); step[:, W // 2:] = 255
    ramp = (400 + (np.arange(W)[None, :] + np.arange(H)[:, None]) % 500).astype(np.uint16)
    cases.append((step, ramp))
    for b, d in cases:
        det = _oracle(b, d)
        assert h.match(b, d, 99.0)[0] == 0
        _check_front_end(h, det, W, H)


def test_similarity_maps_bit_exact(vga):
    W, H, b, d, det, ts, h = vga
    assert h.match(b, d, 75.0)[0] == 0
    for t in list(range(0, ts.n_templates, 23)) + [ts.n_templates - 1]:
        assert np.array_equal(h.debug_similarity(t, W, H), det.similarity(t)), "template %d" % t


@pytest.mark.parametrize("thr", [75.0, 62.5, 55.0, 51.0])
def test_match_list_identical(vga, thr):
    W, H, b, d, det, ts, h = vga
    rc, got = h.match(b, d, thr)
    want = det.match(thr)
    assert rc == 0
    assert len(got) == len(want)
    assert np.array_equal(got, want)
    if len(want):
        assert got[0] == want[0]                          # the only match the reference pipeline consumes


def test_large_candidate_count_takes_the_multi_kernel_sort(vga):
    W, H, b, d, det, ts, h = vga
    # low thresholds bring the raw threshold down to the chance level (2*nf): thousands of candidates (2.4k / 9.6k / 39k raw
    # on this scene), which exercises the one-CTA sort with all 1,024 threads (<= 8,192 keys), the multi-kernel sort beyond
    # that, the second D2H chunk and, at 0 %, the overflow path
    for thr in (30.0, 20.0, 10.0, 0.0):
        raw = det.match(thr, canonical=False)
        want = det.match(thr)
        rc, got = h.match(b, d, thr, capacity=1 << 16)
        if len(raw) > (1 << 16):
            assert rc == fb.FL_ERR_CAPACITY
            continue
        assert rc == 0 and len(got) == len(want) and np.array_equal(got, want)
        assert len(raw) > 2048


def test_staged_and_baseline_similarity_kernels_agree(vga):
    W, H, b, d, det, ts, h = vga
    assert h.match(b, d, 60.0)[0] == 0
    assert h.uses_staged()                       # 600 eligible templates at VGA -> shared-memory-staged kernel
    staged = h.match(b, d, 60.0)[1]
    h.force_baseline(True)
    try:
        rc, base = h.match(b, d, 60.0)
        assert rc == 0 and not h.uses_staged()
    finally:
        h.force_baseline(False)
    want = det.match(60.0)
    assert np.array_equal(staged, want) and np.array_equal(base, want)


def test_debug_options_leave_the_match_list_unchanged(vga):
    """fl_debug_option: every developer switch of a handle (include/fealess_b200.h FL_OPT_*) gives the oracle's list - per-wave
    front-end launches, refinement and sort as separate launches, the in-kernel timelines - alone and together."""
    W, H, b, d, det, ts, _ = vga
    want = {thr: det.match(thr) for thr in (75.0, 50.0)}
    for opts in ([fb.FL_OPT_FE_WAVES], [fb.FL_OPT_SPLIT_REFINE], [fb.FL_OPT_TRACE], [fb.FL_OPT_FE_WAVES, fb.FL_OPT_SPLIT_REFINE, fb.FL_OPT_TRACE]):
        h = fb.Handle((5, 8), (0, 1), W, H, max_candidates=1 << 17)
        for o in opts:
            assert h.debug_option(o, 1) == 0
        h.upload_templates(ts)
        n0 = h.launch_count()
        for thr, w in want.items():
            rc, got = h.match(b, d, thr, capacity=1 << 16)
            assert rc == 0 and np.array_equal(got, w), (opts, thr)
        per_frame = (h.launch_count() - n0) / 2
        if fb.FL_OPT_FE_WAVES in opts or fb.FL_OPT_SPLIT_REFINE in opts:
            assert per_frame > 3                                   # the default frame is 3 launches: front end, similarity, refine + sort
        if fb.FL_OPT_TRACE in opts:
            tr = h.staged_trace()
            assert len(tr) > 0 and (tr[:, 4] >= tr[:, 0]).all() and (tr[:, 0] > 0).all()
            if fb.FL_OPT_FE_WAVES not in opts:                     # the job timeline belongs to the single-launch front end
                fe = h.fe_trace()
                assert len(fe) >= 4 and (fe[:, 3] >= fe[:, 2]).all() and (fe[:, 1] > 0).all()
        else:
            with pytest.raises(RuntimeError):
                h.staged_trace()
        # switching an option off again restores the default path
        for o in opts:
            assert h.debug_option(o, 0) == 0
        assert h.match(b, d, 75.0)[0] == 0                         # (re-plans the staged kernel after a change of FL_OPT_TRACE)
        n0 = h.launch_count()
        rc, got = h.match(b, d, 75.0)
        assert rc == 0 and np.array_equal(got, want[75.0]) and h.launch_count() - n0 == 3
        assert h.debug_option(fb.FL_OPT_FE_FORCED_WAVES) == 0
        h.close()
    with pytest.raises(RuntimeError):
        fb.Handle((5, 8), (0, 1), W, H).debug_option(99, 1)


def test_in_grid_dependency_timeout_falls_back_to_wave_launches(vga):
    """The single-launch front end waits in-grid for producer tiles.  If a dependency never arrives (FL_OPT_FE_DEP_TIMEOUT_TEST makes
    the next frame's first one unsatisfiable) the kernel gives up instead of trapping the context, the host re-runs the frame with one
    launch per wave, and the handle stays on wave launches: the caller sees a correct frame and a usable handle."""
    W, H, b, d, det, ts, _ = vga
    h = fb.Handle((5, 8), (0, 1), W, H)
    h.upload_templates(ts)
    want = det.match(75.0)
    rc, got = h.match(b, d, 75.0)
    assert rc == 0 and np.array_equal(got, want) and h.debug_option(fb.FL_OPT_FE_FORCED_WAVES) == 0
    h.debug_option(fb.FL_OPT_FE_DEP_TIMEOUT_TEST, 1)
    rc, got = h.match(b, d, 75.0)                              # times out in the grid (~1 s), re-run in wave mode
    assert rc == 0 and np.array_equal(got, want)
    assert h.debug_option(fb.FL_OPT_FE_FORCED_WAVES) == 1
    b2, d2 = synth.make_frame(W, H, 3)
    det.process(b2, d2)
    rc, got = h.match(b2, d2, 75.0)                            # later frames: wave launches, still right, context alive
    assert rc == 0 and np.array_equal(got, det.match(75.0))
    det.process(b, d)
    # the device-resident entry points take the same path
    import torch
    tb, td = torch.from_numpy(b).cuda(), torch.from_numpy(d).cuda()
    h2 = fb.Handle((5, 8), (0, 1), W, H)
    h2.upload_templates(ts)
    h2.debug_option(fb.FL_OPT_FE_DEP_TIMEOUT_TEST, 1)
    h2.match_device_async(tb.data_ptr(), td.data_ptr(), W, H, 75.0)
    h2.match_wait()
    assert np.array_equal(h2.match_fetch(1 << 14), want) and h2.debug_option(fb.FL_OPT_FE_FORCED_WAVES) == 1
    h.close(); h2.close()


def test_templates_with_features_on_the_box_border_and_ragged_sets(vga):
    """Features at x == width / y == height make similarity() read past the end of a linear-memory row (flat addressing);
    a template larger than the frame has no valid position at all; a set that is not eligible for the staged kernel (64+
    coarse features in total is impossible, so: modalities with different boxes) must fall back to the baseline kernel."""
    W, H, b, d, det0, _, _ = vga
    T = (5, 8)
    q = [det0.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(300, W, H, T, seed=77, quantized=q, planted_fraction=0.2)
    LM = 4
    for t in range(0, 300, 3):                   # push some features onto the far corner of the box at both levels
        for e in range(LM):
            hd = ts.headers[t * LM + e]
            ts.features[hd[5]] = (hd[0], hd[1], ts.features[hd[5], 2])
            ts.features[hd[5] + 1] = (hd[0], 0, ts.features[hd[5] + 1, 2])
    ts.headers[5 * LM + 2, 0] = ts.headers[5 * LM + 3, 0] = 400       # coarsest-level template wider than the 320-px level
    det = F.Detector(T); det.process(b, d); det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    for thr in (70.0, 50.0):
        rc, got = h.match(b, d, thr)
        assert rc == 0 and h.uses_staged() and np.array_equal(got, det.match(thr))
    ts.headers[7 * LM + 3, 0] += 8               # depth modality box differs from the colour one -> not eligible
    det.set_templates(ts); h.upload_templates(ts)
    rc, got = h.match(b, d, 50.0)
    assert rc == 0 and not h.uses_staged() and np.array_equal(got, det.match(50.0))


def test_class_filter(vga):
    W, H, b, d, det, ts, h = vga
    for filt in ([2], [0, 3], [1, 99]):
        rc, got = h.match(b, d, 55.0, class_filter=filt)
        assert rc == 0 and np.array_equal(got, det.match(55.0, class_filter=[c for c in filt if c < 4]))


def test_golden_small_fixture_with_masks(golden_dir):
    z = np.load(os.path.join(golden_dir, "linemod_small.npz"))
    W, H = int(z["W"]), int(z["H"])
    ts = tset_from_npz(z, synth)
    h = fb.Handle(tuple(int(t) for t in z["T"]), (0, 1), 640, 480)      # handle larger than the frame
    h.upload_templates(ts)
    for thr in (75, 55):
        rc, got = h.match(z["bgr"], z["depth"], float(thr))
        assert rc == 0 and np.array_equal(got, z["final_%d" % thr])
    rc, got, q = h.match(z["bgr"], z["depth"], 60.0, masks=[z["mask_0"], z["mask_1"]], want_quantized=True)
    assert rc == 0
    for i in range(4):
        assert np.array_equal(q[i], z["mquantized_%d" % i])
    assert np.array_equal(got, z["mfinal_60"])
    rc, got = h.match(z["bgr"], z["depth"], 55.0, class_filter=[1])
    assert np.array_equal(got, z["final_55_class1"])
    # back to unmasked on the same handle
    rc, got = h.match(z["bgr"], z["depth"], 55.0)
    assert np.array_equal(got, z["final_55"])


def test_golden_vga_hashes(golden_dir, vga):
    W, H, b, d, det, ts, h = vga
    z = np.load(os.path.join(golden_dir, "linemod_vga_hashes.npz"))
    if sha(b) != str(z["bgr_sha"]):
        pytest.skip("numpy RNG stream differs from the fixture's")
    assert h.match(b, d, 75.0)[0] == 0
    assert [sha(h.debug_quantized(l, m, W, H)) for l in range(2) for m in range(2)] == [str(s) for s in z["quantized_sha"]]
    assert [sha(h.debug_lm(l, m, lab, W, H)) for l in range(2) for m in range(2) for lab in range(8)] == [str(s) for s in z["lm_sha"]]


def test_error_codes_mirror_the_reference(vga):
    W, H, b, d, det, ts, h = vga
    b2, d2 = synth.make_frame(632, 480, 0)               # 632 % 5 != 0
    assert h.match(b2, d2, 75.0)[0] == fb.FL_ERR_GEOMETRY
    assert h.match(np.zeros((960, 1280, 3), np.uint8), np.zeros((960, 1280), np.uint16), 75.0)[0] == fb.FL_ERR_SIZE   # exceeds capacity
    rc, got = h.match(b, d, 75.0)                        # the handle still works afterwards
    assert rc == 0 and np.array_equal(got, det.match(75.0))
    bad = synth.make_templates(3, seed=1)
    bad.headers[0, 6] = 64
    h2 = fb.Handle()
    with pytest.raises(fb.FealessError) as e:
        h2.upload_templates(bad)
    assert e.value.rc == fb.FL_ERR_FEATURES
    empty = synth.make_templates(0)
    h2.upload_templates(empty)
    rc, got = h2.match(b, d, 75.0)
    assert rc == 0 and len(got) == 0


def test_three_levels_720p():
    W, H, T = 1280, 720, (5, 8, 5)                       # 1280x720, 640x360, 320x180
    b, d = synth.make_frame(W, H, 7)
    det = _oracle(b, d, T)
    q = [det.quantized(l, m) for l in range(3) for m in range(2)]
    ts = synth.make_templates(200, W, H, T, n_classes=15, seed=5, quantized=q, planted_fraction=0.1, max_size=160)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    h.keep_spread(True)
    rc, got = h.match(b, d, 70.0)
    assert rc == 0
    _check_front_end(h, det, W, H, L=3)
    want = det.match(70.0)
    assert len(want) > 0 and np.array_equal(got, want)


def test_four_levels_1080p():
    """BASELINE config C5's geometry: 1920x1080, four pyramid levels, T = {5,5,5,5} (the only T that divides every level,
    SURVEY.md 8d): 16 front-end jobs in one launch with a three-deep pyrDown chain, the depth job writing three pyramid levels,
    partial tiles at 240x135, three refinement levels inside the fused refinement + sort launch."""
    W, H, T = 1920, 1080, (5, 5, 5, 5)
    b, d = synth.make_frame(W, H, 3)
    det = _oracle(b, d, T)
    q = [det.quantized(l, m) for l in range(4) for m in range(2)]
    ts = synth.make_templates(300, W, H, T, n_classes=5, seed=15, quantized=q, planted_fraction=0.1, max_size=160)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    h.keep_spread(True)
    rc, got = h.match(b, d, 70.0)
    assert rc == 0
    _check_front_end(h, det, W, H, L=4)
    want = det.match(70.0)
    assert len(want) > 0 and np.array_equal(got, want)
    h.close()


@pytest.mark.parametrize("W,H,T", [(240, 160, (5, 8)), (250, 160, (5, 5)), (480, 270, (5, 5)), (80, 60, (5, 5))])
def test_front_end_ragged_sizes(W, H, T):
    """Partial tiles, widths that are not multiples of 4 (unaligned load / store paths), odd level-1 sizes, tiny frames."""
    b, d = synth.make_frame(W, H, 11)
    det = _oracle(b, d, T)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(synth.make_templates(0))
    h.keep_spread(True)
    rc, got, q = h.match(b, d, 80.0, want_quantized=True)
    assert rc == 0 and len(got) == 0
    for i in range(4):
        assert np.array_equal(q[i], det.quantized(i // 2, i % 2)), "quantized %d" % i
    _check_front_end(h, det, W, H)
    h.close()


def test_single_modality_colour_only():
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 9)
    det = F.Detector(T, modality_kind=(0,))
    assert det.process(b, d) == 0
    q = [det.quantized(l, 0) for l in range(2)]
    ts = synth.make_templates(100, W, H, T, n_modalities=1, seed=8, quantized=q, planted_fraction=0.2)
    det.set_templates(ts)
    h = fb.Handle(T, (0,), W, H)
    h.upload_templates(ts)
    rc, got = h.match(b, None, 70.0)
    assert rc == 0 and np.array_equal(got, det.match(70.0)) and len(got) > 0
    assert h.match(None, d, 70.0)[0] == fb.FL_ERR_SIZE    # the colour modality has no source


def _full_size_case(W, H, T, n_templates, n_classes, thresholds, seed=1, planted=0.01, max_candidates=1 << 16):
    """Whole-list parity at a BASELINE.json configuration's real size: the C oracle over EVERY template (OpenMP over templates,
    result-identical to one thread - tests/test_oracle_golden.py), so a missed match among the non-matching templates would show."""
    b, d = synth.make_frame(W, H, 0)
    det = _oracle(b, d, T)
    L = len(T)
    q = [det.quantized(l, m) for l in range(L) for m in range(2)]
    ts = synth.make_templates(n_templates, W, H, T, n_classes=n_classes, seed=seed, quantized=q, planted_fraction=planted)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H, max_candidates=max_candidates)
    h.upload_templates(ts)
    h.keep_spread(True)
    out = []
    for thr in thresholds:
        want = det.match(thr, n_threads=os.cpu_count() or 1)
        rc, got = h.match(b, d, thr, capacity=max_candidates)
        assert rc == 0 and len(got) == len(want) > 0, (thr, rc, len(got), len(want))
        assert np.array_equal(got, want), thr
        out.append(got)
    assert h.uses_staged()
    _check_front_end(h, det, W, H, L=L)
    return b, d, ts, det, h, out


def test_c2_full_oracle_8k_templates():
    """BASELINE config C2 (640x480, 8,000 templates, match only): full lists at thresholds 75 and 60."""
    W, H, T = 640, 480, (5, 8)
    b, d, ts, det, h, (got75, got60) = _full_size_case(W, H, T, 8000, 1, (75.0, 60.0))
    assert len(got60) > len(got75)
    # determinism across calls, and shard invariance: two half-shards reproduce the full result
    assert np.array_equal(h.match(b, d, 75.0)[1], got75)
    from fealess_b200 import sharded
    parts = []
    for r in range(2):
        sh, gids = sharded.shard_template_set(ts, r, 2)
        hs = fb.Handle(T, (0, 1), W, H)
        hs.upload_templates(sh)
        hs.set_template_ids(gids)
        parts.append(hs.match(b, d, 75.0)[1])
        hs.close()
    assert np.array_equal(canonical(np.concatenate(parts)), got75)
    h.close()


def test_c4_full_oracle_720p_15_classes_x_2000():
    """BASELINE config C4 geometry and size: 1280x720, 15 classes x 2,000 templates, T = {5, 8} (1280 % 5 == 720 % 5 == 0,
    640 % 8 == 360 % 8 == 0)."""
    b, d, ts, det, h, (got,) = _full_size_case(1280, 720, (5, 8), 30000, 15, (75.0,), planted=0.004)
    assert len(np.unique(got["class_idx"])) >= 8
    rc, only = h.match(b, d, 75.0, class_filter=[3, 11])
    assert rc == 0 and np.array_equal(only, det.match(75.0, class_filter=[3, 11], n_threads=os.cpu_count() or 1))
    h.close()


def test_c5_full_oracle_1080p_four_levels_32k():
    """BASELINE config C5 geometry and size: 1920x1080, 4 pyramid levels, 32,000 templates, T = {5, 5, 5, 5} (T_l must
    divide both dimensions of level l; {5, 8, ..} is invalid at 960x540)."""
    b, d, ts, det, h, (got,) = _full_size_case(1920, 1080, (5, 5, 5, 5), 32000, 16, (75.0,), planted=0.004)
    h.close()


def test_device_resident_and_sharded_api_single_gpu():
    import torch
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = _oracle(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(400, W, H, T, n_classes=2, seed=31, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    want = det.match(60.0)
    h = fb.Handle(T, (0, 1), W, H, max_candidates=1 << 15)
    h.upload_templates(ts)
    tb = torch.from_numpy(b).cuda()
    td = torch.from_numpy(d.view(np.int16)).cuda()
    torch.cuda.synchronize()
    h.match_device(tb.data_ptr(), td.data_ptr(), W, H, 60.0)
    assert np.array_equal(h.match_fetch(), want)
    from fealess_b200 import sharded
    sm = sharded.ShardedMatcher(h, ts, 0, 1, capacity=4096)
    sm.match_device(tb.data_ptr(), td.data_ptr(), W, H, 60.0)
    assert np.array_equal(sm.fetch(), want)
    assert h.launch_count() > 0


def test_read_linemod_file_then_match(tmp_path):
    """readLinemod (linemod_if.cpp:36-47): a template file in the reference's layout -> device database -> the oracle's matches;
    and the async halves of match_device give the same list as the synchronous call."""
    from fealess_b200 import linemod_io
    b, d = synth.make_frame(640, 480, 0)
    det = _oracle(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(40, n_classes=2, seed=16, quantized=q, planted_fraction=0.25)
    det.set_templates(ts)
    want = det.match(75.0)
    D0 = fb.Detector()
    D0.add_template_set(ts)
    path = str(tmp_path / "linemod_templates.yml")
    linemod_io.write_linemod(D0, path)
    D = linemod_io.read_linemod(path)
    rc, matches = D.match([b, d], 75.0)
    assert rc == 0 and len(matches) == len(want) > 0
    for m, w in zip(matches, want):
        assert (m.x, m.y, np.float32(m.similarity), m.class_id, m.template_id) == (w["x"], w["y"], w["similarity"], "obj%02d" % w["class_idx"], w["template_id"])
    import torch
    h = D._handle
    tb, td = torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()
    h.match_device_async(tb.data_ptr(), td.data_ptr(), 640, 480, 75.0)
    with pytest.raises(fb.FealessError) as e:                  # one frame in flight per handle
        h.match_device_async(tb.data_ptr(), td.data_ptr(), 640, 480, 75.0)
    assert e.value.rc == fb.FL_ERR_STATE
    h.match_wait()
    assert np.array_equal(h.match_fetch(), want)


def test_reference_facing_detector_mirror():
    b, d = synth.make_frame(640, 480, 0)
    det = _oracle(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(60, n_classes=3, seed=6, quantized=q, planted_fraction=0.2)
    det.set_templates(ts)
    want = det.match(75.0)
    D = fb.Detector()
    D.add_template_set(ts)
    quant = []
    rc, matches = D.match([b, d], 75.0, quantized_images=quant)
    assert rc == 0 and len(matches) == len(want) > 0 and len(quant) == 4
    for m, w in zip(matches, want):
        assert (m.x, m.y, np.float32(m.similarity), m.class_id, m.template_id) == (w["x"], w["y"], w["similarity"], "obj%02d" % w["class_idx"], w["template_id"])
    rc, only = D.match([b, d], 75.0, class_ids=["obj01", "nope"])
    assert rc == 0 and all(m.class_id == "obj01" for m in only)
    assert len(only) == len(det.match(75.0, class_filter=[1]))


def test_fused_shard_exchange_in_one_launch_world1():
    """fl_match_shard_exchange_device_async with a world of one: the rank pushes its block into its own exchange buffer, signals and
    waits on itself, and the same launch refines + sorts - the whole peer-memory path on a single GPU, twice (both parities)."""
    import torch
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = _oracle(b, d, T)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(500, W, H, T, n_classes=3, seed=51, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    cap = 2048
    xbuf = torch.zeros(int(fb.lib().fl_exchange_buffer_bytes(1, cap)), dtype=torch.uint8, device="cuda")
    block = torch.zeros((cap + 1) * 5, dtype=torch.int32, device="cuda")
    tb, td = torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()
    for epoch, thr in ((1, 70.0), (2, 60.0), (3, 70.0)):
        h.match_shard_exchange_device_async(tb.data_ptr(), td.data_ptr(), W, H, thr, 0, 1, [xbuf.data_ptr()], cap, block.data_ptr(), epoch)
        h.match_wait()
        want = det.match(thr)
        assert len(want) > 0 and np.array_equal(h.match_fetch(), want), thr


def test_exchange_block_overflow_is_reported():
    """A rank that emits more candidates than its exchange block holds publishes the RAW count, so that the merged list is reported
    as incomplete (FL_ERR_CAPACITY) instead of silently truncated (peer-memory path, world of one)."""
    import torch
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 0)
    det = _oracle(b, d, T)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(500, W, H, T, n_classes=3, seed=51, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    assert len(det.match(60.0, canonical=False)) > 8
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    cap = 8
    xbuf = torch.zeros(int(fb.lib().fl_exchange_buffer_bytes(1, cap)), dtype=torch.uint8, device="cuda")
    block = torch.zeros((cap + 1) * 5, dtype=torch.int32, device="cuda")
    tb, td = torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()
    h.match_shard_exchange_device_async(tb.data_ptr(), td.data_ptr(), W, H, 60.0, 0, 1, [xbuf.data_ptr()], cap, block.data_ptr(), 1)
    h.match_wait()
    with pytest.raises(fb.FealessError) as e:
        h.match_fetch()
    assert e.value.rc == fb.FL_ERR_CAPACITY
    assert len(h.match_fetch(allow_truncated=True)) <= cap
    h.close()


@pytest.mark.parametrize("world", [2, 8])
def test_gather_layout_on_one_gpu(world):
    """The multi-GPU data path with the collective replaced by a concatenation: `world` handles hold the template shards,
    each writes its candidate block [header | records]; the blocks laid out as all_gather_into_tensor would lay them out
    are merged by fl_sort_unique_blocks_device and must equal the single-handle result and the oracle."""
    import torch
    from fealess_b200 import sharded
    W, H, T = 640, 480, (5, 8)
    b, d = synth.make_frame(W, H, 3)
    det = _oracle(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(700, W, H, T, n_classes=3, seed=41, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    want = det.match(65.0)
    assert len(want) > 0
    tb = torch.from_numpy(b).cuda()
    td = torch.from_numpy(d.view(np.int16)).cuda()
    cap = 512
    blocks = torch.zeros(world * sharded.block_ints(cap), dtype=torch.int32, device="cuda")
    handles = []
    for r in range(world):
        sh, gids = sharded.shard_template_set(ts, r, world)
        hs = fb.Handle(T, (0, 1), W, H)
        hs.upload_templates(sh)
        hs.set_template_ids(gids)
        blk = blocks[r * sharded.block_ints(cap):(r + 1) * sharded.block_ints(cap)]
        torch.cuda.synchronize()
        hs.match_shard_device(tb.data_ptr(), td.data_ptr(), W, H, 65.0, sharded.records_view(blk).data_ptr(), cap, blk.data_ptr())
        hs.sync()
        handles.append(hs)
    counts = blocks.view(world, -1)[:, 0].cpu().numpy()
    assert counts.sum() >= len(want) and counts.max() <= cap
    handles[0].sort_unique_blocks_device(blocks.data_ptr(), world, cap)
    got = handles[0].match_fetch()                     # (raises if any per-list count reads as an overflow)
    assert np.array_equal(got, want)
    for hs in handles:
        hs.close()


def test_sanitizer_case_passes_without_the_tool():
    """tools/racecheck_case.py is the pass compute-sanitizer is pointed at (tools/gpu_check.sh); it checks itself against the oracle,
    so it also has to pass on its own."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "racecheck_case.py")], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0 and "RACECHECK CASE PASS" in r.stdout, r.stdout[-2000:]

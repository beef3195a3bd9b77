"""GPU (-m gpu): fl_pipe - several frames in flight on one GPU (include/fealess_b200.h).  Every frame's list has to equal the
oracle's Detector::match list for THAT frame, in submission order, whatever the depth; the state errors have to be reported."""
import numpy as np
import pytest

import fealess_b200 as fb
import fl_oracle_py as F
from fealess_b200 import synth

pytestmark = pytest.mark.gpu
W, H, T = 640, 480, (5, 8)


@pytest.fixture(scope="module")
def stream_case():
    frames = [synth.make_frame(W, H, i) for i in range(5)]
    det = F.Detector(T)
    assert det.process(*frames[0]) == 0
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(700, W, H, T, n_classes=3, seed=5, quantized=q, planted_fraction=0.05)
    det.set_templates(ts)
    want = []
    for b, d in frames:
        assert det.process(b, d) == 0
        want.append(det.match(70.0))
    assert sum(len(w) for w in want) > 20 and len({len(w) for w in want}) > 1
    return frames, ts, want


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_stream_of_frames_in_order(stream_case, depth):
    frames, ts, want = stream_case
    p = fb.Pipe(depth, T, (0, 1), W, H)
    p.upload_templates(ts)
    n = 23
    got = []
    for i in range(n):
        if p.in_flight() == depth:
            got.append(p.collect())
        p.submit(*frames[i % 5], 70.0)
    assert p.in_flight() == min(depth, n)
    while p.in_flight():
        got.append(p.collect())
    assert len(got) == n
    for i, (rc, m) in enumerate(got):
        assert rc == 0 and np.array_equal(m, want[i % 5]), i
    p.close()


def test_state_errors_and_batch(stream_case):
    import torch
    frames, ts, want = stream_case
    p = fb.Pipe(2, T, (0, 1), W, H)
    with pytest.raises(fb.FealessError) as e:
        p.collect()                                              # nothing in flight
    assert e.value.rc == fb.FL_ERR_STATE
    p.upload_templates(ts)
    p.submit(*frames[0], 70.0)
    p.submit(*frames[1], 70.0)
    with pytest.raises(fb.FealessError) as e:
        p.submit(*frames[2], 70.0)                               # depth frames in flight already
    assert e.value.rc == fb.FL_ERR_STATE
    with pytest.raises(fb.FealessError) as e:
        p.upload_templates(ts)                                   # not while frames are in flight
    assert e.value.rc == fb.FL_ERR_STATE
    assert np.array_equal(p.collect()[1], want[0]) and np.array_equal(p.collect()[1], want[1])
    # a frame with a bad geometry is refused at submit and does not occupy a slot
    bb, dd = synth.make_frame(640, 488, 0)
    p2 = fb.Pipe(2, T, (0, 1), 640, 488)
    with pytest.raises(fb.FealessError) as e:
        p2.submit(bb, dd, 70.0)
    assert e.value.rc == fb.FL_ERR_GEOMETRY and p2.in_flight() == 0
    p2.close()
    # batch entry: 11 frames through 2 slots, a class filter, and a capacity that truncates some lists
    order = [i % 5 for i in range(11)]
    rc, lists = p.match_batch([frames[i] for i in order], 70.0)
    assert rc == 0 and all(np.array_equal(lists[k], want[i]) for k, i in enumerate(order))
    rc, lists = p.match_batch([frames[i] for i in order], 70.0, class_filter=[1])
    assert rc == 0 and all(np.array_equal(lists[k], want[i][want[i]["class_idx"] == 1]) for k, i in enumerate(order))
    cap = min(len(w) for w in want) + 1
    rc, lists = p.match_batch([frames[i] for i in order], 70.0, capacity_per_frame=cap)
    assert rc == fb.FL_ERR_CAPACITY and all(np.array_equal(lists[k], want[i][:cap]) for k, i in enumerate(order))
    assert p.match_batch([], 70.0) == (0, [])
    # page-locked host frames (read by DMA while in flight) and device-resident frames
    pinned = [(torch.from_numpy(b).pin_memory().numpy(), torch.from_numpy(d.view(np.int16)).pin_memory().numpy().view(np.uint16)) for b, d in frames]
    dev = [(torch.from_numpy(b).cuda(), torch.from_numpy(d.view(np.int16)).cuda()) for b, d in frames]
    for k in range(8):
        if p.in_flight() == 2:
            i = k - 2
            assert np.array_equal(p.collect()[1], want[i % 5])
        if k % 2:
            p.submit(*pinned[k % 5], 70.0)
        else:
            p.submit_device(dev[k % 5][0].data_ptr(), dev[k % 5][1].data_ptr(), W, H, 70.0)
    assert np.array_equal(p.collect()[1], want[6 % 5]) and np.array_equal(p.collect()[1], want[7 % 5])
    p.close()


def test_blocking_wait_gives_the_same_lists(stream_case):
    """fl_set_blocking_wait: the host thread sleeps on a blocking event instead of spinning; nothing else changes."""
    frames, ts, want = stream_case
    h = fb.Handle(T, (0, 1), W, H)
    h.upload_templates(ts)
    h.set_blocking_wait(True)
    for i in (0, 3, 1):
        rc, got = h.match(*frames[i], 70.0)
        assert rc == 0 and np.array_equal(got, want[i])
    h.set_blocking_wait(False)
    rc, got = h.match(*frames[2], 70.0)
    assert rc == 0 and np.array_equal(got, want[2])
    h.close()


def test_template_sharded_pipes_on_one_gpu(stream_case):
    """fl_pipe_set_exchange: two pipes = the two ranks of a template-sharded detector, both on this GPU, exchange buffers in plain device
    memory; driven in lock-step from one thread, every frame's merged list on both ranks equals the single detector's."""
    import torch
    from fealess_b200 import sharded
    frames, ts, want = stream_case
    world, depth, cap = 2, 3, 1024
    nbytes = int(fb.lib().fl_exchange_buffer_bytes(world, cap))
    xbuf = [[torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)] for _ in range(depth)]     # [slot][rank]
    blocks = [[torch.zeros(sharded.block_ints(cap), dtype=torch.int32, device="cuda") for _ in range(depth)] for _ in range(world)]
    pipes = []
    for r in range(world):
        p = fb.Pipe(depth, T, (0, 1), W, H)
        shard, gids = sharded.shard_template_set(ts, r, world)
        p.upload_templates(shard)
        p.set_template_ids(gids)
        p.set_exchange(r, world, cap, [[xbuf[i][q].data_ptr() for q in range(world)] for i in range(depth)], [b.data_ptr() for b in blocks[r]])
        pipes.append(p)
    torch.cuda.synchronize()
    n, got = 14, [[], []]
    for i in range(n):
        if pipes[0].in_flight() == depth:
            for r in range(world):
                got[r].append(pipes[r].collect())
        for r in range(world):                                   # both ranks submit frame i before either is waited for
            pipes[r].submit(*frames[i % 5], 70.0)
    while pipes[0].in_flight():
        for r in range(world):
            got[r].append(pipes[r].collect())
    for r in range(world):
        assert len(got[r]) == n
        for i, (rc, m) in enumerate(got[r]):
            assert rc == 0 and np.array_equal(m, want[i % 5]), (r, i)
    for p in pipes:
        p.close()

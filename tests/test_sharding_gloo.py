"""CPU, world_size 2, gloo: the N > 1 host logic of template-sharded matching (shard assignment, global template ids,
the single all-gather of [header | records] blocks, merge) gives exactly the single-process result.  The per-shard
candidate lists come from the C oracle here (the GPU path is exercised by tests/test_gpu_match.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fl_oracle_py as F
from fealess_b200 import MATCH_DTYPE, sharded, synth
from helpers import canonical


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    b, d = synth.make_frame(320, 160, seed=0xA11CE)
    det = F.Detector((5, 8))
    det.process(b, d)
    q = [det.quantized(l, m) for l in range(2) for m in range(2)]
    ts = synth.make_templates(41, 320, 160, (5, 8), n_classes=3, seed=11, quantized=q, planted_fraction=0.3, max_size=96, min_size=24)
    return b, d, ts


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, d, ts = _scene()
        shard, gids = sharded.shard_template_set(ts, rank, world)
        det = F.Detector((5, 8))
        det.set_templates(shard)
        det.process(b, d)
        raw = det.match(55.0, canonical=False)
        # shard-local template ids -> global ids (what Handle.set_template_ids does on the device)
        first = {}
        for t in range(shard.n_templates):
            first.setdefault(int(shard.class_of[t]), t)
        for r in raw:
            r["template_id"] = gids[first[int(r["class_idx"])] + r["template_id"]]
        cap = 512
        block = torch.zeros(sharded.block_ints(cap), dtype=torch.int32)
        recs = sharded.records_view(block)
        recs[: len(raw) * 5] = torch.from_numpy(np.ascontiguousarray(raw).view(np.int32).reshape(-1).copy())
        sharded.pack_block(block, recs, len(raw))
        g = sharded.gather_blocks(block, world)
        counts, _ = sharded.split_blocks(g)
        merged = sharded.merge_on_host(g.numpy())
        np.save(os.path.join(out_dir, "merged_%d.npy" % rank), canonical(merged))
        np.save(os.path.join(out_dir, "counts_%d.npy" % rank), counts.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_match_equals_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b, d, ts = _scene()
    det = F.Detector((5, 8))
    det.set_templates(ts)
    det.process(b, d)
    want = det.match(55.0)
    raw_all = det.match(55.0, canonical=False)
    m0 = np.load(tmp_path / "merged_0.npy")
    m1 = np.load(tmp_path / "merged_1.npy")
    assert len(want) > 5
    assert np.array_equal(m0, want) and np.array_equal(m1, want)      # every rank ends with the identical final list
    c = np.load(tmp_path / "counts_0.npy")
    assert c.sum() == len(raw_all) and (c > 0).all()


def test_shard_assignment_is_a_partition():
    for n, w in [(41, 2), (8000, 8), (7, 8), (0, 4)]:
        parts = [sharded.shard_indices(n, r, w) for r in range(w)]
        allidx = np.sort(np.concatenate(parts)) if n else np.zeros(0, np.int64)
        assert np.array_equal(allidx, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    ts = synth.make_templates(10, n_classes=2, seed=1)
    sh, gids = sharded.shard_template_set(ts, 1, 3)        # templates 1, 4, 7
    assert sh.n_templates == 3 and list(sh.class_of) == [0, 0, 1] and list(gids) == [1, 4, 2]
    h, f = sh.template(2, 1, 0)
    h0, f0 = ts.template(7, 1, 0)
    assert np.array_equal(f, f0) and list(h[:5]) == list(h0[:5])


def _fake_refine(indices):
    """A stand-in for the batched ICP launch: a record that depends only on the hypothesis index."""
    from fealess_b200 import ICP_RESULT_DTYPE
    out = np.zeros(len(indices), ICP_RESULT_DTYPE)
    for j, k in enumerate(indices):
        out["R"][j] = np.arange(9, dtype=np.float32) + 10 * k
        out["T"][j] = (k, 2 * k, (k * 37) % 11)
        out["dist_mean"][j] = 0.25 + (k * 7) % 5
        out["iterations"][j] = k % 10
        out["n_points"][j] = 1000 + (k * 131) % 400
    return out


def _refine_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for n_hyp in (0, 1, 2, 5, 16):
            calls = []

            def fn(idx):
                calls.append(list(idx))
                return _fake_refine(idx)
            merged = sharded.refine_sharded(n_hyp, fn, rank, world)
            assert calls == ([list(range(rank, n_hyp, world))] if len(range(rank, n_hyp, world)) else [])   # empty shares launch nothing
            np.save(os.path.join(out_dir, "refined_%d_%d.npy" % (n_hyp, rank)), merged)
    finally:
        dist.destroy_process_group()


def test_two_rank_hypothesis_sharded_refinement_equals_single_process(tmp_path):
    """Top-K hypotheses dealt k % world, ONE all-gather of the pose records, every rank ends with the complete list in hypothesis
    order (the order nonMaximumSuppression depends on)."""
    import fl_oracle_py as Fo
    world = 2
    mp.spawn(_refine_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for n_hyp in (0, 1, 2, 5, 16):
        want = sharded.refine_sharded(n_hyp, _fake_refine, 0, 1)                          # world of one: no collective
        assert len(want) == n_hyp
        for rank in range(world):
            got = np.load(tmp_path / ("refined_%d_%d.npy" % (n_hyp, rank)))
            assert got.tobytes() == want.tobytes()
    # the greedy NMS over the gathered list (hypothesis order) is therefore the same on every rank and for every world size
    full = _fake_refine(np.arange(16))
    a = Fo.nms(full["T"], full["n_points"].astype(np.int32), full["dist_mean"], 12.0)
    b = Fo.nms(np.load(tmp_path / "refined_16_1.npy")["T"], full["n_points"].astype(np.int32), full["dist_mean"], 12.0)
    assert np.array_equal(a, b) and 0 < len(a) < 16
    assert [list(sharded.hypothesis_share(5, r, 2)) for r in range(2)] == [[0, 2, 4], [1, 3]]

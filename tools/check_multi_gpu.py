"""Multi-GPU parity check (run under torchrun on N GPUs of one box): the template-sharded matcher, with both exchange
modes, must return exactly the oracle's match list on every rank, frame after frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist
import fealess_b200 as fb
from fealess_b200 import sharded, synth
import fl_oracle_py as F

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(6)]
det = F.Detector(T); det.process(*frames[0])
q = [det.quantized(l, m) for l in range(2) for m in range(2)]
ts = synth.make_templates(1200, W, H, T, n_classes=3, seed=51, quantized=q, planted_fraction=0.05)
det.set_templates(ts)
want = []
for b, d in frames:
    det.process(b, d); want.append(det.match(70.0))
ok_all = True
for mode in ("p2p", "nccl"):
    h = fb.Handle(T, (0, 1), W, H, device=local)
    try:
        sm = sharded.ShardedMatcher(h, ts, rank, world, capacity=1024, device=dev, exchange=mode)
    except Exception as e:
        print("rank", rank, "mode", mode, "unavailable:", repr(e)[:300], flush=True)
        ok_all = False
        continue
    d_frames = [(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for b, d in frames]
    torch.cuda.synchronize(); dist.barrier()
    ok = True
    for rep in range(3):
        for i, (tb, td) in enumerate(d_frames):
            sm.match_device(tb.data_ptr(), td.data_ptr(), W, H, 70.0)
            got = sm.fetch()
            same = len(got) == len(want[i]) and bool((got == want[i]).all())
            ok &= same
            if not same:
                print("rank", rank, "mode", mode, "frame", i, "MISMATCH", len(got), len(want[i]), flush=True)
    print("rank", rank, "mode", mode, "(effective %s)" % sm.exchange, "PASS" if ok else "FAIL", "matches/frame", [len(w) for w in want], flush=True)
    ok_all &= ok
    sm.close()
# several frames in flight per rank (sharded.ShardedPipe): device frames and host frames (page-locked and pageable), both exchanges
pinned = [(torch.from_numpy(b).pin_memory().numpy(), torch.from_numpy(d.view(np.int16)).pin_memory().numpy().view(np.uint16)) for b, d in frames]
for mode in ("p2p", "nccl"):
    try:
        pipe = sharded.ShardedPipe(lambda: fb.Handle(T, (0, 1), W, H, device=local), ts, rank, world, depth=3, capacity=1024, device=dev, exchange=mode)
    except Exception as e:
        print("rank", rank, "pipe mode", mode, "unavailable:", repr(e)[:300], flush=True)
        ok_all = False
        continue
    torch.cuda.synchronize(); dist.barrier()
    ok, got = True, []
    n = 20
    for i in range(n):
        if pipe.in_flight() == 3:
            got.append(pipe.collect().fetch().copy())
        k = i % 6
        if i % 3 == 0:
            pipe.submit_device(d_frames[k][0].data_ptr(), d_frames[k][1].data_ptr(), W, H, 70.0)
        elif i % 3 == 1:
            pipe.submit_host(pinned[k][0], pinned[k][1], 70.0)
        else:
            pipe.submit_host(frames[k][0], frames[k][1], 70.0)
    while pipe.in_flight():
        got.append(pipe.collect().fetch().copy())
    for i in range(n):
        same = len(got[i]) == len(want[i % 6]) and bool((got[i] == want[i % 6]).all())
        ok &= same
        if not same:
            print("rank", rank, "pipe mode", mode, "frame", i, "MISMATCH", len(got[i]), len(want[i % 6]), flush=True)
    print("rank", rank, "pipe mode", mode, "(effective %s)" % pipe.exchange, "PASS" if ok else "FAIL", flush=True)
    ok_all &= ok
    pipe.close()
t = torch.tensor([1 if ok_all else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI-GPU PARITY", "PASS" if int(t.item()) else "FAIL", flush=True)
code = 0 if int(t.item()) else 1
torch.cuda.synchronize()
dist.destroy_process_group()
sys.stdout.flush(); sys.stderr.flush()
os._exit(code)                                                  # (no interpreter teardown: it frees torch tensors in arbitrary order)

import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np, torch, torch.distributed as dist
import fealess_b200 as fb
from fealess_b200 import sharded, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, H, T = 640, 480, (5, 8)
frames = [synth.make_frame(W, H, i) for i in range(4)]
h = fb.Handle(T, (0, 1), W, H, device=local)
h.upload_templates(synth.make_templates(0))
rc, _, q = h.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
ts = synth.make_templates(8000 * world, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
for mode in ("p2p", "nccl"):
    hh = fb.Handle(T, (0, 1), W, H, device=local)
    sm = sharded.ShardedMatcher(hh, ts, rank, world, capacity=2048, device=dev, exchange=mode)
    d_frames = [(torch.from_numpy(b).to(dev), torch.from_numpy(d.view(np.int16)).to(dev)) for b, d in frames]
    torch.cuda.synchronize(); dist.barrier()
    stamps = []
    for i in range(20):
        sm.match_device(d_frames[i % 4][0].data_ptr(), d_frames[i % 4][1].data_ptr(), W, H, 75.0)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for i in range(200):
        sm.match_device(d_frames[i % 4][0].data_ptr(), d_frames[i % 4][1].data_ptr(), W, H, 75.0)
        c = np.zeros(16, np.int32)
        fb.lib().fl_debug_get(hh._h, 5, 0, 0, 0, C.c_void_p(c.ctypes.data), C.c_size_t(64))
        stamps.append(c[8:11].copy())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 200 * 1e6
    st = np.array(stamps)
    print("rank", rank, mode, sm.exchange, "wall us/frame %.1f" % dt, "exchange ns: push %.0f wait %.0f sort %.0f (medians)" % tuple(np.median(st, 0)), "wait p90 %.0f" % np.percentile(st[:, 1], 90), flush=True)
    hh.close()
dist.destroy_process_group()

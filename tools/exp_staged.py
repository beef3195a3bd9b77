"""Developer experiment (GPU box): staged global-similarity kernel variants (cluster size, phase size) at the bench workload:
stage times by CUDA events, in-kernel timeline (FL_TRACE), and parity of the match list against the baseline kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FL_TRACE"] = "1"
import ctypes as C
import numpy as np
import fealess_b200 as fb
from fealess_b200 import synth

W, H, T = 640, 480, (5, 8)
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 8000


def trace(h):
    buf = np.zeros(8192 * 8, np.uint64)
    n = fb.lib().fl_debug_get(h._h, 4, 0, 0, 0, C.c_void_p(buf.ctypes.data), C.c_size_t(buf.nbytes))
    if n <= 0:
        return None
    per_warp = buf[n * 8 + 8:n * 8 + 8 + n * 128].reshape(n, 32, 4).astype(np.int64)      # [cta][warp]{loop end, warp end}
    global PER_WARP
    PER_WARP = per_warp
    return buf[:n * 8].reshape(n, 8).astype(np.int64), buf[n * 8:n * 8 + 2].astype(np.int64)


def main():
    frames = [synth.make_frame(W, H, i) for i in range(4)]
    h0 = fb.Handle(T, (0, 1), W, H)
    h0.upload_templates(synth.make_templates(0))
    rc, _, q = h0.match(frames[0][0], frames[0][1], 75.0, want_quantized=True)
    ts = synth.make_templates(NT, W, H, T, seed=1, quantized=q, planted_fraction=0.01)
    h0.upload_templates(ts)
    h0.force_baseline(True)
    ref = [h0.match(b, d, 75.0)[1] for b, d in frames]
    print("baseline matches per frame", [len(r) for r in ref])
    for cl, kb in [(1, 2)]:
        os.environ["FL_SS_CLUSTER"] = str(cl)
        os.environ["FL_SS_NBUF"] = str(kb)
        h = fb.Handle(T, (0, 1), W, H)
        h.upload_templates(ts)
        h.profile(True)
        ok = True
        st = np.zeros(4)
        n = 0
        for it in range(24):
            b, d = frames[it % 4]
            rc, m = h.match(b, d, 75.0)
            ok &= rc == 0 and len(m) == len(ref[it % 4]) and bool((m == ref[it % 4]).all())
            if it >= 4:
                st += h.last_stage_ms(); n += 1
        st /= n
        tr = trace(h)
        msg = "cluster %d nbuf %3d staged %d parity %s | stage us: fe %.1f sim %.1f refine %.1f sort %.1f" % (
            cl, kb, h.uses_staged(), ok, *(1e3 * st))
        if tr is not None:
            tr, stamps = tr
            msg += " | before-gap %.1f after-gap %.1f stamp-to-stamp %.1f last-warp-end-after-warp0 %.1f producer-done %.1f" % (
                (tr[:, 0].min() - stamps[0]) / 1e3, (stamps[1] - tr[:, 6].max()) / 1e3, (stamps[1] - stamps[0]) / 1e3,
                (tr[:, 6] - tr[:, 4]).max() / 1e3, (tr[:, 7] - tr[:, 0]).mean() / 1e3)
            t0 = tr[:, 0].min()
            span = tr[:, 4].max() - t0
            msg += " | trace us: span %.1f start-skew %.1f init %.1f first-data %.1f loop[min %.1f mean %.1f max %.1f] tail %.1f" % (
                span / 1e3, (tr[:, 0].max() - t0) / 1e3, (tr[:, 1] - tr[:, 0]).mean() / 1e3, (tr[:, 2] - tr[:, 1]).mean() / 1e3,
                (tr[:, 3] - tr[:, 2]).min() / 1e3, (tr[:, 3] - tr[:, 2]).mean() / 1e3, (tr[:, 3] - tr[:, 2]).max() / 1e3,
                (tr[:, 4] - tr[:, 3]).mean() / 1e3)
        print(msg, flush=True)
        if tr is not None:
            pw = PER_WARP
            live = pw[:, :, 0] > 0
            emis = (pw[:, :, 1] - pw[:, :, 0])[live] / 1e3
            t0 = tr[:, 0].min()
            loop_end = (pw[:, :, 0][live] - t0) / 1e3
            warp_end = (pw[:, :, 1][live] - t0) / 1e3
            d = np.where(live, pw[:, :, 1] - pw[:, :, 0], 0)
            for idx in np.argsort(d.ravel())[::-1][:6]:
                c, w = divmod(int(idx), 32)
                print("  slow warp: cta %d warp %d loop end %.1f us emission %.2f us | loop end -> before atomic %.2f, atomic %.2f, after atomic -> end %.2f"
                      % (c, w, (pw[c, w, 0] - t0) / 1e3, d[c, w] / 1e3, (pw[c, w, 2] - pw[c, w, 0]) / 1e3, (pw[c, w, 3] - pw[c, w, 2]) / 1e3, (pw[c, w, 1] - pw[c, w, 3]) / 1e3))
            print("per-warp: loop end us p50 %.1f p99 %.1f max %.1f | warp end us p50 %.1f p99 %.1f max %.1f | emission us p50 %.2f p90 %.2f p99 %.2f max %.2f, warps > 2 us: %d of %d"
                  % (np.percentile(loop_end, 50), np.percentile(loop_end, 99), loop_end.max(), np.percentile(warp_end, 50), np.percentile(warp_end, 99), warp_end.max(),
                     np.percentile(emis, 50), np.percentile(emis, 90), np.percentile(emis, 99), emis.max(), int((emis > 2).sum()), emis.size))
        h.close()


if __name__ == "__main__":
    main()

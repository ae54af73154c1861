"""Phase timeline of the single-step lean kernel under full load (library built with -DLEAN_TRACE, see
tools/lean_variants.sh): %globaltimer stamps of lane 0 of every warp, launches back to back.
  0 entry | 1 tables written | 2 after griddepcontrol.wait | 3 slices arrived (mbarrier) | 4 setup done |
  5 moves + bounding-box tests | 6 reward / flags | 7 rows stored | 8 state written back
Prints, per stamp, the 10 / 50 / 90 % points over the warps of one launch in ns after the launch's first entry, and how
far the next launch's first entry / first post-wait stamp lies behind this launch's last exit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gym_ballenv_b200 import BallVecEnv
T = 48
wl = os.environ.get("WL", "c3")
spec = bench.workload_spec(wl)
os.environ["BALLENV_NO_ROLLOUT"] = "1"
for n in [int(x) for x in os.environ.get("NS", "65536").split(",")]:
    lanes = 2
    warps = (n + 63) // 64 * (64 * lanes // 32)
    buf = torch.zeros(64 * warps * 16, dtype=torch.int64, device="cuda:0")
    os.environ["BALLENV_TRACE_PTR"] = str(buf.data_ptr())
    env = BallVecEnv(n, window=spec["window"], config=bench.env_config(spec), seed=0, device="cuda:0")
    env.reset()
    a = torch.randint(0, 9, (T, n), device="cuda:0")
    out = env.alloc_rollout(T, keep_all_obs=True)
    for _ in range(int(os.environ.get("WARM", "30"))):   # into the steady state: episodes end at every step
        env.step_many(a, keep_all_obs=True, out=out)
    torch.cuda.synchronize()
    base = (int(os.environ.get("WARM", "30")) * T) % 64
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step_many(a, keep_all_obs=True, out=out)
    e1.record()
    torch.cuda.synchronize()
    print("%s n=%d: %.2f us per launch (traced build)" % (wl, n, e0.elapsed_time(e1) * 1e3 / T))
    tr = buf.cpu().numpy().reshape(64, warps, 16)
    # launch numbers continue from the warm-up call: rows (T + i) % 64 hold launch i of the timed call
    order = [(base + i) % 64 for i in range(T)]
    mid = order[20:28]
    for k in range(9):
        rel = [tr[L, :, k] - tr[L, :, 0].min() for L in mid]
        rel = np.stack(rel).astype(np.float64)
        print("  stamp %d: p10 %7.0f  p50 %7.0f  p90 %7.0f  max %7.0f ns" %
              (k, np.percentile(rel, 10), np.percentile(rel, 50), np.percentile(rel, 90), rel.max()))
    per = [tr[order[i + 1], :, 0].min() - tr[order[i], :, 0].min() for i in range(20, 28)]
    gap0 = [tr[order[i + 1], :, 0].min() - tr[order[i], :, 8].max() for i in range(20, 28)]
    gap2 = [tr[order[i + 1], :, 2].min() - tr[order[i], :, 8].max() for i in range(20, 28)]
    print("  launch period (first entry to first entry): %s ns" % per)
    print("  next launch's first entry minus this launch's last exit: %s ns" % gap0)
    print("  next launch's first stamp after the dependency wait minus this launch's last exit: %s ns" % gap2)
    dur = np.stack([tr[L, :, 8] - tr[L, :, 0] for L in mid]).astype(np.float64)
    print("  warp lifetime: p10 %.0f p50 %.0f p90 %.0f ns;  blocks per SM (launch %d): %s" %
          (np.percentile(dur, 10), np.percentile(dur, 50), np.percentile(dur, 90), 20,
           np.bincount(np.bincount(tr[mid[0], ::4, 15].astype(np.int64), minlength=148))))
    # the stragglers of one launch: phase durations of the last warps to leave, against the medians
    # order of the stamps in time: 0 1 2 3 9 4 10 11 5 6 12 13 7 8
    seq = [0, 1, 2, 3, 9, 4, 10, 11, 5, 6, 12, 13, 7, 8]
    names = ["tables", "dep-wait", "slices", "dirty-load", "syncwarp", "philox", "moves+scan", "static scan", "reward",
             "reset", "raster", "expand+store", "write-back"]
    for L in mid[3:6]:
        t0 = tr[L, :, 0].min()
        ts = tr[L][:, seq].astype(np.int64)
        ph = np.diff(ts, axis=1)
        med = np.median(ph, axis=0)
        print("  launch row %d: median phase durations (ns): %s" % (L, dict(zip(names, med.astype(int).tolist()))))
        print("  p99 phase durations (ns): %s" % dict(zip(names, np.percentile(ph, 99, axis=0).astype(int).tolist())))
        fl = tr[L, :, 14]
        print("  warps with flags: goal-change %d mixed %d near %d fin %d rescan %d" %
              tuple(int(((fl >> b) & 1).sum()) for b in range(5)))
        for b, nm in ((0, "goal-change"), (2, "near"), (3, "fin")):
            m = ((fl >> b) & 1) == 1
            if m.any() and (~m).any():
                print("    %-11s: median phases with %s  /  without %s" %
                      (nm, np.median(ph[m], axis=0).astype(int).tolist(), np.median(ph[~m], axis=0).astype(int).tolist()))
        last = np.argsort(tr[L, :, 8])[-10:]
        for w in last:
            print("    warp %5d sm %3d flags %2d entry %6d wait-done %6d exit %6d  phases %s" %
                  (w, tr[L, w, 15], fl[w], tr[L, w, 0] - t0, tr[L, w, 2] - t0, tr[L, w, 8] - t0, ph[w].tolist()))
    env.close()

"""Development aid: per-warp cycle accounting of k_solve_df.  Needs a -DB2D_DF_PROFILE build (B2D_LIB=...)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, edyn_b200 as E
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = E.scenes.mixed_pile(side)
w = E.scenes.build_world(scene); w.step(150); w.sync()
buf = np.zeros(4096, np.uint8)
OFF = int(os.environ.get("DBG_OFF", "1200"))     # offsetof(Counters, dbg)
def dbg():
    w.l.b2d_debug_counters(w.h, buf.ctypes.data_as(C.c_void_p), C.c_uint32(4096))
    return buf.view(np.uint64)[OFF // 8: OFF // 8 + 16].copy()
a = dbg(); w.step(10); w.sync(); b = dbg()
d = (b - a).astype(np.float64) / 10
nw = d[6]
print("warps %.0f  chunk passes/warp %.1f" % (nw, d[4] / nw))
print("per warp: total %.0f kcyc, stalled %.0f kcyc (%.0f iterations), progress %.0f kcyc (%.0f iterations)" % (d[5]/nw/1e3, d[0]/nw/1e3, d[2]/nw, d[1]/nw/1e3, d[3]/nw))
print("per chunk pass: %.0f cyc total, stalled %.0f cyc in %.2f iterations (%.0f cyc each), progress %.0f cyc in %.2f iterations (%.0f cyc each)" % (
    d[5]/d[4], d[0]/d[4], d[2]/d[4], d[0]/max(d[2],1), d[1]/d[4], d[3]/d[4], d[1]/max(d[3],1)))

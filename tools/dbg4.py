"""Development aid: per-warp cycle accounting of k_solve_df (needs a -DB2D_DF_PROFILE build of libb2d.so)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, edyn_b200 as E
side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = E.scenes.mixed_pile(side)
w = E.scenes.build_world(scene); w.step(150); w.sync()
buf = np.zeros(4096, np.uint8)
OFF = int(os.environ.get("DBG_OFF", "0"))
def dbg():
    w.l.b2d_debug_counters(w.h, buf.ctypes.data_as(C.c_void_p), C.c_uint32(4096))
    return buf.view(np.uint64)[OFF // 8: OFF // 8 + 16].copy()
a = dbg(); w.step(10); w.sync(); b = dbg()
d = (b - a).astype(np.float64) / 10
nw = d[6]
print("warps %.0f  chunk-ops/warp %.1f" % (nw, d[4] / nw))
print("per warp: total %.0f kcyc, stalled %.0f kcyc (%.0f iterations), progress %.0f kcyc (%.0f iterations)" % (d[5]/nw/1e3, d[0]/nw/1e3, d[2]/nw, d[1]/nw/1e3, d[3]/nw))
print("per chunk-op: %.0f cyc total, stalled %.0f cyc in %.2f iterations (%.0f cyc each), progress %.0f cyc in %.2f iterations (%.0f cyc each)" % (
    d[5]/d[4], d[0]/d[4], d[2]/d[4], d[0]/max(d[2],1), d[1]/d[4], d[3]/d[4], d[1]/max(d[3],1)))
e = d[8:]
print("per chunk-op cycles: hdr+ticket wait %.0f | first poll %.0f | later polls %.0f (%.2f polls, %.0f each) | row wait %.0f | normal solve %.0f | friction solve %.0f | solves %.2f" % (
    e[0]/d[4], e[1]/d[4], e[2]/d[4], e[7]/d[4], e[2]/max(e[7],1), e[3]/d[4], e[4]/d[4], e[6]/d[4], e[5]/d[4]))
print("per solve execution: row wait %.0f, normal %.0f, friction %.0f (each on half of the chunk-ops)" % (e[3]/e[5], 2*e[4]/e[5], 2*e[6]/e[5]))

"""Development aid: phase-by-phase comparison of the device path against the oracle with verbose output."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import edyn_b200 as E
from oracle import oracle as O

def pairset(p, ordered=False):
    return {tuple(x) if ordered else tuple(sorted(x)) for x in p.tolist()}

def contacts_by_pair(c):
    out = {}
    for k in range(len(c['num'])):
        out[tuple(c['pairs'][k].tolist())] = (int(c['num'][k]), c['pts'][k], c['att'][k], c['lifetime'][k])
    return out

def compare_contacts(g, o, tol=1e-5, verbose=True):
    G, Oc = contacts_by_pair(g), contacts_by_pair(o)
    bad = 0; maxerr = 0.0; npts = 0
    if set(G) != set(Oc):
        print("  manifold key sets differ:", len(set(G) ^ set(Oc))); bad += len(set(G) ^ set(Oc))
    for k in set(G) & set(Oc):
        ng, pg, ag, lg = G[k]; no, po, ao, lo = Oc[k]
        if ng != no:
            bad += 1
            if verbose and bad < 6: print("  npts differ", k, ng, no)
            continue
        npts += ng
        e = np.abs(pg[:ng] - po[:ng]).max() if ng else 0.0
        maxerr = max(maxerr, float(e))
        if e > tol or not np.array_equal(ag[:ng], ao[:ng]) or not np.array_equal(lg[:ng], lo[:ng]):
            bad += 1
            if verbose and bad < 6: print("  point mismatch", k, "err", e, ag[:ng], ao[:ng], lg[:ng], lo[:ng])
    return bad, maxerr, npts

def lockstep(scene, steps, resync=True, verbose=False):
    w = E.scenes.build_world(scene)
    o = O.OracleWorld(vel_iters=scene['settings']['velocity_iterations'], pos_iters=scene['settings']['position_iterations'])
    o.add_bodies(scene['bodies'])
    if scene['hinges']:
        h = scene['hinges']; o.add_hinges(h['a'], h['b'], h['pivot_a'], h['pivot_b'], h['axis_a'], h['axis_b'])
    if scene['exclusions'] is not None: o.add_exclusions(*scene['exclusions'])
    worst = 0
    for s in range(steps):
        w.run_phases(1); o.run_phases(1)
        gp, op = pairset(w.pairs(), True), pairset(o.pairs(), True)
        if gp != op: print(f"step {s}: ORDERED pair sets differ: {len(gp ^ op)} (unordered diff {len(pairset(w.pairs()) ^ pairset(o.pairs()))})")
        w.run_phases(2); o.run_phases(2)
        bad, me, npts = compare_contacts(w.contacts(), o.contacts(), verbose=verbose)
        if bad: print(f"step {s}: narrowphase mismatches {bad} (maxerr {me:.2e}, pts {npts})")
        w.run_phases(4); o.run_phases(4)
        if not np.array_equal(w.islands(), o.islands()): print(f"step {s}: island labels differ", (w.islands() != o.islands()).sum())
        w.run_phases(8)
        hi, pr = w.solver_order(); o.set_order(hi, pr); o.run_phases(8)
        g, c = w.download_state(inv_IW=True), o.state()
        errs = {k: float(np.abs(g[k] - c[k]).max()) for k in ('pos', 'orn', 'linvel', 'angvel', 'aabb', 'inv_IW')}
        worst = max(worst, errs['pos'])
        if verbose or s % 10 == 0 or errs['pos'] > 1e-4:
            st = w.stats()
            print(f"step {s}: errs " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + f" | manifolds {st['manifolds']} pts {st['contact_points']} colors {st['contact_colors']}/{st['hinge_colors']} islands {st['islands']} err {st['error_flags']}")
        if resync: o.set_state(g['pos'], g['orn'], g['linvel'], g['angvel'])
    return worst

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    t = time.time()
    if which in ("all", "hello"): print("== hello"); lockstep(E.scenes.hello_world(), 120)
    if which in ("all", "boxes"): print("== boxes 4^3"); lockstep(E.scenes.boxes_on_plane(4, jitter=0.01), 100)
    if which in ("all", "spheres"): print("== spheres"); lockstep(E.scenes.spheres_in_box(6, 4, 6), 100)
    if which in ("all", "mixed"): print("== mixed 6^3"); lockstep(E.scenes.mixed_pile(6), 150)
    if which in ("all", "chains"): print("== chains"); lockstep(E.scenes.hinge_chains(4, 4), 100)
    print("elapsed", time.time() - t)

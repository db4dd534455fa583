"""Test infrastructure: the oracle behind the host-adapter surface that edyn_b200.dist.ShardedWorld drives, so the
multi-rank host logic (partitioning, bounds exchange, island migration) can run on CPU with a gloo process group.
Only tests import this module; the product path (edyn_b200.World) never does."""
import numpy as np

from oracle import oracle as O


class OracleBackedWorld:
    def __init__(self, scene, device=0, **kw):
        kw.pop("max_bodies", None); kw.pop("max_hinges", None)        # the oracle has no capacities
        st = dict(scene["settings"]); st.update(kw)
        self.o = O.OracleWorld(vel_iters=st["velocity_iterations"], pos_iters=st["position_iterations"])
        self._defs, self._hinges = [], []
        self.hinge_alive = np.zeros(0, bool)
        self.removed = np.zeros(0, bool)
        self.add_bodies(scene["bodies"])
        if scene["hinges"]:
            h = scene["hinges"]
            self.add_hinges(h["a"], h["b"], h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
        self.exclusions = set()
        if scene["exclusions"] is not None:
            self.add_exclusions(*scene["exclusions"])

    @property
    def num_bodies(self):
        return self.o.num_bodies

    def add_bodies(self, soa):
        n = len(soa["kind"])
        first = self.o.add_bodies(soa)
        self._defs.append({k: np.array(v) for k, v in soa.items() if v is not None and len(v) == n})
        self.removed = np.concatenate([self.removed, np.zeros(n, bool)])
        return first

    def body_defs(self, ids):
        if len(self._defs) > 1:
            self._defs = [{k: np.concatenate([d[k] for d in self._defs]) for k in self._defs[0]}]
        return {k: v[np.asarray(ids, np.int64)].copy() for k, v in self._defs[0].items()}

    def add_hinges(self, a, b, pa, pb, xa, xb):
        n = len(a)
        self.o.add_hinges(a, b, pa, pb, xa, xb)
        f = np.float32
        self._hinges.append(dict(a=np.asarray(a, np.uint32), b=np.asarray(b, np.uint32), pivot_a=np.asarray(pa, f).reshape(n, 3),
                                 pivot_b=np.asarray(pb, f).reshape(n, 3), axis_a=np.asarray(xa, f).reshape(n, 3), axis_b=np.asarray(xb, f).reshape(n, 3)))
        self.hinge_alive = np.concatenate([self.hinge_alive, np.ones(n, bool)])

    def add_exclusions(self, a, b):
        a, b = np.asarray(a, np.uint32), np.asarray(b, np.uint32)
        self.o.add_exclusions(a, b)
        self.exclusions |= {(min(x, y), max(x, y)) for x, y in zip(a.tolist(), b.tolist())}

    def hinge_defs(self):
        if len(self._hinges) > 1:
            self._hinges = [{k: np.concatenate([h[k] for h in self._hinges]) for k in self._hinges[0]}]
        return self._hinges[0] if self._hinges else None

    def remove_bodies(self, ids):
        self.o.remove_bodies(ids)
        self.removed[np.asarray(ids, np.int64)] = True
        h = self.hinge_defs()
        if h is not None:
            self.hinge_alive &= ~(self.removed[h["a"]] | self.removed[h["b"]])

    def step(self, n=1):
        self.o.step(n)

    def download_state(self, aabb=True, **kw):
        return self.o.state()

    def islands(self):
        return self.o.islands()

    def contacts(self):
        return self.o.contacts()

    def upload_contacts(self, pairs, num, pts, att, lifetime=None):
        self.o.set_contacts(pairs, num, pts, att, lifetime)

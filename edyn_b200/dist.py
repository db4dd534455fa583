"""Multi-GPU sharding of one scene across the ranks of a torch.distributed group.

Simulation islands are independent for narrowphase -> solve -> integrate (reference docs/Design.md:205-209,
src/edyn/dynamics/solver.cpp:408-428), so the unit of distribution is the island: every rank owns a set of whole
islands plus a replica of the static bodies (multi_island_resident in the reference, comp/island.hpp:39-41).  There is
no collective on the data path.  The only exchange is the broadphase one: each step every rank publishes the bounding
box of the dynamic bodies it owns (6 floats, all_gather); two ranks whose boxes -- inflated by the broadphase margin --
touch could grow a cross-rank contact and therefore a cross-rank island.  That event is detected
(`ShardedWorld.step` returns the offending rank pairs) and resolved by body migration: the islands of the higher rank
that reach into the lower rank's box move there (`ShardedWorld.migrate`), so an island never spans two GPUs.

Everything here is host/plumbing code: numpy for the partitioning, torch.distributed (NCCL on GPUs, gloo in the CPU
tests) for the exchange.  The simulation itself runs in libb2d.so.
"""
import numpy as np

from .rigidbody import DYNAMIC, SHAPE_BOX, SHAPE_CAPSULE, SHAPE_NONE, SHAPE_PLANE, SHAPE_SPHERE

MARGIN = 0.02 * 1.3            # manifold separation threshold, broadphase.hpp:18


def _rotation_matrices(q):
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s = 2.0 / (x * x + y * y + z * z + w * w)
    R = np.empty((len(q), 3, 3), np.float64)
    R[:, 0, 0] = 1 - s * (y * y + z * z); R[:, 0, 1] = s * (x * y - w * z); R[:, 0, 2] = s * (x * z + w * y)
    R[:, 1, 0] = s * (x * y + w * z); R[:, 1, 1] = 1 - s * (x * x + z * z); R[:, 1, 2] = s * (y * z - w * x)
    R[:, 2, 0] = s * (x * z - w * y); R[:, 2, 1] = s * (y * z + w * x); R[:, 2, 2] = 1 - s * (x * x + y * y)
    return R


def host_aabbs(b, idx):
    """AABB half extents (util/aabb_util.cpp:42-88) of bodies `idx`, on the host, for partitioning only."""
    kind, p = b["shape_kind"][idx], b["shape_params"][idx].astype(np.float64)
    R = np.abs(_rotation_matrices(b["orn"][idx].astype(np.float64)))
    half = np.zeros((len(idx), 3))
    s, c, bx = kind == SHAPE_SPHERE, kind == SHAPE_CAPSULE, kind == SHAPE_BOX
    half[s] = p[s, 0:1]
    if c.any():
        ax = p[c, 2].astype(np.int64)
        half[c] = R[c][np.arange(c.sum()), :, ax] * p[c, 1:2] + p[c, 0:1]
    half[bx] = np.einsum("nij,nj->ni", R[bx], p[bx, :3])
    return half


def initial_islands(scene, reach=MARGIN):
    """Union-find over dynamic bodies whose AABBs, inflated by `reach`, intersect, plus all joints: the simulation
    islands the first broadphase would produce (larger `reach` = merge bodies that are merely close).
    Returns a label per body (smallest member id; -1 for non-dynamic bodies)."""
    b = scene["bodies"]
    n = len(b["kind"])
    dyn = np.where((b["kind"] == DYNAMIC))[0]
    parent = np.arange(n)

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    def union(x, y):
        rx, ry = find(x), find(y)
        if rx != ry:
            if rx < ry:
                parent[ry] = rx
            else:
                parent[rx] = ry

    if scene.get("hinges"):
        for x, y in zip(scene["hinges"]["a"].tolist(), scene["hinges"]["b"].tolist()):
            union(x, y)
    shaped = dyn[b["shape_kind"][dyn] != SHAPE_NONE]
    if len(shaped):
        half = host_aabbs(b, shaped)
        pos = b["pos"][shaped].astype(np.float64)
        size = float(2 * half.max() + reach)
        cell = np.floor(pos / size).astype(np.int64)
        cells = {}
        for i in range(len(shaped)):
            cells.setdefault(tuple(cell[i]), []).append(i)
        for key, members in cells.items():
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in (-1, 0, 1):
                        other = cells.get((key[0] + dx, key[1] + dy, key[2] + dz))
                        if other is None:
                            continue
                        oth = np.asarray(other)
                        for i in members:
                            hit = np.all(np.abs(pos[oth] - pos[i]) <= half[oth] + half[i] + reach, axis=1)
                            for j in oth[hit]:
                                if j != i:
                                    union(int(shaped[i]), int(shaped[j]))
    lab = np.full(n, -1, np.int64)
    for i in dyn:
        lab[i] = find(i)
    return lab


def partition(scene, world_size, labels=None):
    """Assign whole islands to ranks: islands ordered by centroid along x, cut into `world_size` runs of roughly equal
    body count.  Static bodies go to every rank.  Returns owner[n] (rank, or -1 = replicated)."""
    b = scene["bodies"]
    lab = initial_islands(scene) if labels is None else labels
    owner = np.full(len(lab), -1, np.int64)
    ids = np.unique(lab[lab >= 0])
    if len(ids) == 0:
        return owner
    cx = np.array([b["pos"][lab == i, 0].mean() for i in ids])
    size = np.array([(lab == i).sum() for i in ids])
    order = np.argsort(cx, kind="stable")
    total, acc, rank = size.sum(), 0, 0
    for k in order:
        # move to the next rank once this one holds its share (never leave a later rank empty if islands remain)
        if acc >= (rank + 1) * total / world_size and rank < world_size - 1:
            rank += 1
        owner[lab == ids[k]] = rank
        acc += size[k]
    return owner


def shard(scene, rank, world_size, owner=None):
    """The sub-scene rank `rank` simulates: its islands + replicated static bodies, hinges and exclusions remapped."""
    b = scene["bodies"]
    owner = partition(scene, world_size) if owner is None else owner
    keep = np.where((owner == rank) | (owner < 0))[0]
    remap = np.full(len(owner), -1, np.int64)
    remap[keep] = np.arange(len(keep))
    out = {k: (v[keep] if v is not None else None) for k, v in b.items()}
    hinges = None
    if scene.get("hinges"):
        h = scene["hinges"]
        sel = (owner[h["a"]] == rank)
        hinges = {k: (remap[v[sel]].astype(np.uint32) if k in ("a", "b") else v[sel]) for k, v in h.items()}
    excl = None
    if scene.get("exclusions") is not None:
        ea, eb = scene["exclusions"]
        sel = (remap[ea] >= 0) & (remap[eb] >= 0)
        excl = (remap[ea[sel]].astype(np.uint32), remap[eb[sel]].astype(np.uint32))
    dyn = int(((owner == rank)).sum())
    return dict(name=f"{scene['name']}@{rank}/{world_size}", bodies=out, hinges=hinges if hinges and len(hinges["a"]) else None,
                exclusions=excl if excl is not None and len(excl[0]) else None, settings=scene["settings"], dynamic=dyn,
                global_ids=keep)


def overlapping_ranks(bounds, margin=MARGIN):
    """bounds: (world_size, 6) array of min/max of each rank's dynamic bodies (NaN rows = rank owns nothing).
    Returns the list of rank pairs whose boxes, inflated by `margin`, intersect."""
    out = []
    n = len(bounds)
    for i in range(n):
        for j in range(i + 1, n):
            a, b = bounds[i], bounds[j]
            if np.isnan(a).any() or np.isnan(b).any():
                continue
            if np.all(a[0:3] - margin <= b[3:6]) and np.all(a[3:6] + margin >= b[0:3]):
                out.append((i, j))
    return out


def boxes_touch(a, b, margin=MARGIN):
    """a: (n, 6) AABBs, b: one (6,) box; closed-interval intersection with `a` inflated by margin (geom.cpp:762-770)."""
    return np.all(a[:, 0:3] - margin <= b[3:6], axis=1) & np.all(a[:, 3:6] + margin >= b[0:3], axis=1)


class ShardedWorld:
    """One rank's share of a scene, the per-step bounds exchange and island migration.

    `dist` is torch.distributed (initialised) or None.  `world_factory(scene, device=..., **kw)` builds the rank's
    world (default: the device world of scenes.build_world; the CPU tests inject an oracle-backed stand-in).

    Migration (mirrors merge_islands' "move into the other island", island_manager.cpp:297-350): when the
    boxes of two ranks come within the broadphase margin, the ranks gather each other's island AABBs and the higher
    rank hands every island of its own that touches an island of the lower rank over to it -- body definitions with their current state, the joints between them and
    their contact manifolds (points, lifetimes and warm-start impulses), collision exclusions -- and destroys its copies.  Bodies are named
    by scene-global ids on the wire; static bodies are replicated, so a manifold against the ground keeps its partner.
    The transfer is one all_gather_object (sizes first, then payloads: NCCL on GPUs, gloo on CPU) and only happens on
    the steps where `overlapping_ranks` is non-empty."""

    def __init__(self, scene, rank, world_size, dist=None, device=0, world_factory=None, owner=None, **kw):
        if world_factory is None:
            from .scenes import build_world as world_factory
        self.rank, self.world_size, self.dist = rank, world_size, dist
        # owner: rank per body (-1 = replicated static), whole islands per rank; default = balanced x-slabs
        self.owner = partition(scene, world_size) if owner is None else np.asarray(owner, np.int64)
        self.local = shard(scene, rank, world_size, self.owner)
        # any rank may end up owning every island: capacity for the whole scene
        kw.setdefault("max_bodies", len(self.owner))
        kw.setdefault("max_hinges", len(scene["hinges"]["a"]) if scene.get("hinges") else 0)
        self.world = world_factory(self.local, device=device, **kw)
        self.global_of_local = [int(g) for g in self.local["global_ids"]]
        self.local_of_global = {g: l for l, g in enumerate(self.global_of_local)}
        self.next_global = len(self.owner)              # ids for bodies created later (none yet): keep unique per scene
        self.dynamic_local = np.where(self.local["bodies"]["kind"] == DYNAMIC)[0]
        self.migrated_in = self.migrated_out = 0

    def local_bounds(self, aabb):
        if len(self.dynamic_local) == 0:
            return np.full(6, np.nan, np.float32)
        a = aabb[self.dynamic_local]
        return np.concatenate([a[:, 0:3].min(axis=0), a[:, 3:6].max(axis=0)]).astype(np.float32)

    def exchange_bounds(self, bounds):
        """all_gather of 6 floats per rank (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
        if self.dist is None or self.world_size == 1:
            return bounds[None, :]
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        mine = torch.from_numpy(bounds.copy()).to(dev)
        out = [torch.empty_like(mine) for _ in range(self.world_size)]
        self.dist.all_gather(out, mine)
        return torch.stack(out).cpu().numpy()

    # ------------------------------------------------------------------ migration
    def island_boxes(self, st):
        """(labels, boxes): the AABB of every island this rank owns (update_island_aabbs, sys/update_aabbs.cpp:106-138)."""
        dyn = self.dynamic_local
        if len(dyn) == 0:
            return np.zeros(0, np.int64), np.zeros((0, 6), np.float32)
        lab = self.world.islands().astype(np.int64)[dyn]
        ids, inv = np.unique(lab, return_inverse=True)
        a = st["aabb"][dyn]
        box = np.empty((len(ids), 6), np.float32)
        for k in range(3):
            mn = np.full(len(ids), np.inf, np.float32); np.minimum.at(mn, inv, a[:, k]); box[:, k] = mn
            mx = np.full(len(ids), -np.inf, np.float32); np.maximum.at(mx, inv, a[:, 3 + k]); box[:, 3 + k] = mx
        return ids, box

    def _select_outgoing(self, pairs, all_boxes, my_labels):
        """Per destination rank: local ids of the islands this rank hands over (it is the higher rank of the pair):
        those whose island AABB, inflated by the broadphase margin, touches an island AABB of the destination."""
        dests = sorted(i for i, j in pairs if j == self.rank)
        if not dests or len(self.dynamic_local) == 0:
            return {}
        ids, mine = my_labels, all_boxes[self.rank]
        lab = self.world.islands().astype(np.int64)
        dyn = self.dynamic_local
        out, taken = {}, np.zeros(len(ids), bool)
        for dst in dests:
            theirs = all_boxes[dst]
            if len(theirs) == 0:
                continue
            hit = np.zeros(len(ids), bool)
            for b in theirs:
                hit |= boxes_touch(mine, b)
            hit &= ~taken
            if hit.any():
                out[dst] = dyn[np.isin(lab[dyn], ids[hit])]
                taken |= hit
        return out

    def _pack(self, ids, st, contacts):
        """Everything rank `dst` needs to continue simulating bodies `ids` (local ids), in scene-global naming."""
        w = self.world
        gol = np.asarray(self.global_of_local, np.int64)
        defs = w.body_defs(ids)
        for k in ("pos", "orn", "linvel", "angvel"):
            defs[k] = st[k][ids].copy()
        msg = dict(gid=gol[ids], defs=defs, hinges=None, contacts=None, exclusions=None)
        sel = np.zeros(w.num_bodies, bool); sel[ids] = True
        ex = [(a, b) for a, b in getattr(w, "exclusions", ()) if sel[a] and sel[b]]       # collision_exclusion among the movers
        if ex:
            ex = np.asarray(ex, np.int64)
            msg["exclusions"] = (gol[ex[:, 0]], gol[ex[:, 1]])
        h = w.hinge_defs()
        if h is not None:
            m = w.hinge_alive & sel[h["a"]] & sel[h["b"]]
            if m.any():
                msg["hinges"] = {k: (gol[v[m]] if k in ("a", "b") else v[m].copy()) for k, v in h.items()}
        if len(contacts["pairs"]):
            pr = contacts["pairs"].astype(np.int64)
            m = sel[pr[:, 0]] | sel[pr[:, 1]]
            if m.any():
                msg["contacts"] = dict(pairs=gol[pr[m]], num=contacts["num"][m].copy(), pts=contacts["pts"][m].copy(),
                                       att=contacts["att"][m].copy(), lifetime=contacts["lifetime"][m].copy())
        return msg

    def _unpack(self, msg):
        w = self.world
        first = w.add_bodies(msg["defs"])
        n = len(msg["gid"])
        for k, g in enumerate(msg["gid"].tolist()):
            self.local_of_global[g] = first + k
        self.global_of_local.extend(int(g) for g in msg["gid"])
        new_ids = np.arange(first, first + n)
        self.dynamic_local = np.concatenate([self.dynamic_local, new_ids[msg["defs"]["kind"] == DYNAMIC]])
        to_local = np.vectorize(self.local_of_global.__getitem__, otypes=[np.uint32])
        if msg["hinges"] is not None:
            h = msg["hinges"]
            w.add_hinges(to_local(h["a"]), to_local(h["b"]), h["pivot_a"], h["pivot_b"], h["axis_a"], h["axis_b"])
        if msg.get("exclusions") is not None:
            w.add_exclusions(to_local(msg["exclusions"][0]), to_local(msg["exclusions"][1]))
        return n, (None if msg["contacts"] is None else dict(msg["contacts"], pairs=to_local(msg["contacts"]["pairs"])))

    def migrate(self, pairs, bounds, st):
        """Collective: every rank of the group must call it on the same step (they all see the same `pairs`)."""
        # rank boxes are coarse (two ranks' regions may interleave): decide on island boxes, which every rank gathers
        my_labels, my_boxes = self.island_boxes(st)
        all_boxes = [None] * self.world_size
        self.dist.all_gather_object(all_boxes, my_boxes)
        out = self._select_outgoing(pairs, all_boxes, my_labels)
        outbox = {}
        if out:
            contacts = self.world.contacts()
            for dst, ids in out.items():
                outbox[dst] = self._pack(ids, st, contacts)
            gone = np.concatenate(list(out.values()))
            self.world.remove_bodies(gone)
            self.dynamic_local = np.setdiff1d(self.dynamic_local, gone)
            self.migrated_out += len(gone)
        boxes = [None] * self.world_size
        self.dist.all_gather_object(boxes, outbox)
        incoming = [boxes[src][self.rank] for src in range(self.world_size) if boxes[src] and self.rank in boxes[src]]
        if not incoming:
            return 0
        mine = self.world.contacts()
        parts = [mine]
        got = 0
        for msg in incoming:
            n, c = self._unpack(msg)
            got += n
            if c is not None:
                parts.append(c)
        if len(parts) > 1:
            self.world.upload_contacts(*[np.concatenate([p[k] for p in parts]) for k in ("pairs", "num", "pts", "att", "lifetime")])
        self.migrated_in += got
        return got

    def step(self, n=1, check=True, migrate=True):
        """n fixed steps on this rank's islands, then the cross-rank AABB exchange.  Returns the rank pairs whose
        island groups came within the broadphase margin of each other (empty list = shards still independent); with
        `migrate` the offending islands have changed rank by the time the call returns."""
        self.world.step(n)
        if not check:
            return []
        st = self.world.download_state(aabb=True)
        bounds = self.exchange_bounds(self.local_bounds(st["aabb"]))
        pairs = overlapping_ranks(bounds)
        if pairs and migrate and self.dist is not None and self.world_size > 1:
            self.migrate(pairs, bounds, st)
        return pairs

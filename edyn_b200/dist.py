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
import time

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
        isdyn = b["kind"] == DYNAMIC
        for x, y in zip(scene["hinges"]["a"].tolist(), scene["hinges"]["b"].tolist()):
            if isdyn[x] and isdyn[y]:                      # static / kinematic nodes do not connect islands
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
    body count (an island joins the next rank once the current one holds its share; a rank is never skipped while
    islands remain).  Static bodies go to every rank.  Returns owner[n] (rank, or -1 = replicated)."""
    b = scene["bodies"]
    lab = initial_islands(scene) if labels is None else np.asarray(labels, np.int64)
    owner = np.full(len(lab), -1, np.int64)
    dyn = np.where(lab >= 0)[0]
    if len(dyn) == 0:
        return owner
    ids, inv = np.unique(lab[dyn], return_inverse=True)
    size = np.bincount(inv, minlength=len(ids))
    cx = np.bincount(inv, weights=b["pos"][dyn, 0].astype(np.float64), minlength=len(ids)) / size
    order = np.argsort(cx, kind="stable")
    before = np.cumsum(size[order]) - size[order]                      # bodies placed before each island
    share = (before * world_size) // size.sum()                        # rank whose share the island starts in
    k = np.arange(len(order))
    rank_sorted = np.minimum(np.minimum.accumulate(share - k) + k, world_size - 1)     # at most one rank further per island
    rank_of = np.empty(len(ids), np.int64)
    rank_of[order] = rank_sorted
    owner[dyn] = rank_of[inv]
    return owner


def device_islands(scene, device=0):
    """Island label per body (-1 for non-dynamic) as the simulation itself sees them at step 0: one broadphase +
    narrowphase + island pass of the WHOLE scene on `device` (connected components of island_manager's graph).  The
    host-side `initial_islands` is O(minutes) beyond ~10^5 bodies; this is what bench.py partitions with."""
    from .scenes import build_world
    from .world import PH_BROAD, PH_ISLANDS, PH_NARROW
    w = build_world(scene, device=device)
    w.run_phases(PH_BROAD | PH_NARROW | PH_ISLANDS)
    lab = w.islands().astype(np.int64)
    w.close()
    lab[lab == 0xFFFFFFFF] = -1
    return lab


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
        # a joint lives where its dynamic endpoint(s) live; a static / kinematic endpoint is replicated on every rank
        oa, ob = owner[h["a"]], owner[h["b"]]
        sel = ((oa == rank) | (ob == rank)) & ((oa == rank) | (oa < 0)) & ((ob == rank) | (ob < 0))
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
    """bounds: (world_size, 6) array of min/max of each rank's dynamic bodies (NaN or inverted rows = rank owns nothing).
    Returns the list of rank pairs (i < j) whose boxes, inflated by `margin`, intersect."""
    b = np.asarray(bounds, np.float64)
    lo, hi = b[:, None, 0:3], b[:, None, 3:6]
    with np.errstate(invalid="ignore"):
        hit = np.all((lo - margin <= hi.transpose(1, 0, 2)) & (hi + margin >= lo.transpose(1, 0, 2)), axis=2)
    i, j = np.nonzero(np.triu(hit, 1))
    return list(zip(i.tolist(), j.tolist()))


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
            # both endpoints travel, or one travels and the other is a replicated static body
            static = np.zeros(w.num_bodies, bool); static[:len(self.local["bodies"]["kind"])] = self.local["bodies"]["kind"] != DYNAMIC
            m = w.hinge_alive & (sel[h["a"]] | sel[h["b"]]) & (sel[h["a"]] | static[h["a"]]) & (sel[h["b"]] | static[h["b"]])
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


# ======================================================================================================================
# Device-resident flavour (what bench.py --gpus N and the GPU tests run): nothing but 24-byte rank boxes crosses the
# host per step, and a hand-over moves DEVICE blobs over the communicator.

class TorchComm:
    """torch.distributed transport (NCCL over NVLink on GPUs).  All tensors are device tensors; collectives are issued
    on the caller's current stream, which DeviceShardedWorld sets to the world's own stream."""

    def __init__(self, dist, rank, world_size):
        self.dist, self.rank, self.world_size = dist, rank, world_size
        self.bytes_sent = 0                         # payload bytes this rank contributed to collectives / sends

    def all_gather(self, t):
        import torch
        out = torch.empty((self.world_size,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t.contiguous())
        self.bytes_sent += t.numel() * t.element_size()
        return out

    def warmup(self, device):
        """NCCL sets its point-to-point channels up lazily (hundreds of milliseconds on first use): touch every pair once
        so a hand-over inside a timed region pays for the transfer only."""
        import torch
        sent = self.bytes_sent
        # a small and a blob-sized message: NCCL connects further channels the first time a pair moves megabytes
        for nbytes in (16, 8 << 20):
            send = {p: torch.zeros(nbytes, dtype=torch.uint8, device=device) for p in range(self.world_size) if p != self.rank}
            self.exchange(send, {p: nbytes for p in send}, device)
            torch.cuda.synchronize()
        self.all_gather(torch.zeros(8, dtype=torch.float32, device=device))
        self.all_gather(torch.zeros((4096, 8), dtype=torch.float32, device=device))
        self.all_gather(torch.zeros(1, dtype=torch.int32, device=device))
        torch.cuda.synchronize()
        self.bytes_sent = sent

    def exchange(self, send, recv_bytes, device):
        """send: {dst: uint8 tensor}; recv_bytes: {src: nbytes}.  Grouped ncclSend / ncclRecv."""
        import torch
        recv = {src: torch.empty(n, dtype=torch.uint8, device=device) for src, n in recv_bytes.items()}
        ops = [self.dist.P2POp(self.dist.isend, t, dst) for dst, t in sorted(send.items())]
        ops += [self.dist.P2POp(self.dist.irecv, t, src) for src, t in sorted(recv.items())]
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()
        self.bytes_sent += sum(t.numel() for t in send.values())
        return recv


class ThreadComm:
    """The same interface for `world_size` ranks living in threads of ONE process on ONE device: the GPU tests use it so
    the hand-over path is exercised (and the native library loaded) on a single-GPU box too."""

    class Shared:
        def __init__(self, world_size):
            import threading
            self.world_size = world_size
            self.barrier = threading.Barrier(world_size)
            self.slots = [None] * world_size
            self.mail = {}

    def __init__(self, shared, rank):
        self.s, self.rank, self.world_size = shared, rank, shared.world_size
        self.bytes_sent = 0

    def warmup(self, device):
        pass

    def all_gather(self, t):
        import torch
        torch.cuda.current_stream().synchronize()
        self.s.slots[self.rank] = t
        self.s.barrier.wait()
        out = torch.stack([x.clone() for x in self.s.slots])
        torch.cuda.current_stream().synchronize()
        self.s.barrier.wait()
        self.bytes_sent += t.numel() * t.element_size()
        return out

    def exchange(self, send, recv_bytes, device):
        import torch
        torch.cuda.current_stream().synchronize()
        for dst, t in send.items():
            self.s.mail[(self.rank, dst)] = t
        self.s.barrier.wait()
        recv = {src: self.s.mail[(src, self.rank)].clone() for src in recv_bytes}
        torch.cuda.current_stream().synchronize()
        self.s.barrier.wait()
        for dst in send:
            del self.s.mail[(self.rank, dst)]
        self.bytes_sent += sum(t.numel() for t in send.values())
        return recv


class DeviceShardedWorld:
    """One rank's share of a scene with the cross-rank exchange done on the device.

    Per step (`step`): b2d_step; the box of the rank's dynamic bodies is reduced on the device straight into the
    communicator's send buffer (b2d_device_bounds), all-gathered (24 B per rank), and the N x 6 floats are the only
    thing the host reads -- it has to, because whether a hand-over happens decides what the NEXT step's broadphase sees.

    When two rank boxes come within the broadphase margin (`_handover`): each rank computes its island AABBs on the
    device and lists those near a peer (b2d_island_halo); the lists are all-gathered; an island that touches an island
    of a lower rank is handed over whole to the lowest such rank (b2d_handover_plan, merge_islands' "move into the
    other island", island_manager.cpp:297-350); the blobs (b2d_handover_pack: bodies with state, manifolds with
    points / lifetimes / warm-start impulses, joints, exclusions, named by entity) travel as device buffers over grouped
    send / recv, sizes having been all-gathered first; b2d_handover_unpack appends them.  The check repeats until no
    island moves (an island arriving at rank r may itself touch an island of a still lower rank)."""

    def __init__(self, scene, rank, world_size, comm, device=0, owner=None, labels=None, slack=0.25, pipeline=False, **kw):
        import torch
        from .scenes import build_world
        self.torch = torch
        self.rank, self.world_size, self.comm, self.device = rank, world_size, comm, device
        self.owner = partition(scene, world_size, labels) if owner is None else np.asarray(owner, np.int64)
        self.local = shard(scene, rank, world_size, self.owner)
        n_local = len(self.local["bodies"]["kind"])
        n_all = len(self.owner)
        # room for arrivals: `slack` of the whole scene on top of the own share (ids of departed bodies are not reused)
        kw.setdefault("max_bodies", min(2 * n_all, n_local + int(slack * n_all) + 1024))
        nh_all = len(scene["hinges"]["a"]) if scene.get("hinges") else 0
        nh_local = len(self.local["hinges"]["a"]) if self.local.get("hinges") else 0
        kw.setdefault("max_hinges", (nh_local + int(slack * nh_all) + 1024) if nh_all else 0)
        self.world = build_world(self.local, device=device, **kw)
        self.world.set_entities(0, self.local["global_ids"].astype(np.uint32))
        self.ext = torch.cuda.ExternalStream(self.world.stream, device=device)
        dev = torch.device("cuda", device)
        self.dev = dev
        self.bounds = torch.zeros(8, dtype=torch.float32, device=dev)          # min xyz, max xyz, speed, 0
        # pipeline: the boxes of step k are read while step k + 1 already runs (no host round trip on the GPU's critical
        # path); the decision they trigger takes effect one step later, so every margin is widened by what two bodies
        # can travel towards each other in a step (speeds may double in a collision; gravity adds g dt)
        self.pipeline = bool(pipeline)
        self.dt = float(self.world.fixed_dt)
        self.pending = None                                                  # (pinned host copy of the gathered boxes, event)
        self.host_boxes = [torch.zeros((world_size, 8), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.host_events = [torch.cuda.Event() for _ in range(2)]
        self.flip = 0
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.rec_cap = kw["max_bodies"]
        self.records = torch.zeros((self.rec_cap, 8), dtype=torch.float32, device=dev)
        self.migrated_in = self.migrated_out = 0
        self.handover_rounds = 0
        self.halo_checks = 0
        self.dynamic = int(self.local["dynamic"])           # dynamic bodies this rank owns right now
        self.last_pairs = []
        self.handover_ms = []                               # host wall time of every hand-over round (plan .. unpack)
        self.handover_phases = []                           # the same, split: halo / plan / pack / exchange / unpack
        with torch.cuda.stream(self.ext):
            comm.warmup(dev)

    def close(self):
        # pinned buffers and events must go before the CUDA context does (interpreter teardown order is not ours)
        self.torch.cuda.synchronize()
        self.host_boxes = self.host_events = None
        self.pending = None
        self.world.close()

    def _gather_bounds(self):
        self.world.device_bounds(self.bounds.data_ptr())
        g = self.comm.all_gather(self.bounds)
        return g, g.cpu().numpy()                            # the one per-step host read: N x 8 floats

    def _margin(self, gh, lookahead):
        """Broadphase margin, widened by `lookahead` steps of closing travel at the fastest speeds seen on any rank."""
        if not lookahead:
            return MARGIN
        vmax = float(np.nanmax(gh[:, 6])) if len(gh) else 0.0
        return MARGIN + lookahead * self.dt * (4.0 * vmax + 2.0 * 9.8 * self.dt)

    def _resolve(self, g, gh, lookahead):
        """Hand islands over until no two rank boxes are within the margin of each other (blocking)."""
        margin = self._margin(gh, lookahead)
        pairs = overlapping_ranks(gh[:, :6], margin)
        while pairs:
            self.halo_checks += 1
            t0 = time.perf_counter()
            self.world.set_halo_margin(margin)
            moved = self._handover(pairs, g[:, :6].contiguous())
            if moved == 0:
                break
            self.handover_rounds += 1
            self.handover_ms.append((time.perf_counter() - t0) * 1e3)
            g, gh = self._gather_bounds()
            margin = self._margin(gh, lookahead)
            pairs = overlapping_ranks(gh[:, :6], margin)
        return pairs

    def exchange(self):
        """After a step: rank boxes, and a hand-over if any two came within the broadphase margin of each other."""
        torch = self.torch
        with torch.cuda.stream(self.ext):
            if not self.pipeline:
                g, gh = self._gather_bounds()
                pairs = self._resolve(g, gh, 0)
            else:
                # enqueue this step's boxes (reduction, all-gather, copy to pinned memory), then look at the PREVIOUS step's
                self.world.device_bounds(self.bounds.data_ptr())
                g = self.comm.all_gather(self.bounds)
                k = self.flip
                self.host_boxes[k].copy_(g, non_blocking=True)
                self.host_events[k].record(self.ext)
                prev, self.pending, self.flip = self.pending, k, 1 - k
                pairs = []
                if prev is not None:
                    self.host_events[prev].synchronize()         # a step old: long done
                    gh = self.host_boxes[prev].numpy()
                    if overlapping_ranks(gh[:, :6], self._margin(gh, 1)):
                        g, gh = self._gather_bounds()            # blocking path, on the current state
                        pairs = self._resolve(g, gh, 1)
                        self.pending = None                      # the boxes in flight predate the hand-over
        self.last_pairs = pairs
        return pairs

    def _handover(self, pairs, g):
        torch, w, N, r = self.torch, self.world, self.world_size, self.rank
        mask = 0
        for i, j in pairs:
            if i == r:
                mask |= 1 << j
            if j == r:
                mask |= 1 << i
        tm = {}
        t_ = time.perf_counter()
        w.island_halo(g.data_ptr(), N, r, mask, self.records.data_ptr(), self.rec_cap, self.count.data_ptr())
        nrec = self.comm.all_gather(self.count).cpu().numpy().reshape(-1).astype(np.int64)
        if nrec[r] > self.rec_cap:
            raise RuntimeError("halo record buffer overflow")
        if nrec.sum() == 0:
            return 0
        mx = int(nrec.max())
        allrec = self.comm.all_gather(self.records[:mx])                           # (N, mx, 8): 32 B per boundary island
        recs = torch.cat([allrec[p, :int(nrec[p])] for p in range(N)]).contiguous()
        my_b = int(nrec[:r].sum())
        tm["halo"] = (time.perf_counter() - t_) * 1e3; t_ = time.perf_counter()
        plan = w.handover_plan(recs.data_ptr(), my_b, my_b + int(nrec[r]), N)      # (N, 4) host
        tm["plan"] = (time.perf_counter() - t_) * 1e3; t_ = time.perf_counter()
        allplan = self.comm.all_gather(torch.from_numpy(plan.astype(np.int32)).to(self.dev)).cpu().numpy()   # (N, N, 4)
        moved = int(allplan[:, :, 0].sum())
        if moved == 0:
            return 0
        send = {}
        for dst in range(N):
            if plan[dst].any():
                nbytes = w.handover_bytes(plan[dst])
                blob = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
                w.handover_pack(dst, blob.data_ptr(), nbytes)
                send[dst] = blob
        tm["pack"] = (time.perf_counter() - t_) * 1e3; t_ = time.perf_counter()
        recv_bytes = {src: w.handover_bytes(allplan[src, r]) for src in range(N) if allplan[src, r].any()}
        got = self.comm.exchange(send, recv_bytes, self.dev)
        tm["exchange"] = (time.perf_counter() - t_) * 1e3; t_ = time.perf_counter()
        for src in sorted(got):
            c = w.handover_unpack(got[src].data_ptr(), got[src].numel())
            self.migrated_in += int(c[0])
        tm["unpack"] = (time.perf_counter() - t_) * 1e3
        self.handover_phases.append({k: round(v, 3) for k, v in tm.items()})
        out = int(plan[:, 0].sum())
        self.migrated_out += out
        self.dynamic += int(allplan[:, r, 0].sum()) - out
        return moved

    def step(self, n=1, exchange=True):
        for _ in range(n):
            self.world.step(1)
            if exchange:
                self.exchange()
        return self.last_pairs

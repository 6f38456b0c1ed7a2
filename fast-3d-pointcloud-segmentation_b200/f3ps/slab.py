"""Slab mode: ONE very large cloud on several GPUs (BASELINE config 5).

The reference runs pcl::SupervoxelClustering::extract + Clustering::cluster on the whole cloud in one thread
(src/supervoxel_clustering.cpp:348-367, 408-449).  Here every rank (one process per GPU, torch.distributed)
starts with an arbitrary share of the points -- e.g. some of the scan positions of a merged room scan -- and the cloud
is cut into spatial slabs: contiguous ranges of the x-major Morton key of PCL's adjacency octree.  Because that
key order IS PCL's leaf order, a slab's voxels are a contiguous slice of the single-GPU voxel table and every
ordered float sum keeps its order, so the result is bit-identical to one handle processing the whole cloud
(tests/test_gpu_slab.py).

Exchanges (all on the stream the handle runs on; NCCL over NVLink when the backend is nccl):
  K1  all-reduce MIN/MAX of the transformed bounding box (6 floats)      -> same frame / key grid on every rank
      all-reduce SUM of a 4096-bin histogram of the keys' top bits       -> equal-count splitters
      all-to-all of the points, 16 bytes each, to the rank owning their key range (stable: input order survives)
      all-gather of the voxel slices (40 bytes per voxel)                -> replicated voxel table
  K3  normals of the owned slice, all-gather of the slices (20 bytes per voxel)
  K5  per sweep: the owned slice of the steal table (4 bytes per voxel) + a 4-byte convergence flag (MAX);
      per round: owner words of the slice (4 bytes per voxel) + helper sizes (SUM)
  K2 (hash + 27 probes), K4 (seed grid), the per-round helper lists / centroid folds and K6 run on the replicated
  tables on every rank; K7 does not shard ("replicas only": one graph, strictly serial order) and is replayed by
  every rank, so every rank ends with the complete result and no broadcast is needed.
"""
import numpy as np

from . import binding


class _DevView:
    """A raw device pointer as something torch.as_tensor understands (zero copy)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_view(torch, ptr, n, elem_bytes, dtype, device):
    """torch tensor [n, elem_bytes / itemsize] aliasing n elements at device pointer ptr."""
    if n == 0 or not ptr:
        return torch.empty((0, max(1, elem_bytes // torch.empty((), dtype=dtype).element_size())), dtype=dtype, device=device)
    raw = torch.as_tensor(_DevView(ptr, n * elem_bytes), device=device)
    return raw.view(dtype).view(n, -1)


def choose_splitters(hist, world, shift):
    """Equal-count cuts of the key space from the global histogram of the keys' top bits.
    Returns world-1 ascending Morton keys; rank r owns [splitters[r-1], splitters[r])."""
    hist = np.asarray(hist, np.int64)
    total = int(hist.sum())
    cum = np.cumsum(hist)
    cuts = []
    for r in range(1, world):
        target = (total * r) // world
        b = int(np.searchsorted(cum, target, side="left")) + 1      # first bin boundary at or after the target
        b = min(max(b, cuts[-1] if cuts else 0), hist.shape[0])
        cuts.append(b)
    return np.array([np.uint64(b) << np.uint64(shift) for b in cuts], np.uint64)


def slice_bounds(counts):
    b = np.zeros(len(counts) + 1, np.int64)
    b[1:] = np.cumsum(counts)
    return b


class Comm:
    """The collectives the slab driver needs, on torch.distributed.  backend nccl: straight on the device tensors.
    backend gloo (CPU tests, or several ranks sharing one GPU in the parity test): staged through host memory."""

    def __init__(self, dist=None, torch=None):
        self.dist, self.torch = dist, torch
        self.on = dist is not None and dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank() if self.on else 0
        self.world = dist.get_world_size() if self.on else 1
        self.staged = self.on and dist.get_backend() != "nccl"
        self.bytes_moved = 0

    def all_reduce(self, t, op):
        if not self.on or self.world == 1:
            return t
        d = self.dist
        rop = {"min": d.ReduceOp.MIN, "max": d.ReduceOp.MAX, "sum": d.ReduceOp.SUM}[op]
        self.bytes_moved += t.numel() * t.element_size()
        if self.staged and t.is_cuda:
            h = t.cpu()
            d.all_reduce(h, op=rop)
            t.copy_(h)
        else:
            d.all_reduce(t, op=rop)
        return t

    def all_gather_int(self, value):
        if not self.on or self.world == 1:
            return [int(value)]
        t = self.torch.tensor([int(value)], dtype=self.torch.int64)
        out = [self.torch.zeros(1, dtype=self.torch.int64) for _ in range(self.world)]
        if not self.staged:
            dev = self.torch.device("cuda", self.torch.cuda.current_device())
            t = t.to(dev); out = [o.to(dev) for o in out]
        self.dist.all_gather(out, t)
        return [int(o[0]) for o in out]

    def all_gather_vec(self, vec):
        """vec: int64 numpy [k] -> [world, k]"""
        if not self.on or self.world == 1:
            return np.asarray(vec, np.int64)[None, :]
        t = self.torch.from_numpy(np.ascontiguousarray(vec, np.int64))
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        if not self.staged:
            dev = self.torch.device("cuda", self.torch.cuda.current_device())
            t = t.to(dev); out = [o.to(dev) for o in out]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def gather_slices(self, full, bounds):
        """full: tensor whose rows [bounds[r], bounds[r+1]) are valid on rank r; afterwards all rows are valid everywhere."""
        if not self.on or self.world == 1:
            return full
        for r in range(self.world):
            lo, hi = int(bounds[r]), int(bounds[r + 1])
            if hi <= lo:
                continue
            part = full[lo:hi]
            self.bytes_moved += part.numel() * part.element_size()
            if self.staged and part.is_cuda:
                h = part.cpu() if r == self.rank else self.torch.empty(part.shape, dtype=part.dtype)
                self.dist.broadcast(h, src=r)
                if r != self.rank:
                    part.copy_(h)
            else:
                self.dist.broadcast(part, src=r)
        return full

    def gather_slices_flag(self, full, bounds, flag):
        """gather_slices(full) + all-reduce MAX of the one-element `flag` in ONE collective: every rank contributes its
        slice padded to the longest one with the flag appended (all-gather), the slices are copied into place.
        Returns the reduced flag on the host."""
        if not self.on or self.world == 1:
            return int(flag.item())
        if self.staged:
            self.gather_slices(full, bounds)
            self.all_reduce(flag, "max")
            return int(flag.item())
        torch = self.torch
        width = full.shape[1]
        lens = [int(bounds[r + 1] - bounds[r]) for r in range(self.world)]
        cap = max(lens) * width + 1
        key = (full.dtype, cap, full.device)
        if getattr(self, "_gs_key", None) != key:
            self._gs_send = torch.empty(cap, dtype=full.dtype, device=full.device)
            self._gs_recv = torch.empty((self.world, cap), dtype=full.dtype, device=full.device)
            self._gs_key = key
        send, recv = self._gs_send, self._gs_recv
        lo, hi = int(bounds[self.rank]), int(bounds[self.rank + 1])
        if hi > lo:
            send[:(hi - lo) * width].copy_(full[lo:hi].reshape(-1))
        send[cap - 1:cap].copy_(flag.view(full.dtype))
        self.dist.all_gather_into_tensor(recv, send)
        self.bytes_moved += cap * full.element_size() * (self.world - 1)
        for r in range(self.world):
            if r != self.rank and lens[r]:
                full[int(bounds[r]):int(bounds[r + 1])].copy_(recv[r, :lens[r] * width].view(lens[r], width))
        return int(recv[:, cap - 1].max().item())

    def all_to_all_rows(self, send, send_counts, recv_counts):
        """send: [n, k] rows grouped by destination; returns [sum(recv_counts), k] rows grouped by source rank."""
        torch = self.torch
        n_recv = int(np.sum(recv_counts))
        recv = torch.empty((n_recv, send.shape[1]), dtype=send.dtype, device=send.device)
        if not self.on or self.world == 1:
            recv.copy_(send[:n_recv])
            return recv
        self.bytes_moved += int(np.sum(send_counts)) * send.shape[1] * send.element_size()
        if not self.staged:
            self.dist.all_to_all_single(recv, send[:int(np.sum(send_counts))].contiguous(),
                                        [int(c) for c in recv_counts], [int(c) for c in send_counts])
            return recv
        # gloo has no all-to-all: one scatter per source rank, on host tensors
        sb = slice_bounds(send_counts)
        rb = slice_bounds(recv_counts)
        hs = send.cpu()
        for src in range(self.world):
            dst = torch.empty((int(recv_counts[src]), send.shape[1]), dtype=send.dtype)
            if src == self.rank:
                chunks = [hs[int(sb[r]):int(sb[r + 1])].contiguous() for r in range(self.world)]
                # scatter needs equal shapes: pad to the largest chunk
                m = max(1, max(c.shape[0] for c in chunks))
                padded = [torch.zeros((m, send.shape[1]), dtype=send.dtype) for _ in chunks]
                for p, c in zip(padded, chunks):
                    p[:c.shape[0]] = c
                sizes = torch.tensor([m], dtype=torch.int64)
            else:
                padded, sizes = None, torch.zeros(1, dtype=torch.int64)
            self.dist.broadcast(sizes, src=src)
            buf = torch.zeros((int(sizes[0]), send.shape[1]), dtype=send.dtype)
            self.dist.scatter(buf, scatter_list=padded, src=src)
            dst.copy_(buf[:dst.shape[0]])
            recv[int(rb[src]):int(rb[src + 1])].copy_(dst)
        return recv


class SlabSegmenter:
    """One rank of a slab-mode run.  `run(points, threshold)` takes THIS rank's share of the cloud (numpy records of
    16 or 32 bytes, or a (device pointer, n, stride) triple) and leaves the complete result on this rank's handle
    (`self.seg`: counts(), array(...), stage getters -- identical on every rank)."""

    TOP_BITS = 12

    SHARD_EXPAND_MIN_V = 4_000_000      # voxels: below this the per-sweep exchange costs more than sharding the sweeps saves

    def __init__(self, comm, device=0, vccs=None, merge=None, shard_expand=None):
        import torch
        self.torch = torch
        self.comm = comm
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.seg = binding.Segmenter(device=device, stream=self.stream.cuda_stream)
        self.seg.set_vccs_params(**(vccs or {}))
        self.seg.set_merge_params(**(merge or {}))
        self.info = {}
        self._views = {}
        # K5: None = by size (sweeps sharded over the slabs with a steal-table exchange per sweep when V >= SHARD_EXPAND_MIN_V, else
        # every rank runs the persistent expansion kernel on the replicated voxel table: measured on the 50 M-point scan, V = 0.6 M:
        # 8.8 ms replicated against 13.9 / 18.0 / 37.6 ms sharded over 2 / 4 / 8 GPUs); True / False force it (tests run both)
        self.shard_expand = shard_expand

    def close(self):
        self.seg.close()

    def _view(self, name, dtype):
        ptr, n, eb = self.seg.slab_array(name)
        key = (ptr, n, eb, dtype)                     # the handle's buffers alternate between a few addresses: wrap each once
        v = self._views.get(key)
        if v is None:
            if len(self._views) > 64:
                self._views.clear()
            v = self._views[key] = device_view(self.torch, ptr, n, eb, dtype, self.device)
        return v

    def run(self, points, threshold=0.2, merge=True):
        torch, comm, seg = self.torch, self.comm, self.seg
        ev = []

        def tick(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record(self.stream)
            ev.append((name, e))

        self._views.clear()
        with torch.cuda.stream(self.stream):
            tick("start")
            seg.slab_reset()
            if isinstance(points, tuple):
                ptr, n_local, stride = points
                seg.set_input_device(ptr, n_local, stride)
            else:
                n_local = int(points.shape[0])
                seg.set_input(points)
            # ---- K1a: the frame of the whole cloud ----
            box = torch.zeros(8, dtype=torch.int32, device=self.device)
            seg.slab_bbox(box.data_ptr())
            if comm.world > 1:
                sgn = torch.tensor(-2 ** 31, dtype=torch.int32, device=self.device)
                ob = box ^ sgn                                   # unsigned order -> signed order
                comm.all_reduce(ob[0:3], "min"); comm.all_reduce(ob[3:7], "max")
                box = ob ^ sgn
            seg.slab_set_frame(box.data_ptr())
            # ---- K1b: keys, splitters, routing ----
            hist = torch.zeros(1 << self.TOP_BITS, dtype=torch.int32, device=self.device)
            used, shift = seg.slab_keys(self.TOP_BITS, hist.data_ptr())
            comm.all_reduce(hist, "sum")
            splitters = choose_splitters(hist[: 1 << used].cpu().numpy(), comm.world, shift)
            send = torch.empty((max(1, n_local), 4), dtype=torch.float32, device=self.device)
            send_counts = seg.slab_route(comm.world, splitters, send.data_ptr())
            tick("route")
            count_matrix = comm.all_gather_vec(send_counts)      # [src, dst]
            recv_counts = count_matrix[:, comm.rank]
            recv = comm.all_to_all_rows(send, send_counts, recv_counts)
            del send
            tick("all_to_all")
            # ---- K1c: voxels of the owned slab, then the replicated table ----
            n_recv = int(recv.shape[0])
            if n_recv:
                seg.set_input_device(recv.data_ptr(), n_recv, 16)
            else:
                seg.set_input(np.zeros((0, 4), np.float32).view(np.dtype([("xyzc", "<f4", (4,))])).reshape(0))
            seg.voxelize()
            v_local = int(seg.counts().n_voxels)
            vb = slice_bounds(comm.all_gather_int(v_local))
            V = int(vb[-1])
            lo, hi = int(vb[comm.rank]), int(vb[comm.rank + 1])
            full_xyz = torch.empty((V, 4), dtype=torch.float32, device=self.device)
            full_rgb = torch.empty((V, 4), dtype=torch.float32, device=self.device)
            full_key = torch.empty((V, 1), dtype=torch.int64, device=self.device)
            if v_local:
                full_xyz[lo:hi].copy_(self._view("vox_xyz", torch.float32)[:v_local])
                full_rgb[lo:hi].copy_(self._view("vox_rgb", torch.float32)[:v_local])
                full_key[lo:hi].copy_(self._view("vox_key", torch.int64)[:v_local])
            tick("voxelize")
            comm.gather_slices(full_xyz, vb); comm.gather_slices(full_rgb, vb); comm.gather_slices(full_key, vb)
            seg.slab_set_voxels(full_xyz.data_ptr(), full_rgb.data_ptr(), full_key.data_ptr(), V, lo, hi)
            del full_xyz, full_rgb, full_key, recv
            tick("gather_voxels")
            # ---- K2 (replicated), K3 (owned slice + exchange), K4 (replicated) ----
            seg.neighbors()
            tick("neighbors")
            seg.normals()
            if V:
                comm.gather_slices(self._view("vox_normal", torch.float32), vb)
                comm.gather_slices(self._view("vox_curv", torch.float32), vb)
            tick("normals")
            seg.seeds()
            tick("seeds")
            # ---- K5: sweeps over the owned slice, steal table exchanged after every sweep -- or, for a voxel table too small for
            #      that to pay, the single-GPU persistent kernel on the replicated table (identical results, no exchange) ----
            shard = self.shard_expand if self.shard_expand is not None else (comm.world > 1 and V >= self.SHARD_EXPAND_MIN_V)
            sweeps = 0
            rounds = 0
            if not shard:
                seg.expand()
                sweeps = int(seg.counts().sweeps)
            else:
                seg.slab_expand_begin()
                rounds = int(seg.counts().rounds)
            flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            for _ in range(rounds):
                for s in range(33):
                    if s == 32:
                        raise binding.F3psError("expansion fixed point not reached within 32 sweeps")
                    seg.slab_expand_sweep(flag.data_ptr())
                    sweeps += 1
                    # the slices of the steal table this sweep wrote + the convergence flag, one collective
                    changed = comm.gather_slices_flag(self._view("steal", torch.int32), vb, flag) if V else 0
                    if changed == 0:
                        break
                if V:
                    comm.gather_slices(self._view("owner_next", torch.int32), vb)
                comm.all_reduce(self._view("count", torch.int32), "sum")
                seg.slab_expand_round_end()
            if shard:
                if rounds == 0:
                    seg.slab_expand_round_end()
                elif V:
                    comm.gather_slices(self._view("dist", torch.float32), vb)
                seg.slab_expand_end()
            tick("expand")
            # ---- K6, K7 on the replicated tables ----
            seg.graph()
            tick("graph")
            if merge:
                seg.merge(threshold)
                tick("merge")
            seg.sync()
        self.stream.synchronize()
        ms = {}
        for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
            ms[n1] = e0.elapsed_time(e1)
        ms["total"] = ev[0][1].elapsed_time(ev[-1][1])
        self.info = {"n_local": n_local, "n_received": n_recv, "v_local": v_local, "V": V, "own": (lo, hi), "sweeps": sweeps,
                     "splitters": [int(x) for x in splitters], "send_counts": [int(c) for c in send_counts],
                     "stage_ms": ms, "bytes_exchanged": comm.bytes_moved}
        return self.info

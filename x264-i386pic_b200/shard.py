"""Work partitioning across the GPUs of one box.  Encode streams shard with no data-path
collective: every rank computes its own streams and returns per-frame cost/MV arrays to its host.  ONE stream sharded
over the GPUs (its search fields and evaluations) has one exchange step, the newest frame's list-0 MV fields, done by
the product's own kernels over peer memory (PeerExchange / x264cu_peer_*); an NCCL all-gather arm is kept to compare."""


def assign_streams(n_streams, world, rank):
    """configs[4]: encode streams round-robin over ranks (64 streams / 8 GPUs -> 8 each)."""
    return [s for s in range(n_streams) if s % world == rank]


def assign_fields(fields, world):
    """configs[3]: one stream's search fields (b, ref, list) spread over the GPUs by (b + ref) mod world;
    the evaluation (p0, p1, b) then runs where b's fields live.  Returns {rank: [fields]}."""
    out = {r: [] for r in range(world)}
    for (b, ref, lst) in fields:
        out[(b + ref) % world].append((b, ref, lst))
    return out


def stream_seed(global_stream_id):
    """synthetic content seed of a stream; depends on the global id only, so any world size sees the same streams."""
    return 7 * global_stream_id + 1


# ---- one encode stream over several GPUs (BASELINE.json configs[3]: 4K, frame-pair sharding) --------------------------
def step_units(n, bframes, rank, world):
    """This rank's share of the step that frame n opens (host/stream.cpp x264cu_stream_push_shard uses the same rule):
    list-0 search (n -> p0) and P evaluation (p0, n, n): rank p0 % world; list-1 search (b -> n) and B evaluations
    (p0, n, b): rank b % world; I evaluation: rank n % world.  Returns (l0_searches [p0], l1_searches [b], evals [(p0, n, b)])."""
    lo = max(0, n - (bframes + 1))
    l0 = [p0 for p0 in range(n - 1, lo - 1, -1) if p0 % world == rank]
    l1 = [b for b in range(lo + 1, n) if b % world == rank]
    ev = [(n, n, n)] if n % world == rank else []
    for p0 in range(n - 1, lo - 1, -1):
        if p0 % world == rank:
            ev.append((p0, n, n))
        ev += [(p0, n, b) for b in range(p0 + 1, n) if b % world == rank]
    return l0, l1, ev


class PeerExchange:
    """The ranks' mailboxes (x264cu_peer_*): created once per context, shared by every ShardedWindow opened on it.
    The IPC handles travel through torch.distributed's object all-gather (host side, once); after that the exchange is
    kernels storing into peer HBM over NVLink and flags polled on the device -- no collective, no host synchronisation."""

    def __init__(self, x264cu, ctx, bframes, rank, world, dist):
        import ctypes as C
        self.x, self.ctx, self.L, self.rank, self.world, self._C = x264cu, ctx, ctx.L, rank, world, C
        self.items = 2 * ((bframes + 1 + world - 1) // world)          # MV field + cost field per owned list-0 search
        handle, off = (C.c_uint8 * 64)(), C.c_uint64()
        self.p = self.L.x264cu_peer_create(ctx.ctx, rank, world, self.items, handle, C.byref(off))
        if not self.p:
            raise RuntimeError("x264cu_peer_create: " + x264cu.last_error())
        mine = (bytes(handle), int(off.value))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        hs = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(h for h, _ in everyone))
        offs = (C.c_uint64 * world)(*[o for _, o in everyone])
        x264cu.check(self.L.x264cu_peer_connect(self.p, hs, offs), "peer_connect")
        dist.barrier()                                                  # nobody pushes before everybody has mapped

    @classmethod
    def local(cls, x264cu, ctxs, bframes):
        """the ranks as contexts of THIS process (one per GPU, or several on one device): mailboxes are connected directly
        (x264cu_peer_connect_local), no IPC handles and no torch.distributed.  Returns one PeerExchange per rank."""
        import ctypes as C
        world = len(ctxs)
        out = []
        for rank, ctx in enumerate(ctxs):
            pe = cls.__new__(cls)
            pe.x, pe.ctx, pe.L, pe.rank, pe.world, pe._C = x264cu, ctx, ctx.L, rank, world, C
            pe.items = 2 * ((bframes + 1 + world - 1) // world)
            handle, off = (C.c_uint8 * 64)(), C.c_uint64()
            pe.p = pe.L.x264cu_peer_create(ctx.ctx, rank, world, pe.items, handle, C.byref(off))
            if not pe.p:
                raise RuntimeError("x264cu_peer_create: " + x264cu.last_error())
            out.append(pe)
        arr = (C.c_void_p * world)(*[pe.p for pe in out])
        for pe in out:
            x264cu.check(pe.L.x264cu_peer_connect_local(pe.p, arr, world), "peer_connect_local")
        return out

    def push(self, ptrs):
        C, n = self._C, len(ptrs)
        a = (C.c_void_p * max(n, 1))(*ptrs)
        sl = (C.c_int * max(n, 1))(*range(n))
        self.x.check(self.L.x264cu_peer_push(self.p, a, sl, n), "peer_push")

    def imports(self, items):
        """items: [(dst pointer, from rank, slot)]"""
        C, n = self._C, len(items)
        a = (C.c_void_p * max(n, 1))(*[i[0] for i in items])
        fr = (C.c_int * max(n, 1))(*[i[1] for i in items])
        sl = (C.c_int * max(n, 1))(*[i[2] for i in items])
        self.x.check(self.L.x264cu_peer_import(self.p, a, fr, sl, n), "peer_import")

    def check(self):
        if self.L.x264cu_peer_error(self.p) != 0:
            raise RuntimeError("peer exchange: an import gave up waiting for a peer's fields")

    def close(self):
        if self.p:
            self.L.x264cu_peer_destroy(self.p)
            self.p = None


class ShardedWindow:
    """One lookahead window held by `world` ranks (one GPU each).  Every rank pushes every frame; a step's searches and
    evaluations are split by step_units(); the list-0 MV fields (+ costs) of the newest frame are the one exchange step
    of the path, between the two phases of x264cu_stream_push_shard: with `peer` (a PeerExchange) the owners store them
    straight into the peers' mailboxes and nothing waits on the host; without it, an NCCL all-gather ordered through
    the host (kept as the plain-collective arm to compare with)."""

    def __init__(self, x264cu, ctx, bframes, rank, world, dist=None, device=None, peer=None):
        import ctypes as C
        self.x, self.ctx, self.B, self.rank, self.world, self.dist = x264cu, ctx, bframes, rank, world, dist
        self.L = ctx.L
        self.win = self.L.x264cu_stream_open(ctx.ctx, bframes, 0)
        assert self.win, x264cu.last_error()
        self.n = 0
        self.fb = ctx.nmb * 4                                   # bytes of an MV field; the cost field has the same size
        self.slots = (bframes + 1 + world - 1) // world
        self.send = self.recv = None
        self.ahead, self.fence = 0, [0, 0]
        self.peer = peer
        if world > 1 and peer is None:
            import torch
            self.torch = torch
            # double-buffered: step n+1 packs while step n is being gathered, step n's imports may still be pending
            # on the context's stream when step n+1 is gathered
            self.send = [torch.zeros(self.slots * 2 * self.fb, dtype=torch.uint8, device=device) for _ in range(2)]
            self.recv = [torch.zeros(world * self.slots * 2 * self.fb, dtype=torch.uint8, device=device) for _ in range(2)]
        self._C = C

    def _field(self, frame, dist_):
        mv, mc = self._C.c_void_p(), self._C.c_void_p()
        self.x.check(self.L.x264cu_field_device(self.ctx.ctx, frame, 0, dist_, self._C.byref(mv), self._C.byref(mc)), "field_device")
        return mv.value, mc.value

    def _phase0(self, n, luma, stride, on_device):
        """upload frame n, enqueue this rank's searches, pack its list-0 fields of frame n for the all-gather"""
        L = self.L
        r = L.x264cu_stream_push_shard(self.win, luma, stride, None, int(on_device), self.rank, self.world, 0)
        if r < 0:
            raise RuntimeError("x264cu_stream_push_shard(0): " + self.x.last_error())
        if self.peer is not None:
            ptrs = []
            if n > 0:
                fn = L.x264cu_stream_frame(self.win, n)
                for p0 in (p0 for p0 in range(n - 1, max(0, n - (self.B + 1)) - 1, -1) if p0 % self.world == self.rank):
                    ptrs += list(self._field(fn, n - p0))
            self.peer.push(ptrs)                 # every step, also with nothing to send: the peers wait for the step's flag
            return
        if self.world > 1 and n > 0:
            fn = L.x264cu_stream_frame(self.win, n)
            lo = max(0, n - (self.B + 1))
            send = self.send[n & 1]
            for k, p0 in enumerate(p0 for p0 in range(n - 1, lo - 1, -1) if p0 % self.world == self.rank):
                mv, mc = self._field(fn, n - p0)
                base = send.data_ptr() + k * 2 * self.fb
                self.x.check(L.x264cu_memcpy_d2d(self.ctx.ctx, base, mv, self.fb), "d2d")
                self.x.check(L.x264cu_memcpy_d2d(self.ctx.ctx, base + self.fb, mc, self.fb), "d2d")
        self.fence[n & 1] = L.x264cu_fence(self.ctx.ctx)

    def push(self, luma, stride, on_device, next_luma=None):
        """frame n = number of frames pushed so far.  next_luma: the frame after it, if the caller already has it (an
        encoder's lookahead queue does): its upload and searches are enqueued BEFORE this step's exchange, so the GPU
        works on them while the host waits for this step's fields and runs the all-gather."""
        L, n = self.L, self.n
        if self.ahead == 0:
            self._phase0(n, luma, stride, on_device)
            self.ahead = 1
        if next_luma is not None:
            self._phase0(n + 1, next_luma, stride, on_device)
            self.ahead = 2
        if self.peer is not None:
            items = []
            if n > 0:
                fn = L.x264cu_stream_frame(self.win, n)
                lo = max(0, n - (self.B + 1))
                for r2 in (r2 for r2 in range(self.world) if r2 != self.rank):
                    for k, p0 in enumerate(p0 for p0 in range(n - 1, lo - 1, -1) if p0 % self.world == r2):
                        mv, mc = self._field(fn, n - p0)
                        items += [(mv, r2, 2 * k), (mc, r2, 2 * k + 1)]
                        self.x.check(L.x264cu_field_set_searched(self.ctx.ctx, fn, 0, n - p0), "set_searched")
            self.peer.imports(items)
        elif self.world > 1 and n > 0:
            fn = L.x264cu_stream_frame(self.win, n)
            lo = max(0, n - (self.B + 1))
            # the searches and the packing ran on the context's stream, NCCL runs on torch's: order them through the host
            self.x.check(L.x264cu_wait(self.ctx.ctx, self.fence[n & 1]), "wait")
            send, recv = self.send[n & 1], self.recv[n & 1]
            self.dist.all_gather_into_tensor(recv, send)
            self.torch.cuda.current_stream().synchronize()
            for r2 in range(self.world):
                if r2 == self.rank:
                    continue
                for k, p0 in enumerate(p0 for p0 in range(n - 1, lo - 1, -1) if p0 % self.world == r2):
                    mv, mc = self._field(fn, n - p0)
                    base = recv.data_ptr() + (r2 * self.slots + k) * 2 * self.fb
                    self.x.check(L.x264cu_memcpy_d2d(self.ctx.ctx, mv, base, self.fb), "d2d")
                    self.x.check(L.x264cu_memcpy_d2d(self.ctx.ctx, mc, base + self.fb, self.fb), "d2d")
                    self.x.check(L.x264cu_field_set_searched(self.ctx.ctx, fn, 0, n - p0), "set_searched")
        r = L.x264cu_stream_push_shard(self.win, None, 0, None, 0, self.rank, self.world, 1)
        if r < 0:
            raise RuntimeError("x264cu_stream_push_shard(1): " + self.x.last_error())
        self.n += 1
        self.ahead -= 1
        return r

    def sync(self):
        self.x.check(self.L.x264cu_flush(self.ctx.ctx), "flush")
        if self.peer is not None:
            self.peer.check()

    def costs(self):
        """this rank's stats of the newest step: {(d0, d1): [cost_est, cost_est_aq, intra_mbs, raw]} (others absent)"""
        out = {}
        st = (self._C.c_int32 * 4)()
        for d0 in range(self.B + 2):
            for d1 in range(self.B + 2 - d0):
                self.L.x264cu_stream_cost(self.win, d0, d1, st)
                if st[0] >= 0:
                    out[(d0, d1)] = list(st)
        return out

    def close(self):
        self.L.x264cu_stream_close(self.win)

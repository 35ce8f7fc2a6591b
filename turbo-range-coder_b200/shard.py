"""Multi-GPU sharding of the batch path: independent blocks per rank, one gather of the compressed chunks.

The path has no data dependency between chunks, so ranks never exchange anything while coding.  The only
exchange is the collection of the packed streams on one rank: an all-gather of the
per-rank compressed lengths (8 bytes each) fixes every rank's offset, then each rank sends its payload to
the destination rank, which receives it at that offset (NCCL has no gatherv; grouped send/recv is its
idiom).  Works with any torch.distributed backend: NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def partition(n_blocks: int, world: int):
    """Contiguous block ranges per rank: block b belongs to rank b * world // n_blocks."""
    bounds = [-(-r * n_blocks // world) for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_bytes(total_len: int, block_len: int, world: int):
    """Byte ranges [start, end) of every rank's slice when the buffer is cut into block_len blocks."""
    n_blocks = -(-total_len // block_len)
    return [(min(a * block_len, total_len), min(b * block_len, total_len)) for a, b in partition(n_blocks, world)]


def gather_compressed(payload: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor = None):
    """Collect every rank's packed stream (1-D uint8 tensor, any length) on rank `dst`.

    Returns (buffer, offsets) on `dst` -- offsets is an int64 tensor of world+1 prefix sums, rank r's
    stream is buffer[offsets[r]:offsets[r+1]] -- and (None, offsets) elsewhere.
    """
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = torch.tensor([payload.numel()], dtype=torch.int64, device=payload.device)
    lens = torch.empty(world, dtype=torch.int64, device=payload.device)
    dist.all_gather_into_tensor(lens, mine, group=group)
    lens_h = lens.cpu()
    offs = torch.zeros(world + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens_h, 0)
    if world == 1:
        return payload, offs
    if rank == dst:
        total = int(offs[-1])
        buf = out if out is not None and out.numel() >= total else torch.empty(total, dtype=torch.uint8, device=payload.device)
        buf[int(offs[dst]):int(offs[dst + 1])] = payload
        ops = [dist.P2POp(dist.irecv, buf[int(offs[r]):int(offs[r + 1])], r, group) for r in range(world)
               if r != dst and lens_h[r] > 0]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        return buf[:total], offs
    if payload.numel() > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, payload, dst, group)]):
            w.wait()
    return None, offs


class PeerGather:
    """Gather of the packed streams over NVLink peer memory, driven from the device (CUDA only).

    Rank `dst` owns `depth` sets of world x slot_bytes slots (+ one length and one flag word per slot); every rank maps the
    buffer through CUDA IPC.  `push()` launches a copy kernel that reads the byte count from device memory (the encoder's
    out_off[n]) and stores the stream into slot (seq % depth, rank) -- for remote ranks the 128-bit stores cross NVSwitch --
    then publishes length and sequence number with release semantics (csrc/pack.cuh k_push).  `wait_all(seq)` on `dst` blocks
    a stream until every rank's push `seq` has landed: after it the gathered streams can be consumed on the device.
    `fetch()` is the mirror (scatter by the directory): a rank copies a slot of the gathered buffer back into local memory.
    Nothing synchronises with the host and no collective is called per step.
    Layout on `dst`: slot (d, r) at (d*world + r)*slot_bytes; lengths then flags (uint64 each) behind the slots.
    `split`: copy engines (streams) that share the predicted part of a push (see push(hint_bytes)).
    """

    def __init__(self, slot_bytes: int, dst: int = 0, group=None, depth: int = 2, split: int = 1):
        import ctypes
        import importlib
        self.ct = ctypes
        self.trc = importlib.import_module("turbo-range-coder_b200")
        lib = self.trc.lib
        for f in (lib.trc_dev_alloc, lib.trc_ipc_export, lib.trc_ipc_open, lib.trc_ipc_close, lib.trc_push_dev, lib.trc_wait_flags_dev, lib.trc_ack_dev,
                  lib.trc_memcpy_dev, lib.trc_dev_free):
            f.restype = ctypes.c_int
        self.world, self.rank, self.dst, self.group, self.depth = dist.get_world_size(group), dist.get_rank(group), dst, group, depth
        self.slot_bytes = (int(slot_bytes) + 255) & ~255
        nslot = depth * self.world
        self.off_lens = nslot * self.slot_bytes
        self.off_flags = self.off_lens + 8 * nslot
        self.off_status = self.off_flags + 8 * nslot
        self.off_ack = self.off_status + 64                 # one acknowledgement word per slot set
        total = self.off_ack + 8 * depth + 256
        self.base = ctypes.c_void_p()
        self.local = ctypes.c_void_p()                      # local scratch: completion counters of this rank's pushes (one per depth)
        self.trc._check(lib.trc_dev_alloc(ctypes.byref(self.local), ctypes.c_size_t(256)), "trc_dev_alloc")
        handle = [None]
        if self.rank == dst:
            self.trc._check(lib.trc_dev_alloc(ctypes.byref(self.base), ctypes.c_size_t(total)), "trc_dev_alloc")   # zeroed
            h = (ctypes.c_ubyte * 64)()
            self.trc._check(lib.trc_ipc_export(self.base, h), "trc_ipc_export")
            handle = [bytes(h)]
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank != dst:
            h = (ctypes.c_ubyte * 64).from_buffer_copy(handle[0])
            self.trc._check(lib.trc_ipc_open(h, ctypes.byref(self.base)), "trc_ipc_open")
        self.total = total
        self.seq = 0                                        # pushes issued by this rank so far
        # one cudaMemcpyAsync to a peer is ONE copy engine (~180 GB/s over NVLink: with 4 GPUs a 72.5 MB push took 0.40 ms and set the
        # step time); the predicted part of a push is cut into `split` pieces on as many streams so that several engines carry it
        self.split = max(1, int(split))
        self.extra = [torch.cuda.Stream() for _ in range(self.split - 1)]
        self.ev_fork = torch.cuda.Event()
        self.ev_join = [torch.cuda.Event() for _ in range(self.split - 1)]

    def _slot(self, d, r):
        return d * self.world + r

    def slot_ptr(self, r, seq=None):
        d = (self.seq if seq is None else seq) % self.depth
        return self.base.value + self._slot(d, r) * self.slot_bytes

    def len_ptr(self, r, seq=None):
        d = (self.seq if seq is None else seq) % self.depth
        return self.base.value + self.off_lens + 8 * self._slot(d, r)

    def flag_ptr(self, r, seq=None):
        d = (self.seq if seq is None else seq) % self.depth
        return self.base.value + self.off_flags + 8 * self._slot(d, r)

    def push(self, payload: torch.Tensor, d_len_ptr: int, stream=None, hint_bytes: int = 0) -> int:
        """payload: the local packed stream buffer (16-byte aligned); d_len_ptr: device address of its uint64 length.
        hint_bytes: a prediction of the length known on the host (the previous step's, say; 0 = none): that many bytes travel by
        copy engine (cudaMemcpyAsync to the peer mapping) and the kernel only moves what lies beyond them and publishes, so the SMs
        stay with the coders.  A wrong hint costs bandwidth (too long) or SM time (too short), never correctness.
        Returns the sequence number of this push (1, 2, ...): the same number on every rank names the same step."""
        st = stream if stream is not None else torch.cuda.current_stream()
        self.seq += 1
        s, ct, lib = self.seq, self.ct, self.trc.lib
        ack = self.base.value + self.off_ack + 8 * (s % self.depth)
        need = max(0, s - self.depth)
        skip = min(int(hint_bytes), self.slot_bytes, payload.numel()) & ~15
        if skip:
            if need:                                          # the slot must be free before the copy engine touches it
                self.trc._check(lib.trc_wait_flags_dev(ct.c_void_p(ack), None, ct.c_uint(1), ct.c_uint64(need), None, ct.c_void_p(st.cuda_stream)),
                                "trc_wait_flags_dev (ack)")
            dst0, src0 = self.slot_ptr(self.rank, s), payload.data_ptr()
            nsp = self.split if (skip >= (1 << 20) and self.rank != self.dst) else 1
            part = (skip // nsp) & ~255
            if nsp > 1:
                self.ev_fork.record(st)                     # the pieces start after the slot is free and the payload is complete
            for i in range(nsp):
                lo, hi = i * part, (skip if i == nsp - 1 else (i + 1) * part)
                si = st if i == 0 else self.extra[i - 1]
                if i:
                    si.wait_event(self.ev_fork)
                self.trc._check(lib.trc_memcpy_dev(ct.c_void_p(dst0 + lo), ct.c_void_p(src0 + lo), ct.c_size_t(hi - lo), ct.c_void_p(si.cuda_stream)),
                                "trc_memcpy_dev")
                if i:
                    self.ev_join[i - 1].record(si)
            for i in range(1, nsp):
                st.wait_event(self.ev_join[i - 1])          # the publishing kernel below runs after every piece
        rc = lib.trc_push_dev(ct.c_void_p(self.slot_ptr(self.rank, s)), ct.c_void_p(payload.data_ptr()), ct.c_void_p(d_len_ptr),
                              ct.c_size_t(0), ct.c_size_t(self.slot_bytes), ct.c_void_p(self.len_ptr(self.rank, s)),
                              ct.c_void_p(self.flag_ptr(self.rank, s)), ct.c_uint64(s),
                              ct.c_void_p(self.local.value + 4 * (s % self.depth)),
                              None if skip else ct.c_void_p(ack), ct.c_uint64(need), ct.c_size_t(skip), ct.c_void_p(st.cuda_stream))
        self.trc._check(rc, "trc_push_dev")
        return s

    def wait_all(self, seq: int, stream=None):
        """(dst only) make `stream` wait until push `seq` of every rank has landed (flags acquired on the device)."""
        assert self.rank == self.dst
        st = stream if stream is not None else torch.cuda.current_stream()
        ct, d = self.ct, seq % self.depth
        rc = self.trc.lib.trc_wait_flags_dev(ct.c_void_p(self.base.value + self.off_flags + 8 * self._slot(d, 0)),
                                             ct.c_void_p(self.base.value + self.off_lens + 8 * self._slot(d, 0)), ct.c_uint(self.world),
                                             ct.c_uint64(seq), ct.c_void_p(self.base.value + self.off_status), ct.c_void_p(st.cuda_stream))
        self.trc._check(rc, "trc_wait_flags_dev")

    def ack(self, seq: int, stream=None):
        """(dst only) the consumer is done with the slot set of push `seq` (stream-ordered): producers may overwrite it.
        push() of step seq + depth waits for this on the device -- without it a fast producer would lap the consumer."""
        assert self.rank == self.dst
        st = stream if stream is not None else torch.cuda.current_stream()
        rc = self.trc.lib.trc_ack_dev(self.ct.c_void_p(self.base.value + self.off_ack + 8 * (seq % self.depth)), self.ct.c_uint64(seq),
                                      self.ct.c_void_p(st.cuda_stream))
        self.trc._check(rc, "trc_ack_dev")

    def fetch(self, r: int, seq: int, out: torch.Tensor, stream=None):
        """The mirror of push (scatter by the directory): copy slot (seq, r) of the gathered buffer -- its byte count is read from
        the gathered lengths, on the device -- into the local tensor `out`.  The caller orders this after the slot is complete
        (own slot: stream order after push(); any slot on dst: after wait_all())."""
        st = stream if stream is not None else torch.cuda.current_stream()
        ct = self.ct
        rc = self.trc.lib.trc_push_dev(ct.c_void_p(out.data_ptr()), ct.c_void_p(self.slot_ptr(r, seq)), ct.c_void_p(self.len_ptr(r, seq)),
                                       ct.c_size_t(0), ct.c_size_t(out.numel() & ~15), None, None, ct.c_uint64(0), None, None, ct.c_uint64(0),
                                       ct.c_size_t(0), ct.c_void_p(st.cuda_stream))
        self.trc._check(rc, "trc_push_dev (fetch)")

    def read_slot(self, r, nbytes, device, seq=None):
        """(dst only) copy slot r into a fresh tensor -- verification helper."""
        out = torch.empty(nbytes, dtype=torch.uint8, device=device)
        rc = self.trc.lib.trc_memcpy_dev(self.ct.c_void_p(out.data_ptr()), self.ct.c_void_p(self.slot_ptr(r, seq)), self.ct.c_size_t(nbytes),
                                         self.ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.trc._check(rc, "trc_memcpy_dev")
        return out

    def read_lens(self, device, seq=None):
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        rc = self.trc.lib.trc_memcpy_dev(self.ct.c_void_p(out.data_ptr()), self.ct.c_void_p(self.len_ptr(0, seq)), self.ct.c_size_t(8 * self.world),
                                         self.ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.trc._check(rc, "trc_memcpy_dev")
        return out

    def overflowed(self, device) -> bool:
        """(dst only) did any pushed stream exceed its slot?  (reads the status word written by wait_all)"""
        out = torch.empty(1, dtype=torch.int32, device=device)
        rc = self.trc.lib.trc_memcpy_dev(self.ct.c_void_p(out.data_ptr()), self.ct.c_void_p(self.base.value + self.off_status), self.ct.c_size_t(4),
                                         self.ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.trc._check(rc, "trc_memcpy_dev")
        return bool(int(out.item()) & 1)

    def close(self):
        if self.base.value:
            if self.rank == self.dst:
                self.trc.lib.trc_dev_free(self.base)
            else:
                self.trc.lib.trc_ipc_close(self.base)
            self.base = self.ct.c_void_p()
        if self.local.value:
            self.trc.lib.trc_dev_free(self.local)
            self.local = self.ct.c_void_p()

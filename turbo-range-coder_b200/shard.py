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

    Rank `dst` owns a buffer of world x slot_bytes (+ one length word per rank); every rank maps it through CUDA IPC
    and `push()` launches a copy kernel that reads the byte count from device memory (the encoder's out_off[n]) and
    stores the stream into slot `rank` -- for remote ranks the 128-bit stores cross NVSwitch.  Nothing synchronises
    with the host and no collective is called per step, so the transfer overlaps whatever runs on other streams.
    Layout on `dst`: slot r at r*slot_bytes, lengths (uint64) at world*slot_bytes + 8*r.
    """

    def __init__(self, slot_bytes: int, dst: int = 0, group=None):
        import ctypes
        import importlib
        self.ct = ctypes
        self.trc = importlib.import_module("turbo-range-coder_b200")
        lib = self.trc.lib
        for f in (lib.trc_dev_alloc, lib.trc_ipc_export, lib.trc_ipc_open, lib.trc_ipc_close, lib.trc_push_dev, lib.trc_memcpy_dev, lib.trc_dev_free):
            f.restype = ctypes.c_int
        self.world, self.rank, self.dst, self.group = dist.get_world_size(group), dist.get_rank(group), dst, group
        self.slot_bytes = (int(slot_bytes) + 255) & ~255
        total = self.world * self.slot_bytes + 8 * self.world + 256
        self.base = ctypes.c_void_p()
        handle = [None]
        if self.rank == dst:
            self.trc._check(lib.trc_dev_alloc(ctypes.byref(self.base), ctypes.c_size_t(total)), "trc_dev_alloc")
            h = (ctypes.c_ubyte * 64)()
            self.trc._check(lib.trc_ipc_export(self.base, h), "trc_ipc_export")
            handle = [bytes(h)]
        dist.broadcast_object_list(handle, src=dst, group=group)
        if self.rank != dst:
            h = (ctypes.c_ubyte * 64).from_buffer_copy(handle[0])
            self.trc._check(lib.trc_ipc_open(h, ctypes.byref(self.base)), "trc_ipc_open")
        self.total = total

    def slot_ptr(self, r):
        return self.base.value + r * self.slot_bytes

    def len_ptr(self, r):
        return self.base.value + self.world * self.slot_bytes + 8 * r

    def push(self, payload: torch.Tensor, d_len_ptr: int, stream=None):
        """payload: the local packed stream buffer (16-byte aligned); d_len_ptr: device address of its uint64 length."""
        st = stream if stream is not None else torch.cuda.current_stream()
        rc = self.trc.lib.trc_push_dev(self.ct.c_void_p(self.slot_ptr(self.rank)), self.ct.c_void_p(payload.data_ptr()),
                                       self.ct.c_void_p(d_len_ptr), self.ct.c_size_t(0), self.ct.c_void_p(self.len_ptr(self.rank)),
                                       self.ct.c_void_p(st.cuda_stream))
        self.trc._check(rc, "trc_push_dev")

    def read_slot(self, r, nbytes, device):
        """(dst only) copy slot r into a fresh tensor -- verification helper."""
        out = torch.empty(nbytes, dtype=torch.uint8, device=device)
        rc = self.trc.lib.trc_memcpy_dev(self.ct.c_void_p(out.data_ptr()), self.ct.c_void_p(self.slot_ptr(r)), self.ct.c_size_t(nbytes),
                                         self.ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.trc._check(rc, "trc_memcpy_dev")
        return out

    def read_lens(self, device):
        out = torch.empty(self.world, dtype=torch.int64, device=device)
        rc = self.trc.lib.trc_memcpy_dev(self.ct.c_void_p(out.data_ptr()), self.ct.c_void_p(self.len_ptr(0)), self.ct.c_size_t(8 * self.world),
                                         self.ct.c_void_p(torch.cuda.current_stream().cuda_stream))
        self.trc._check(rc, "trc_memcpy_dev")
        return out

    def close(self):
        if self.base.value:
            if self.rank == self.dst:
                self.trc.lib.trc_dev_free(self.base)
            else:
                self.trc.lib.trc_ipc_close(self.base)
            self.base = self.ct.c_void_p()

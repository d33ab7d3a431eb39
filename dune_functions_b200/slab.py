"""z-slab partition of the element loop across ranks.

One process per GPU; rank r owns the elements with z in [z_r, z_{r+1}) and a coefficient vector
numbered over its own slab.  The only data that must travel is the lattice plane a slab shares
with its upper neighbour ("halo"): the owner of that plane is the upper slab (the last writer
in the reference's serial loop, interpolate.hh:254-259), so rank r receives its TOP plane from
rank r+1 and sends its BOTTOM plane to rank r-1.

Works with CUDA tensors over NCCL (pack/unpack kernels of the library) and with CPU tensors over
gloo (contiguous ranges from dfx_halo_ranges) — the latter is what the CPU test-suite runs.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi as capi

MAX_RANGES = 256


def halo_ranges(basis, side):
    """[(flat begin, count)] of the lowest (side 0) / highest (side 1) lattice plane"""
    b, c = (C.c_uint64 * MAX_RANGES)(), (C.c_uint64 * MAX_RANGES)()
    n = C.c_int32()
    capi.check(basis._lib.dfx_halo_ranges(basis._h, side, b, c, MAX_RANGES, C.byref(n)))
    return [(int(b[i]), int(c[i])) for i in range(n.value)]


def owned_ranges(basis):
    """[(local flat begin, count, global flat begin)] of the DOFs this slab owns"""
    lb, cnt, gb = (C.c_uint64 * MAX_RANGES)(), (C.c_uint64 * MAX_RANGES)(), (C.c_uint64 * MAX_RANGES)()
    n = C.c_int32()
    capi.check(basis._lib.dfx_owned_ranges(basis._h, lb, cnt, gb, MAX_RANGES, C.byref(n)))
    return [(int(lb[i]), int(cnt[i]), int(gb[i])) for i in range(n.value)]


def halo_size(basis):
    return int(basis._lib.dfx_halo_size(basis._h))


class HaloExchange:
    """Reusable buffers + the send-down / receive-from-above step."""

    def __init__(self, basis, rank, world, device=None, group=None):
        self.basis, self.rank, self.world, self.group = basis, rank, world, group
        self.n = halo_size(basis)
        dev = device if device is not None else "cpu"
        self.send = torch.empty(max(1, self.n), dtype=torch.float64, device=dev)
        self.recv = torch.empty(max(1, self.n), dtype=torch.float64, device=dev)

    def _pack(self, coeff, side, buf):
        if coeff.is_cuda:
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            capi.check(self.basis._lib.dfx_halo_pack(self.basis._h, side, C.c_void_p(coeff.data_ptr()),
                                                     C.c_void_p(buf.data_ptr()), st))
        else:
            o = 0
            for b, c in halo_ranges(self.basis, side):
                buf[o:o + c] = coeff[b:b + c]
                o += c

    def _unpack(self, coeff, side, buf):
        if coeff.is_cuda:
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            capi.check(self.basis._lib.dfx_halo_unpack(self.basis._h, side, C.c_void_p(buf.data_ptr()),
                                                       C.c_void_p(coeff.data_ptr()), st))
        else:
            o = 0
            for b, c in halo_ranges(self.basis, side):
                coeff[b:b + c] = buf[o:o + c]
                o += c

    def __call__(self, coeff):
        """make the top plane of every slab equal to the bottom plane of the slab above it"""
        if self.world == 1 or self.n == 0:
            return coeff
        ops = []
        if self.rank > 0:
            self._pack(coeff, 0, self.send)
            ops.append(dist.P2POp(dist.isend, self.send, self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.irecv, self.recv, self.rank + 1, self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if self.rank < self.world - 1:
            self._unpack(coeff, 1, self.recv)
        return coeff


class PeerHaloExchange:
    """The same step as HaloExchange over NVLink peer memory instead of NCCL send/recv: every rank
    owns a receive block; the upper neighbour's pack kernel stores the plane straight into it and
    raises a counter, the receiver waits for the counter ON THE STREAM (no rendezvous, no host
    synchronisation), unpacks and acknowledges (dfx_halo_link_*, dfx_halo_push / dfx_halo_pull).
    Set-up exchanges the CUDA IPC handles of the blocks through torch.distributed once."""

    def __init__(self, basis, rank, world, group=None):
        self.basis, self.rank, self.world = basis, rank, world
        self.n = halo_size(basis)
        self.step = 0
        self._link = None
        if world == 1 or self.n == 0:
            return
        L = basis._lib
        link = C.c_void_p()
        err = None
        handle = (C.c_uint8 * 64)()
        try:
            capi.check(L.dfx_halo_link_create(basis._h, C.byref(link)))
            self._link = link
            capi.check(L.dfx_halo_link_export(link, handle))
        except capi.DfxError as exc:
            err = exc
        # the collectives below are entered by every rank, whatever happened locally
        handles = [None] * world
        dist.all_gather_object(handles, None if err else bytes(handle), group=group)
        if err is None and all(h is not None for h in handles):
            try:
                if rank > 0:
                    h = (C.c_uint8 * 64).from_buffer_copy(handles[rank - 1])
                    capi.check(L.dfx_halo_link_attach(link, 0, h, None))
                if rank < world - 1:
                    h = (C.c_uint8 * 64).from_buffer_copy(handles[rank + 1])
                    capi.check(L.dfx_halo_link_attach(link, 1, h, None))
            except capi.DfxError as exc:
                err = exc
        elif err is None:
            err = capi.DfxError("a neighbouring rank could not create its halo link")
        dist.barrier(group=group)
        if err is not None:
            self.close()
            raise err

    def __call__(self, coeff):
        if self._link is None:
            return coeff
        self.step += 1
        L = self.basis._lib
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        if self.rank > 0:
            capi.check(L.dfx_halo_push(self.basis._h, self._link, C.c_void_p(coeff.data_ptr()), self.step, st))
        if self.rank < self.world - 1:
            capi.check(L.dfx_halo_pull(self.basis._h, self._link, C.c_void_p(coeff.data_ptr()), self.step, st))
        return coeff

    def error(self):
        return 0 if self._link is None else int(self.basis._lib.dfx_halo_link_error(self._link))

    def close(self):
        if self._link is not None:
            self.basis._lib.dfx_halo_link_destroy(self._link)
            self._link = None


class GlobalGather:
    """Gather of the distributed coefficient vector: every z-slab owns one
    contiguous flat-position range per (leaf, index-set component) and that range is contiguous in the
    GLOBAL numbering too (dfx_owned_ranges), so the gather is a handful of range transfers per pair of
    ranks, received straight into their place in the output — no staging buffer, no padding, no index
    lists.  The metadata (every rank's ranges) is exchanged once at construction.
    root=None: every rank ends up with the whole vector (all-gather); root=r: only rank r does."""

    def __init__(self, basis, global_dimension, group=None, root=None):
        self.group, self.root, self.n = group, root, int(global_dimension)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.mine = owned_ranges(basis)
        allr = [None] * self.world
        dist.all_gather_object(allr, self.mine, group=group)
        self.ranges = allr
        covered = sum(c for r in allr for _, c, _ in r)
        if covered != self.n:
            raise capi.RangeError(f"owned ranges cover {covered} of {self.n} global entries")

    def bytes_received(self):
        """bytes this rank receives from its peers per gather"""
        if self.root is not None and self.rank != self.root:
            return 0
        return 8 * sum(c for r, rr in enumerate(self.ranges) if r != self.rank for _, c, _ in rr)

    def __call__(self, coeff, out=None):
        receiver = self.root is None or self.rank == self.root
        if receiver and out is None:
            out = torch.empty(self.n, dtype=coeff.dtype, device=coeff.device)
        ops = []
        # the same (sender, range) order on both sides of every pair
        for dst in range(self.world):
            if dst == self.rank or (self.root is not None and dst != self.root):
                continue
            for lb, c, _ in self.mine:
                if c:
                    ops.append(dist.P2POp(dist.isend, coeff[lb:lb + c], dst, self.group))
        if receiver:
            for src in range(self.world):
                if src == self.rank:
                    continue
                for _, c, gb in self.ranges[src]:
                    if c:
                        ops.append(dist.P2POp(dist.irecv, out[gb:gb + c], src, self.group))
            for lb, c, gb in self.mine:
                out[gb:gb + c].copy_(coeff[lb:lb + c])
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out


def gather_global(basis, coeff, global_dimension, group=None, root=None):
    """one-shot form of GlobalGather: the whole vector in GLOBAL flat order on every rank (or on `root`)"""
    return GlobalGather(basis, global_dimension, group=group, root=root)(coeff)

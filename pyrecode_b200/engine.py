"""Batch engines over the C ABI: device buffers (torch), pinned staging, launches, D2H of compact results.

WriteEngine  = the per-frame hot path of ReCoDeWriter._reduce_compress (reference: pyrecode/recode_writer.py:430-557)
               for a batch of frames.
ReadEngine   = ReCoDeReader._get_frame_sparse (reference: pyrecode/recode_reader.py:379-462) for a batch of frames:
               inflate -> unpack to triples / dense frames / summed live-view image.

PyTorch only owns memory and streams here; all arithmetic is in librecode_b200.so.
"""
import numpy as np
import torch

from . import _native
from ._native import Context, make_config


class _Slot:
    """One in-flight batch of the write path: its own rc_ctx, CUDA stream, workspace and result buffers."""

    def __init__(self, eng, device):
        self.ctx = Context(device)
        dev = self.ctx.device
        with torch.cuda.device(dev):
            self.stream = torch.cuda.Stream(device=dev)
            self.ws = self.ctx.empty(self.ctx.workspace_bytes(eng.cfg))
            self.records = self.ctx.empty(eng.records_cap)
            self.offsets = self.ctx.zeros(eng.max_frames + 1, torch.int64)
            self.counts = self.ctx.zeros(eng.max_frames, torch.int32)
            self.status = self.ctx.zeros(1, torch.int32)
            self.done = torch.cuda.Event()
        self.frames_dev = None
        self.pin_in = None
        self.pin_off = torch.empty(eng.max_frames + 1, dtype=torch.int64).pin_memory()
        self.pin_cnt = torch.empty(eng.max_frames, dtype=torch.int32).pin_memory()
        self.pin_st = torch.empty(1, dtype=torch.int32).pin_memory()
        self.pin_rec = None
        self.n = 0
        self.h2d = 0
        self.busy = False


class WriteEngine:
    """Batches of frames -> finished part-file records.  `n_slots` batches can be in flight: submit() enqueues the
    host-to-device copy and all kernels of a batch on the slot's stream and returns at once, collect() waits for
    that batch and brings its records to the host, so the copies and the file write of one batch overlap the
    kernels of the next."""

    def __init__(self, ny, nx, itemsize, bit_depth, reduction_level, rc_operation_mode=1, l2_statistics=0,
                 l4_centroiding=0, compression_level=1, max_frames=16, device=None, records_capacity=None, n_slots=1):
        self.ny, self.nx, self.itemsize, self.bit_depth = int(ny), int(nx), int(itemsize), int(bit_depth)
        self.level, self.mode = int(reduction_level), int(rc_operation_mode)
        self.max_frames = int(max_frames)
        self.cfg = make_config(ny, nx, itemsize, bit_depth, reduction_level, rc_operation_mode, l2_statistics,
                               l4_centroiding, compression_level, max_frames)
        self.P = self.ny * self.nx
        self.np_dtype = _native.numpy_dtype(itemsize)
        self.t_dtype = _native.torch_dtype(itemsize)
        probe = Context(device)
        self.records_cap = probe.records_capacity(self.cfg) if records_capacity is None else int(records_capacity)
        probe.close()
        self.slots = [_Slot(self, device) for _ in range(max(1, int(n_slots)))]
        if len(self.slots) > 1:
            for sl in self.slots:                 # several batches in flight: let them share the SMs (rc_set_pipelined)
                sl.ctx.set_pipelined(True)
        self._next = 0
        s0 = self.slots[0]
        self.ctx, self.dev = s0.ctx, s0.ctx.device
        # single-slot views used by the stage API and older callers
        self.ws, self.records, self.offsets, self.counts, self.status = s0.ws, s0.records, s0.offsets, s0.counts, s0.status
        self.thr = None

    # ---- calibration -----------------------------------------------------------------------------
    def set_threshold(self, dark, eps):
        """thr = dark + eps in the source dtype (recode_writer.py:126-137), computed on the device."""
        d = torch.from_numpy(np.ascontiguousarray(dark, dtype=self.np_dtype)).to(self.dev)
        with torch.cuda.device(self.dev):
            self.thr = self.ctx.make_threshold(self.cfg, d, eps)
            torch.cuda.current_stream().synchronize()
        return self.thr

    # ---- input staging ----------------------------------------------------------------------------
    def _to_device(self, frames, slot=None):
        """frames: numpy [n, ny, nx], pinned/pageable torch CPU tensor, or CUDA tensor of the source dtype.
        The copy is enqueued on the current stream."""
        slot = slot or self.slots[0]
        if isinstance(frames, torch.Tensor) and frames.is_cuda:
            assert frames.dtype == self.t_dtype and frames.is_contiguous()
            return frames, 0
        if isinstance(frames, np.ndarray):
            a = np.ascontiguousarray(frames, dtype=self.np_dtype)
            n = a.shape[0]
            if slot.pin_in is None:
                slot.pin_in = torch.empty((self.max_frames, self.ny, self.nx), dtype=self.t_dtype).pin_memory()
            slot.pin_in[:n].numpy()[...] = a
            src = slot.pin_in[:n]
        else:
            src = frames
            n = src.shape[0]
        if slot.frames_dev is None:
            slot.frames_dev = torch.empty((self.max_frames, self.ny, self.nx), dtype=self.t_dtype, device=self.dev)
        dst = slot.frames_dev[:n]
        dst.copy_(src, non_blocking=True)
        return dst, src.numel() * src.element_size()

    def pinned_input(self, n):
        """-> pinned CPU tensor [n, ny, nx] of the slot the next submit() will use: a source that can fill it in place
        (a file read) saves the staging copy.  The slot must have been collected."""
        s = self.slots[self._next]
        if s.busy:
            raise RuntimeError('slot %d still holds an uncollected batch' % self._next)
        if s.pin_in is None:
            s.pin_in = torch.empty((self.max_frames, self.ny, self.nx), dtype=self.t_dtype).pin_memory()
        return s.pin_in[:n]

    # ---- the hot path ------------------------------------------------------------------------------
    def launch(self, frames_dev, n, first_frame_id, slot=0):
        """Asynchronous: all kernels of one batch on the CURRENT stream, using the buffers of `slot`."""
        if self.thr is None:
            raise RuntimeError('set_threshold() must be called first')
        s = self.slots[slot]
        s.ctx.reduce_compress(self.cfg, frames_dev, n, self.thr, int(first_frame_id), s.ws, s.records, s.offsets,
                              s.counts, s.status)

    def submit(self, frames, first_frame_id=0):
        """Enqueue one batch (copy in, kernels, copy of the record table out) on the next slot's stream.
        Returns the slot index to pass to collect()."""
        n = int(frames.shape[0])
        if n > self.max_frames:
            raise ValueError('batch larger than max_frames')
        k = self._next
        s = self.slots[k]
        if s.busy:
            raise RuntimeError('slot %d still holds an uncollected batch' % k)
        self._next = (k + 1) % len(self.slots)
        with torch.cuda.device(self.dev):
            if isinstance(frames, torch.Tensor) and frames.is_cuda:
                s.stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s.stream):
                fd, s.h2d = self._to_device(frames, s)
                self.launch(fd, n, first_frame_id, k)
                s.pin_off[:n + 1].copy_(s.offsets[:n + 1], non_blocking=True)
                s.pin_cnt[:n].copy_(s.counts[:n], non_blocking=True)
                s.pin_st.copy_(s.status, non_blocking=True)
                s.done.record(s.stream)
        s.n = n
        s.busy = True
        return k

    def collect(self, k):
        """-> (records: np.uint8 [total] (valid until the slot is reused), offsets: np.int64 [n+1],
        counts: np.int32 [n], h2d_bytes, d2h_bytes) of the batch submitted on slot k."""
        s = self.slots[k]
        n = s.n
        s.done.synchronize()
        s.busy = False
        if int(s.pin_st[0]) & _native.RC_STATUS_RECORDS_OVERFLOW:
            raise ValueError('Buffer size smaller than compressed data size')
        offs = s.pin_off[:n + 1].numpy().copy()
        counts = s.pin_cnt[:n].numpy().copy()
        total = int(offs[n])
        if s.pin_rec is None or s.pin_rec.numel() < total:
            s.pin_rec = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory()
        with torch.cuda.device(self.dev), torch.cuda.stream(s.stream):
            s.pin_rec[:total].copy_(s.records[:total], non_blocking=True)
            s.stream.synchronize()
        return s.pin_rec[:total].numpy(), offs, counts, s.h2d, total + (n + 1) * 8 + n * 4 + 4

    def reduce_compress(self, frames, first_frame_id=0):
        """Synchronous submit + collect of one batch."""
        return self.collect(self.submit(frames, first_frame_id))

    # ---- stage access (parity tests, c_recode shim) --------------------------------------------------
    def reduce(self, frames):
        """-> (maps: [bytes], packed: [bytes], counts) through rc_reduce."""
        n = int(frames.shape[0])
        with torch.cuda.device(self.dev):
            fd, _ = self._to_device(frames)
            ms = self.ctx.map_stride_words(self.P)
            ps = self.ctx.packed_stride_bytes(self.cfg)
            maps = self.ctx.zeros(n * ms + 16, torch.int32)
            packed = self.ctx.zeros(n * ps + 16)
            pbytes = self.ctx.zeros(n, torch.int32)
            counts = self.ctx.zeros(n, torch.int32)
            self.ctx.reduce(self.cfg, fd, n, self.thr, self.ws, maps, packed, pbytes, counts)
            torch.cuda.synchronize()
            mh = maps.cpu().numpy().view(np.uint8)
            ph = packed.cpu().numpy()
            pb = pbytes.cpu().numpy()
            mb = (self.P + 7) // 8
            out_m = [mh[f * ms * 4:f * ms * 4 + mb].tobytes() for f in range(n)]
            out_p = [ph[f * ps:f * ps + int(pb[f])].tobytes() for f in range(n)]
            return out_m, out_p, counts.cpu().numpy()

    def _stage_workspace(self):
        """workspace of the stage entry points rc_ccl_label / rc_l4_centroids (larger than the write path's)"""
        if getattr(self, '_stage_ws', None) is None:
            self._stage_ws = self.ctx.empty(self.ctx.stage_workspace_bytes(self.cfg))
        return self._stage_ws

    def labels(self, maps_bytes):
        """8-connected labels of packed binary maps -> (int32 [n, ny, nx], k[n]) through rc_ccl_label."""
        n = len(maps_bytes)
        ms = self.ctx.map_stride_words(self.P)
        host = np.zeros((n, ms * 4), dtype=np.uint8)
        for f, m in enumerate(maps_bytes):
            host[f, :len(m)] = np.frombuffer(m, dtype=np.uint8)
        with torch.cuda.device(self.dev):
            maps = torch.from_numpy(host.view(np.int32).reshape(-1)).to(self.dev)
            labels = self.ctx.empty(n * self.P, torch.int32)
            counts = self.ctx.zeros(n, torch.int32)
            self.ctx.ccl_label(self.cfg, maps, n, self._stage_workspace(), labels, counts)
            torch.cuda.synchronize()
            return labels.cpu().numpy().reshape(n, self.ny, self.nx), counts.cpu().numpy()

    def centroids(self, frames):
        """L4 centroid lists -> [float32 [k, 2]] through rc_l4_centroids."""
        n = int(frames.shape[0])
        cap = (self.ny + 1) // 2 * ((self.nx + 1) // 2) + 1
        with torch.cuda.device(self.dev):
            fd, _ = self._to_device(frames)
            cent = self.ctx.zeros(n * cap * 2, torch.float32)
            counts = self.ctx.zeros(n, torch.int32)
            self.ctx.l4_centroids(self.cfg, fd, n, self.thr, self._stage_workspace(), cent, cap, counts)
            torch.cuda.synchronize()
            c = cent.cpu().numpy().reshape(n, cap, 2)
            k = counts.cpu().numpy()
            return [c[f, :int(k[f])].copy() for f in range(n)]


def deflate_batch(ctx, payloads, level=1):
    """zlib-format deflate of a list of byte strings on the GPU -> list of bytes (rc_deflate_zlib)."""
    n = len(payloads)
    if n == 0:
        return []
    sizes = np.array([len(p) for p in payloads], dtype=np.uint32)
    offs = np.zeros(n, dtype=np.uint64)
    pos = 0
    for i, s in enumerate(sizes):
        offs[i] = pos
        pos += (int(s) + 15) // 16 * 16
    blob = np.zeros(pos + 64, dtype=np.uint8)
    for i, p in enumerate(payloads):
        blob[int(offs[i]):int(offs[i]) + len(p)] = np.frombuffer(p, dtype=np.uint8)
    mx = int(sizes.max())
    stride = (ctx.deflate_bound(mx) + 15) // 16 * 16
    with torch.cuda.device(ctx.device):
        d_in = torch.from_numpy(blob).to(ctx.device)
        d_off = torch.from_numpy(offs.view(np.int64)).to(ctx.device)
        d_sz = torch.from_numpy(sizes.view(np.int32)).to(ctx.device)
        out = ctx.empty(n * stride + 64)
        out_bytes = ctx.zeros(n, torch.int32)
        ws = ctx.deflate_zlib(level, d_in, d_off, d_sz, n, mx, out, stride, out_bytes)
        torch.cuda.synchronize()
        del ws
        ob = out_bytes.cpu().numpy()
        oh = out.cpu().numpy()
    return [oh[i * stride:i * stride + int(ob[i])].tobytes() for i in range(n)]


def _stage_streams(ctx, streams):
    n = len(streams)
    sizes = np.array([len(p) for p in streams], dtype=np.uint32)
    offs = np.zeros(n, dtype=np.uint64)
    pos = 0
    for i, s in enumerate(sizes):
        offs[i] = pos
        pos += (int(s) + 15) // 16 * 16
    blob = np.zeros(pos + 64, dtype=np.uint8)
    for i, p in enumerate(streams):
        blob[int(offs[i]):int(offs[i]) + len(p)] = np.frombuffer(p, dtype=np.uint8)
    d_in = torch.from_numpy(blob).to(ctx.device, non_blocking=False)
    d_off = torch.from_numpy(offs.view(np.int64)).to(ctx.device)
    d_sz = torch.from_numpy(sizes.view(np.int32)).to(ctx.device)
    return d_in, d_off, d_sz, int(pos)


def inflate_batch(ctx, streams, out_capacity):
    """zlib-format inflate of a list of byte strings on the GPU -> (list of bytes, status[n]) (rc_inflate_zlib)."""
    n = len(streams)
    if n == 0:
        return [], np.zeros(0, np.uint32)
    stride = (int(out_capacity) + 15) // 16 * 16 + 16
    with torch.cuda.device(ctx.device):
        d_in, d_off, d_sz, _ = _stage_streams(ctx, streams)
        out = ctx.empty(n * stride + 64)
        out_bytes = ctx.zeros(n, torch.int32)
        status = ctx.zeros(n, torch.int32)
        ws = ctx.inflate_zlib(d_in, d_off, d_sz, n, out, stride, out_bytes, status)
        torch.cuda.synchronize()
        ob = out_bytes.cpu().numpy()
        st = status.cpu().numpy().astype(np.uint32)
        del ws
        oh = out.cpu().numpy()
    return [oh[i * stride:i * stride + int(ob[i])].tobytes() for i in range(n)], st


class ReadEngine:
    def __init__(self, ny, nx, itemsize, bit_depth, reduction_level, rc_operation_mode=1, max_frames=16, device=None):
        self.ctx = Context(device)
        self.dev = self.ctx.device
        self.ny, self.nx, self.itemsize, self.bit_depth = int(ny), int(nx), int(itemsize), int(bit_depth)
        self.level, self.mode = int(reduction_level), int(rc_operation_mode)
        self.max_frames = int(max_frames)
        self.P = self.ny * self.nx
        self.cfg = make_config(ny, nx, itemsize, bit_depth, reduction_level, rc_operation_mode, 0, 0, 1, max_frames)
        self.np_dtype = _native.numpy_dtype(itemsize)
        self.t_dtype = _native.torch_dtype(itemsize)
        self.map_bytes = (self.P + 7) // 8
        self.ms = self.ctx.map_stride_words(self.P)
        # the streams are inflated where the unpack kernels read them: binary maps at the library's map stride (whole
        # tiles of 1024 words per frame, the padding stays zero), value streams at their own stride -- no repacking
        self.mstride = self.ms * 4
        self.stride = ((self.P * self.bit_depth + 7) // 8 + 15) // 16 * 16 + 16 if self.level <= 2 else 16
        F = self.max_frames
        with torch.cuda.device(self.dev):
            self.ws = self.ctx.empty(self.ctx.read_workspace_bytes(self.cfg))
            self.maps_buf = self.ctx.zeros(F * self.mstride + 64)
            self.packed_buf = self.ctx.zeros(F * self.stride + 64)
            self.out_bytes = self.ctx.zeros(2 * F, torch.int32)
            self.status = self.ctx.zeros(2 * F, torch.int32)
            self.counts = self.ctx.zeros(F, torch.int32)
            self._inf_ws = None
            self._inf_ws2 = None
            if self.mode == 1:
                self._inf_ws = self.ctx.empty(self.ctx._lib.rc_inflate_workspace_bytes(F, self.mstride))
                if self.level <= 2:
                    self._inf_ws2 = self.ctx.empty(self.ctx._lib.rc_inflate_workspace_bytes(F, self.stride))

    def load(self, map_streams, val_streams):
        """Stage n frames' streams (compressed when mode 1, raw when mode 0) -> device maps / packed values."""
        n = len(map_streams)
        if n > self.max_frames:
            raise ValueError('batch larger than max_frames')
        has_vals = val_streams is not None and self.level <= 2
        F = self.max_frames
        with torch.cuda.device(self.dev):
            if self.mode == 1:
                streams = list(map_streams) + (list(val_streams) if has_vals else [])
                d_in, d_off, d_sz, staged = _stage_streams(self.ctx, streams)
                ns = len(streams)
                # maps occupy stream slots [0, n), values [F, F + n) of out_bytes / status
                self._inf_ws = self.ctx.inflate_zlib(d_in, d_off, d_sz, n, self.maps_buf, self.mstride,
                                                     self.out_bytes, self.status, self._inf_ws)
                if has_vals:
                    self._inf_ws2 = self.ctx.inflate_zlib(d_in, d_off[n:], d_sz[n:], n, self.packed_buf, self.stride,
                                                          self.out_bytes[F:], self.status[F:], self._inf_ws2)
                self._keep = (d_in, d_off, d_sz)
                h2d = staged + ns * 12
            else:
                hm = np.zeros((F, self.mstride), dtype=np.uint8)
                hp = np.zeros((F, self.stride), dtype=np.uint8)
                for f in range(n):
                    if len(map_streams[f]) > self.mstride or (has_vals and len(val_streams[f]) > self.stride):
                        raise ValueError('stream of frame %d is longer than a frame can produce' % f)
                    hm[f, :len(map_streams[f])] = np.frombuffer(map_streams[f], dtype=np.uint8)
                    if has_vals:
                        hp[f, :len(val_streams[f])] = np.frombuffer(val_streams[f], dtype=np.uint8)
                self.maps_buf[:F * self.mstride].copy_(torch.from_numpy(hm.reshape(-1)))
                self.packed_buf[:F * self.stride].copy_(torch.from_numpy(hp.reshape(-1)))
                self.status.zero_()
                ob = np.zeros(2 * F, dtype=np.int32)
                ob[:n] = [len(m) for m in map_streams]
                if has_vals:
                    ob[F:F + n] = [len(v) for v in val_streams]
                self.out_bytes.copy_(torch.from_numpy(ob))
                h2d = hm.nbytes + hp.nbytes
        self.n = n
        self.has_vals = has_vals
        return h2d

    # ---- block staging: the reader reads a batch of records from the file straight into pinned memory
    def block_buffer(self, nbytes, keep=0):
        """-> numpy uint8 view of at least nbytes of pinned host memory (the first `keep` bytes survive a regrow)"""
        cur = getattr(self, '_h_block', None)
        if cur is None or cur.numel() < nbytes:
            cap = max(int(nbytes), 2 * (cur.numel() if cur is not None else 0), 16 << 20)
            new = torch.empty(cap, dtype=torch.uint8).pin_memory()
            if cur is not None and keep:
                new[:keep].copy_(cur[:keep])
            self._h_block = new
            with torch.cuda.device(self.dev):
                self._d_block = self.ctx.empty(cap + 64)
        return self._h_block.numpy()

    def wait_block_free(self):
        """the pinned block may be overwritten again once its host-to-device copy has finished"""
        ev = getattr(self, '_h2d_done', None)
        if ev is not None:
            ev.synchronize()

    def load_block(self, nbytes, map_off, map_sz, val_off=None, val_sz=None, maps_only=False):
        """Inflate n frames whose compressed streams lie in the pinned block (block_buffer) at the given byte
        offsets.  Everything is enqueued on the current CUDA stream; nothing synchronizes."""
        if self.mode != 1:
            raise ValueError('load_block handles compressed records (rc_operation_mode 1) only')
        n = len(map_off)
        if n > self.max_frames:
            raise ValueError('batch larger than max_frames')
        has_vals = val_off is not None and self.level <= 2
        F = self.max_frames
        with torch.cuda.device(self.dev):
            meta = getattr(self, '_h_meta', None)
            if meta is None:
                self._h_meta = meta = torch.empty(2 * F * 3, dtype=torch.int32).pin_memory()
                self._d_meta = self.ctx.empty(2 * F * 3 * 4).view(torch.int32)
            mv = meta.numpy()
            offs = mv[:4 * F].view(np.int64)           # [2F] stream offsets, then [2F] int32 stream sizes
            sizes = mv[4 * F:]
            offs[:n] = map_off
            sizes[:n] = map_sz
            if has_vals:
                offs[F:F + n] = val_off
                sizes[F:F + n] = val_sz
            d_blk = self._d_block
            d_blk[:nbytes].copy_(self._h_block[:nbytes], non_blocking=True)
            self._d_meta.copy_(meta, non_blocking=True)
            self._h2d_done = torch.cuda.Event()
            self._h2d_done.record()
            d_off = self._d_meta[:4 * F].view(torch.int64)
            d_sz = self._d_meta[4 * F:]
            self._inf_ws = self.ctx.inflate_zlib(d_blk, d_off, d_sz, n, self.maps_buf, self.mstride,
                                                 self.out_bytes, self.status, self._inf_ws)
            if has_vals and not maps_only:
                self._inf_ws2 = self.ctx.inflate_zlib(d_blk, d_off[F:], d_sz[F:], n, self.packed_buf, self.stride,
                                                      self.out_bytes[F:], self.status[F:], self._inf_ws2)
        self.n = n
        self.has_vals = has_vals
        return int(nbytes) + meta.numel() * 4

    def check(self, packed_sizes=None):
        """Host-side validation after load(): stream status and sizes (raises ValueError like zlib.error would).
        packed_sizes: the bytes_in_packed_* metadata of the n frames -- the value stream must inflate to exactly that
        many bytes (a valid but short stream would otherwise read as zeros out of the cleared output buffer)."""
        F, n = self.max_frames, self.n
        st = self.status.cpu().numpy()
        ob = self.out_bytes.cpu().numpy()
        for f in range(n):
            if st[f] or (self.has_vals and st[F + f]):
                raise ValueError('corrupt compressed stream in frame %d (status %d/%d)' % (f, st[f], st[F + f]))
            if ob[f] != self.map_bytes:
                raise ValueError('binary map of frame %d inflates to %d bytes, expected %d' % (f, ob[f], self.map_bytes))
            if self.has_vals and packed_sizes is not None and int(ob[F + f]) != int(packed_sizes[f]):
                raise ValueError('value stream of frame %d inflates to %d bytes, the record says %d'
                                 % (f, ob[F + f], int(packed_sizes[f])))
        return ob

    def sparse(self):
        """-> list of uint64 [n_fg, 3] (row, col, value) arrays, the c_recode.get_frame_sparse result."""
        n = self.n
        with torch.cuda.device(self.dev):
            cap = self.P
            tri = self.ctx.empty(n * cap * 3, torch.int64)
            self.ctx.unpack_sparse(self.cfg, self.maps_buf, self.packed_buf, self.stride, n, self.ws, tri, cap, self.counts)
            torch.cuda.synchronize()
            k = self.counts[:n].cpu().numpy()
            out = []
            for f in range(n):
                t = tri[f * cap * 3:f * cap * 3 + int(k[f]) * 3].cpu().numpy().view(np.uint64).reshape(-1, 3)
                out.append(t)
            return out

    def sparse_coo(self):
        """-> list of (rows int32, cols int32, values of the target dtype), one triple of numpy arrays per loaded frame:
        the arrays scipy's coo_matrix keeps.  Same kernel as sparse(); the uint64 triples are narrowed on the device, so
        10 instead of 24 bytes per foreground pixel cross PCIe, in three copies per batch instead of one per frame."""
        n = self.n
        with torch.cuda.device(self.dev):
            cap = self.P
            tri = self.ctx.empty(n * cap * 3, torch.int64)
            self.ctx.unpack_sparse(self.cfg, self.maps_buf, self.packed_buf, self.stride, n, self.ws, tri, cap, self.counts)
            k = self.counts[:n].cpu().numpy().astype(np.int64)
            total = int(k.sum())
            vt = torch.int16 if self.itemsize == 2 else torch.uint8       # int16: the bits of the uint16 values
            rows_d = torch.empty(max(total, 1), dtype=torch.int32, device=self.dev)
            cols_d = torch.empty(max(total, 1), dtype=torch.int32, device=self.dev)
            vals_d = torch.empty(max(total, 1), dtype=vt, device=self.dev)
            off = 0
            for f in range(n):
                kf = int(k[f])
                if kf:
                    t = tri[f * cap * 3:f * cap * 3 + kf * 3].view(kf, 3)
                    rows_d[off:off + kf].copy_(t[:, 0])
                    cols_d[off:off + kf].copy_(t[:, 1])
                    vals_d[off:off + kf].copy_(t[:, 2])
                off += kf
            rows = rows_d.cpu().numpy()
            cols = cols_d.cpu().numpy()
            vals = vals_d.cpu().numpy().view(self.np_dtype)
        out, off = [], 0
        for f in range(n):
            kf = int(k[f])
            out.append((rows[off:off + kf], cols[off:off + kf], vals[off:off + kf]))
            off += kf
        return out

    def summary_stats(self, out_bytes):
        """L2: unpack the per-puddle statistics of the loaded frames -> list of arrays of the target dtype
        (intent of recode_reader.py:473-481, count = bytes * 8 // bit_depth)."""
        F, n = self.max_frames, self.n
        packed = self.packed_buf
        ks = [int(out_bytes[F + f]) * 8 // self.bit_depth for f in range(n)]
        total = sum(ks)
        with torch.cuda.device(self.dev):
            out = self.ctx.zeros(max(total, 1), torch.int64)
            off = 0
            for f in range(n):                      # one launch per frame, one copy to the host for the batch
                if ks[f]:
                    self.ctx.bit_unpack(self.bit_depth, packed[f * self.stride:], ks[f], out[off:])
                off += ks[f]
            host = out.cpu().numpy().astype(self.np_dtype)
        res, off = [], 0
        for f in range(n):
            res.append(host[off:off + ks[f]])
            off += ks[f]
        return res

    def dense(self, total=None, want_dense=True, out=None):
        """-> dense frames [n, ny, nx] on the device (written into `out` when given: a preallocated tensor of the target
        dtype with room for n frames) and/or accumulates into `total`, uint32 [ny*nx]."""
        n = self.n
        with torch.cuda.device(self.dev):
            dense = None
            if want_dense:
                dense = out[:n] if out is not None else torch.empty((n, self.ny, self.nx), dtype=self.t_dtype,
                                                                    device=self.dev)
            self.ctx.unpack_dense(self.cfg, self.maps_buf, self.packed_buf, self.stride, n, self.ws, dense, total,
                                  self.counts)
        return dense

"""`c_recode`-compatible shim: the reference's native extension type (pyrecode/pyrecode.cpp:161-200) with the same
method names and argument meaning, running on the GPU through librecode_b200.so.  Keeps third-party callers of
`c_recode.Reader` working; the reference's defects are fixed behind the same signatures (SURVEY B-1, B-3).

    Reader().create_buffers(ny, nx, bit_depth)                             pyrecode.cpp:57-72
    Reader().get_frame_sparse(level, binary_map, packed_vals, out)        pyrecode.cpp:95-119
    Reader().bit_unpack_pixel_intensities(n_values, packed, out)          pyrecode.cpp:74-93
    Reader().bit_pack_pixel_intensities(sz_packed, n_fg, bit_depth, vals, out)   pyrecode.cpp:121-141
"""
import time

import numpy as np


class Reader:

    def __init__(self, device=None):
        """device: CUDA device index of this Reader's context and buffers (an extension; the reference takes no
        arguments, pyrecode.cpp:41-55); None = the current device"""
        self._device = device
        self.ny = 0
        self.nx = 0
        self.bit_depth = 0
        self._engines = {}
        self._ctx = None

    def create_buffers(self, ny, nx, bit_depth):
        self.ny, self.nx, self.bit_depth = int(ny), int(nx), int(bit_depth)
        return 1

    def _context(self):
        if self._ctx is None:
            from ._native import Context
            self._ctx = Context(self._device)
        return self._ctx

    def _engine(self, level):
        if level not in self._engines:
            from .engine import ReadEngine
            itemsize = 1 if self.bit_depth <= 8 else 2
            self._engines[level] = ReadEngine(self.ny, self.nx, itemsize, self.bit_depth, level,
                                              rc_operation_mode=0, max_frames=1, device=self._device)
        return self._engines[level]

    def get_frame_sparse(self, reduction_level, binary_map, packed_vals, out):
        """fills `out` (writable buffer viewed as uint64) with (row, col, value) triples, returns n foreground.
        As in the reference, only level 1 reads `packed_vals`; every other level emits value 1 for each set map bit
        and ignores the second stream (reader.h:39-41) -- L2 statistics are unpacked by
        bit_unpack_pixel_intensities."""
        level = 1 if reduction_level == 1 else 3
        eng = self._engine(level)
        eng.load([bytes(binary_map)], [bytes(packed_vals)] if level == 1 and packed_vals is not None else None)
        eng.check()                                   # a map of the wrong length raises instead of reading zeros
        tri = eng.sparse()[0]
        dst = np.frombuffer(out, dtype=np.uint64)
        dst[:tri.size] = tri.ravel()
        return int(tri.shape[0])

    def bit_unpack_pixel_intensities(self, n_values, packed, out):
        import torch
        ctx = self._context()
        n = int(n_values)
        src = np.frombuffer(bytes(packed) + b'\0' * 8, dtype=np.uint8)
        d_in = torch.from_numpy(src.copy()).to(ctx.device)
        d_out = ctx.zeros(max(n, 1), torch.int64)
        if n:
            ctx.bit_unpack(self.bit_depth, d_in, n, d_out)
        np.frombuffer(out, dtype=np.uint64)[:n] = d_out[:n].cpu().numpy().view(np.uint64)
        return n

    def bit_pack_pixel_intensities(self, sz_packed, n_fg_pixels, bit_depth, pixvals, packed_pixvals):
        """packs n_fg_pixels uint16 values at bit_depth bits each into packed_pixvals[:sz_packed]; returns ms"""
        import torch
        t0 = time.perf_counter()
        ctx = self._context()
        n = int(n_fg_pixels)
        vals = np.frombuffer(pixvals, dtype=np.uint16, count=n)
        d_vals = torch.from_numpy(vals.copy()).to(ctx.device)
        nb = (n * int(bit_depth) + 7) // 8
        d_out = ctx.zeros((nb + 7) // 4 * 4)
        if n:
            ctx.bit_pack(int(bit_depth), d_vals, n, d_out)
        dst = np.frombuffer(packed_pixvals, dtype=np.uint8)
        dst[:int(sz_packed)] = 0
        dst[:nb] = d_out[:nb].cpu().numpy()
        return (time.perf_counter() - t0) * 1000.0

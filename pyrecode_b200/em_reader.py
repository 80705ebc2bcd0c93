"""SEQ source files for the streaming ingest: a pure-NumPy reader of the Norpix StreamPix
sequence layout that the reference reads through `pims.open` (pyrecode/em_reader.py:243-304; pims is a
requirements.txt dependency that is not installed here, so the published layout is restated:

    bytes 0..1023     header, little-endian: magic 0xFEED, name 'Norpix seq\\n' (UTF-16), version (i32 @28),
                      header_size (i32 @32), description (512 B), width, height, bit_depth, bit_depth_real,
                      image_size_bytes, image_format (6 x u32 @548), allocated_frames (@572), origin (@576),
                      true_image_size (@580: the stride from one frame to the next, image + 8-byte timestamp + padding),
                      suggested_frame_rate (f64 @584), ...
    image_offset      1024 for version < 5, 8192 for StreamPix 6 (version >= 5)
    frame i           width * height pixels at image_offset + i * true_image_size, followed by its timestamp
                      (u32 seconds, u16 milliseconds, u16 microseconds)

"parity unpinned": no reference test or fixture holds a .seq file).  Only what the writer's stream mode needs is here:
shape, dtype, frame ranges straight into a caller-supplied (pinned) buffer.  The acquisition writes the chunk that is
being read, so the frame count comes from the file size, not from `allocated_frames` (the reference falls back to
frame-by-frame reads on IndexError for the same reason, pyrecode/recode_writer.py:330-347).
"""
import os
import struct

import numpy as np

SEQ_MAGIC = 0xFEED
SEQ_HEADER_BYTES = 1024
_NAME = 'Norpix seq\n'.encode('utf-16-le')


class SEQReader:
    """`with SEQReader(path) as f: f[a:b]` -> uint8 / uint16 array [n, ny, nx]; `f.read_into(buf, a, b)` fills a
    caller-owned array (e.g. the numpy view of a pinned torch tensor) without an intermediate copy."""

    def __init__(self, file):
        self._source_filename = file
        self._fp = open(file, 'rb')
        h = self._fp.read(SEQ_HEADER_BYTES)
        if len(h) < SEQ_HEADER_BYTES or struct.unpack_from('<L', h, 0)[0] != SEQ_MAGIC:
            self._fp.close()
            raise ValueError('%s is not a Norpix sequence file' % file)
        hd = {}
        hd['version'], hd['header_size'] = struct.unpack_from('<ll', h, 28)
        (hd['width'], hd['height'], hd['bit_depth'], hd['bit_depth_real'], hd['image_size_bytes'],
         hd['image_format']) = struct.unpack_from('<6L', h, 548)
        hd['allocated_frames'], hd['origin'], hd['true_image_size'] = struct.unpack_from('<3L', h, 572)
        hd['suggested_frame_rate'] = struct.unpack_from('<d', h, 584)[0]
        self._header = hd
        self.header_dict = hd
        if hd['bit_depth'] == 8:
            self._dtype = np.uint8
        elif hd['bit_depth'] == 16:
            self._dtype = np.uint16
        else:
            self._fp.close()
            raise TypeError('Sequence datasets with bit-depth %d is not supported.' % hd['bit_depth'])
        self._image_offset = 8192 if hd['version'] >= 5 else SEQ_HEADER_BYTES
        self._frame_bytes = hd['width'] * hd['height'] * np.dtype(self._dtype).itemsize
        self._stride = hd['true_image_size'] or self._frame_bytes
        if self._stride < self._frame_bytes:
            self._fp.close()
            raise ValueError('true_image_size %d smaller than one frame (%d bytes)' % (self._stride, self._frame_bytes))

    # ---- the subset of EMReaderBase the writer uses (pyrecode/em_reader.py:38-130) -------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        self._fp.close()

    @property
    def dtype(self):
        return self._dtype

    @property
    def header(self):
        return self._header

    def frames_on_disk(self):
        """whole frames present in the file right now"""
        size = os.fstat(self._fp.fileno()).st_size
        if size < self._image_offset + self._frame_bytes:
            return 0
        return (size - self._image_offset - self._frame_bytes) // self._stride + 1

    @property
    def shape(self):
        """(frames, ny, nx): frames actually present, capped by the header's allocated_frames when that is set"""
        n = self.frames_on_disk()
        a = self._header['allocated_frames']
        return (min(n, a) if a else n, self._header['height'], self._header['width'])

    def get_true_shape(self):
        return self.shape

    def read_into(self, out, z0, z1):
        """frames [z0, z1) -> out[:z1 - z0] (C-contiguous [k, ny, nx] of this file's dtype); returns the count read"""
        n = max(0, min(z1, self.shape[0]) - z0)
        if n == 0:
            return 0
        ny, nx = self._header['height'], self._header['width']
        fd = self._fp.fileno()
        flat = out.reshape(out.shape[0], -1).view(np.uint8)
        if self._stride == self._frame_bytes:
            mv = memoryview(flat[:n]).cast('B')
            off, got = self._image_offset + z0 * self._stride, 0
            while got < n * self._frame_bytes:
                k = os.preadv(fd, [mv[got:]], off + got)
                if k <= 0:
                    raise ValueError('truncated sequence file')
                got += k
        else:
            for i in range(n):
                mv = memoryview(flat[i]).cast('B')
                off, got = self._image_offset + (z0 + i) * self._stride, 0
                while got < self._frame_bytes:
                    k = os.preadv(fd, [mv[got:]], off + got)
                    if k <= 0:
                        raise ValueError('truncated sequence file')
                    got += k
        del ny, nx
        return n

    def __getitem__(self, key):
        ny, nx = self._header['height'], self._header['width']
        if isinstance(key, slice):
            idx = range(*key.indices(self.shape[0]))
            if key.step not in (None, 1):
                return np.stack([self[i][0] for i in idx]) if len(idx) else np.zeros((0, ny, nx), self._dtype)
            out = np.empty((len(idx), ny, nx), dtype=self._dtype)
            if len(idx):
                self.read_into(out, idx.start, idx.stop)
            return out
        z = int(key)
        if z < 0 or z >= self.shape[0]:
            raise IndexError(z)
        out = np.empty((1, ny, nx), dtype=self._dtype)
        self.read_into(out, z, z + 1)
        return out

    def serialize_header(self, fp):
        """the reference stores 1024 zero bytes as the source header of a SEQ-sourced file (em_reader.py:296-300)"""
        fp.write(bytes(SEQ_HEADER_BYTES))


def write_seq(path, frames, allocated_frames=None, version=3, pad_to=None, frame_rate=100.0):
    """Write frames [n, ny, nx] (uint8 / uint16) as a Norpix sequence file: test data and synthetic acquisition chunks.
    pad_to: stride between frames (>= frame bytes + 8); default = frame + timestamp rounded up to 8 bytes."""
    a = np.ascontiguousarray(frames)
    if a.dtype not in (np.uint8, np.uint16):
        raise TypeError('uint8 or uint16 frames')
    n, ny, nx = a.shape
    fb = ny * nx * a.dtype.itemsize
    stride = pad_to or (fb + 8 + 7) // 8 * 8
    if stride < fb + 8:
        raise ValueError('pad_to too small')
    h = bytearray(SEQ_HEADER_BYTES)
    struct.pack_into('<L', h, 0, SEQ_MAGIC)
    h[4:4 + len(_NAME)] = _NAME
    struct.pack_into('<ll', h, 28, version, SEQ_HEADER_BYTES)
    struct.pack_into('<6L', h, 548, nx, ny, 8 * a.dtype.itemsize, 8 * a.dtype.itemsize, fb, 100)
    struct.pack_into('<3L', h, 572, n if allocated_frames is None else allocated_frames, 0, stride)
    struct.pack_into('<d', h, 584, frame_rate)
    off = 8192 if version >= 5 else SEQ_HEADER_BYTES
    tmp = path + '.tmp'
    with open(tmp, 'wb') as fp:
        fp.write(h)
        fp.write(bytes(off - SEQ_HEADER_BYTES))
        for i in range(n):
            fp.write(a[i].tobytes())
            fp.write(struct.pack('<LHH', i // 1000, i % 1000, 0))
            fp.write(bytes(stride - fb - 8))
    os.rename(tmp, path)                      # a chunk appears in the watched directory in one piece


def emfile(file, file_type=None, mode='r', buffering=-1):
    """pyrecode/em_reader.py:11-35 for the one source type implemented here"""
    from .misc import rc_cfg as rc
    if mode != 'r':
        raise NotImplementedError("emfile supports only 'r' mode.")
    if file_type == rc.FILE_TYPE_SEQ:
        return SEQReader(file)
    raise NotImplementedError('only SEQ sources are read on the GPU path (MRC needs mrcfile)')

"""ctypes front-end of the CPU oracle (oracle/recode_oracle.c) plus a byte-level
restatement of the ReCoDe container (header / records / merged file).

TEST INFRASTRUCTURE ONLY.  The product package (pyrecode_b200/) never imports this
module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs do, as the checker or as the timed CPU baseline.

Reference lines are cited per function (paths relative to /root/reference).
Parity status is listed in recode_oracle.c's header; L2 statistics and the L4
centroid map are "parity unpinned" (the reference cannot execute them).
"""
import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u64p = ctypes.POINTER(ctypes.c_uint64)
_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when the reference tree is present)."""
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'recode_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(['make', '-C', _HERE, 'all'], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_binarize_u16.restype = ctypes.c_size_t
        L.orc_l1_values_u16.restype = ctypes.c_size_t
        L.orc_bit_pack_u16.restype = ctypes.c_size_t
        L.orc_label8.restype = ctypes.c_int32
        L.orc_unpack_sparse.restype = ctypes.c_int64
        L.orc_reduce_frame_u16.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


def _as_u16(a):
    return np.ascontiguousarray(a, dtype=np.uint16)


# ---------------------------------------------------------------------------
# per-stage functions
# ---------------------------------------------------------------------------
def make_threshold(dark, eps, dtype=np.uint16):
    """thr = dark + eps in the source dtype, wrapping (recode_writer.py:126-137)."""
    dark = np.ascontiguousarray(dark, dtype=dtype)
    thr = np.empty_like(dark)
    if dtype == np.uint16:
        lib().orc_threshold_u16(_p(dark, _u16p), ctypes.c_uint64(int(eps)), _p(thr, _u16p),
                                ctypes.c_size_t(dark.size))
    elif dtype == np.uint8:
        lib().orc_threshold_u8(_p(dark, _u8p), ctypes.c_uint64(int(eps)), _p(thr, _u8p),
                               ctypes.c_size_t(dark.size))
    else:
        raise NotImplementedError(dtype)
    return thr


def binarize(frame, thr):
    """recode_writer.py:437"""
    f, t = _as_u16(frame), _as_u16(thr)
    b = np.empty(f.shape, dtype=np.uint8)
    lib().orc_binarize_u16(_p(f, _u16p), _p(t, _u16p), _p(b, _u8p), ctypes.c_size_t(f.size))
    return b


def l1_values(frame, thr, binary):
    """recode_writer.py:440"""
    f, t = _as_u16(frame), _as_u16(thr)
    b = np.ascontiguousarray(binary, dtype=np.uint8)
    v = np.empty(f.size, dtype=np.uint16)
    n = lib().orc_l1_values_u16(_p(f, _u16p), _p(t, _u16p), _p(b, _u8p), ctypes.c_size_t(f.size),
                                _p(v, _u16p))
    return v[:n].copy()


def pack_map(binary):
    """recode_writer.py:622-634"""
    b = np.ascontiguousarray(binary, dtype=np.uint8).ravel()
    out = np.empty((b.size + 7) // 8, dtype=np.uint8)
    lib().orc_pack_map(_p(b, _u8p), ctypes.c_size_t(b.size), _p(out, _u8p))
    return out


def bit_pack(vals, b):
    """recode_writer.py:637-652"""
    v = _as_u16(vals).ravel()
    out = np.empty((v.size * b + 7) // 8, dtype=np.uint8)
    if v.size:
        lib().orc_bit_pack_u16(_p(v, _u16p), ctypes.c_size_t(v.size), ctypes.c_int(b), _p(out, _u8p))
    return out


def bit_unpack(packed, n, b):
    """intent of reader.h:74-99"""
    pk = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint64)
    if n:
        lib().orc_bit_unpack(_p(pk, _u8p), ctypes.c_size_t(n), ctypes.c_int(b), _p(out, _u64p))
    return out


def label8(binary):
    """scipy.ndimage.label(binary, 3x3 ones) as called at recode_writer.py:443"""
    b = np.ascontiguousarray(binary, dtype=np.uint8)
    ny, nx = b.shape
    labels = np.empty((ny, nx), dtype=np.int32)
    k = lib().orc_label8(_p(b, _u8p), ctypes.c_int(ny), ctypes.c_int(nx), _p(labels, _i32p))
    if k < 0:
        raise MemoryError
    return labels, int(k)


def l2_stats(labels, frame, k, method):
    """intent of converters.py:262-297; method 0/1 = max, 2 = sum (header code,
    recode_writer.py:358-365)"""
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    f = _as_u16(frame)
    out = np.empty(k, dtype=np.uint16)
    lib().orc_l2_stats_u16(_p(lab, _i32p), _p(f, _u16p), ctypes.c_size_t(f.size), ctypes.c_int32(k),
                           ctypes.c_int(1 if method == 2 else 0), _p(out, _u16p))
    return out


def l4_centroids(labels, frame, k, mode=0):
    """converters.py:157-259; mode = header L4_centroiding code (0/1 weighted, 2 max, 3 unweighted)"""
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    f = _as_u16(frame)
    ny, nx = lab.shape
    out = np.empty((k, 2), dtype=np.float32)
    lib().orc_l4_centroids_u16(_p(lab, _i32p), _p(f, _u16p), ctypes.c_int(ny), ctypes.c_int(nx),
                               ctypes.c_int32(k), ctypes.c_int(mode), _p(out, _f32p))
    return out


def centroid_map(centroids, ny, nx):
    """intent of converters.py:300-309"""
    c = np.ascontiguousarray(centroids, dtype=np.float32)
    b = np.empty((ny, nx), dtype=np.uint8)
    lib().orc_centroid_map(_p(c, _f32p), ctypes.c_int32(c.shape[0]), ctypes.c_int(ny), ctypes.c_int(nx),
                           _p(b, _u8p))
    return b


def recalibrate_l1_frame(frame, original_calibration, new_calibration, epsilon=0.0):
    """converters.py:18-45 for one dense frame: float64 add of the calibration difference, clamp, cast back"""
    diff = original_calibration.astype(np.float64) - (new_calibration.astype(np.float64) + epsilon)
    info = np.iinfo(frame.dtype)
    f = frame.astype(np.float64) + diff
    f[f < info.min] = info.min
    f[f > info.max] = info.max
    return f.astype(frame.dtype)


def l1_to_l4_frame(frame, mode=0, transpose=True):
    """converters.py:84-100 for one dense dark-subtracted frame -> bool [ny, nx] centroid image.
    The reference stores centroid [row_c, col_c] at (col_c, row_c) (:100): transpose=True reproduces that."""
    ny, nx = frame.shape
    labels, k = label8(frame > 0)
    out = np.zeros((ny, nx), dtype=bool)
    if k:
        c = np.round(l4_centroids(labels, frame, k, mode)).astype(np.int64)
        if transpose:
            out[c[:, 1], c[:, 0]] = True
        else:
            out[c[:, 0], c[:, 1]] = True
    return out


def median_std(stack):
    """_median_std_nb (pyrecode/utils/calibration.py:48-57): per-pixel np.median / np.std over the frames, as float32"""
    d = np.asarray(stack)
    return np.median(d, axis=0).astype(np.float32), np.std(d.astype(np.float64), axis=0).astype(np.float32)


def pixel_thresholds(stack, thr, k, as_run=True):
    """_get_pixel_thresh_2 (pyrecode/utils/calibration.py:27-45): of the values above thr the k + 1 largest (float32
    minimum where there are fewer), ascending; (top[0] + top[1]) / 2 with a float32 sum.  as_run: the live reference
    never removes a found maximum (:42 stores a float32 minimum into an unsigned list: no effect), so all k + 1 kept
    values are the maximum."""
    d = np.asarray(stack)
    n, ny, nx = d.shape
    fmin = np.finfo(np.float32).min
    out = np.empty((ny, nx), dtype=np.float32)
    with np.errstate(over='ignore'):
        for r in range(ny):
            for c in range(nx):
                v = np.sort(d[:, r, c][d[:, r, c] > thr[r, c]].astype(np.float32))[::-1][:k + 1]
                top = np.full(k + 1, fmin, dtype=np.float32)
                top[:v.size] = v
                if as_run and v.size:
                    top[:] = v[0]
                top.sort()
                out[r, c] = np.float32(np.float64(np.float32(top[0] + top[1])) / 2)
    return out


def unpack_sparse(ny, nx, b, map_bytes, val_bytes, level):
    """reader.h:10-68; returns uint64 [n, 3] (row, col, value)"""
    m = np.frombuffer(bytes(map_bytes), dtype=np.uint8)
    v = np.frombuffer(bytes(val_bytes) + b'\0' * 8, dtype=np.uint8)
    out = np.empty((ny * nx, 3), dtype=np.uint64)
    n = lib().orc_unpack_sparse(ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(b), _p(m, _u8p), _p(v, _u8p),
                                _p(out, _u64p), ctypes.c_int(level))
    return out[:n].copy()


def unpack_dense(ny, nx, b, map_bytes, val_bytes, level):
    m = np.frombuffer(bytes(map_bytes), dtype=np.uint8)
    v = np.frombuffer(bytes(val_bytes) + b'\0' * 8, dtype=np.uint8)
    out = np.empty((ny, nx), dtype=np.uint16)
    lib().orc_unpack_dense_u16(ctypes.c_int(nx), ctypes.c_int(ny), ctypes.c_int(b), _p(m, _u8p), _p(v, _u8p),
                               _p(out, _u16p), ctypes.c_int(level))
    return out


def reduce_frame(frame, thr, level, bit_depth, l2_statistics=0, l4_centroiding=0, _scratch={}):
    """Whole per-frame reduction (recode_writer.py:436-477) -> (map_bytes, packed_bytes, count).
    count = foreground pixels (L1/L3) or puddles (L2/L4)."""
    f, t = _as_u16(frame), _as_u16(thr)
    ny, nx = f.shape
    n = ny * nx
    key = n
    if key not in _scratch:
        _scratch.clear()
        _scratch[key] = (np.empty(n * 16 + 64, dtype=np.uint8), np.empty((n + 7) // 8, dtype=np.uint8),
                         np.empty(n * 2 + 8, dtype=np.uint8))
    scratch, map_out, packed = _scratch[key]
    sizes = (ctypes.c_uint64 * 3)()
    rc = lib().orc_reduce_frame_u16(_p(f, _u16p), _p(t, _u16p), ctypes.c_int(ny), ctypes.c_int(nx),
                                    ctypes.c_int(level), ctypes.c_int(bit_depth),
                                    ctypes.c_int(1 if l2_statistics == 2 else 0),
                                    ctypes.c_int(l4_centroiding),
                                    _p(map_out, _u8p), _p(packed, _u8p), _p(scratch, _u8p), sizes)
    if rc != 0:
        raise RuntimeError('oracle reduce failed: %d' % rc)
    return map_out[:sizes[1]].tobytes(), packed[:sizes[2]].tobytes(), int(sizes[0])


# ---------------------------------------------------------------------------
# container restatement
# ---------------------------------------------------------------------------
# v0.2 header: (name, bytes) in file order -- recode_header.py:57-94
HEADER_FIELDS = [
    ('uid', 8), ('version_major', 1), ('version_minor', 1), ('is_intermediate', 1), ('reduction_level', 1),
    ('rc_operation_mode', 1), ('is_bit_packed', 1), ('target_bit_depth', 1), ('nx', 4), ('ny', 4), ('nz', 4),
    ('frame_metadata_size', 1), ('num_non_standard_frame_metadata', 1), ('L2_statistics', 1),
    ('L4_centroiding', 1), ('compression_scheme', 1), ('compression_level', 1), ('source_file_type', 1),
    ('source_header_length', 2), ('source_header_position', 1), ('source_file_name', 100),
    ('calibration_file_name', 100), ('calibration_threshold_epsilon', 8), ('has_calibration_data', 1),
    ('frame_offset', 4), ('calibration_frame_offset', 4), ('num_calibration_frames', 4),
    ('source_bit_depth', 1), ('source_dtype', 1), ('target_dtype', 1), ('checksum', 32), ('futures', 219)]
HEADER_LEN = sum(b for _, b in HEADER_FIELDS)
assert HEADER_LEN == 512
UID = 158966344846346


def parse_header(buf):
    """-> dict of the v0.2 header (recode_header.py:188-249)."""
    d, off = {}, 0
    for name, nb in HEADER_FIELDS:
        raw = buf[off:off + nb]
        if name in ('source_file_name', 'calibration_file_name'):
            d[name] = raw.decode('utf-8', 'replace')
        elif name in ('checksum', 'futures'):
            d[name] = bytes(raw)
        else:
            d[name] = int.from_bytes(raw, 'little')
        off += nb
    return d


def build_header(fields):
    """serialize a v0.2 header (recode_header.py:257-275): ints little-endian, names space padded."""
    out = bytearray()
    for name, nb in HEADER_FIELDS:
        v = fields.get(name, 0)
        if name in ('source_file_name', 'calibration_file_name'):
            s = str(v)[:nb].ljust(nb, ' ')
            out += s.encode('utf-8')
        elif name in ('checksum', 'futures'):
            out += bytes(nb)
        else:
            out += int(v).to_bytes(nb, 'little')
    return bytes(out)


def metadata_fields(level, mode):
    """structures.py:18-46 -- names of the per-frame uint32 metadata fields (without frame_id)."""
    if level in (1, 2):
        s = 'pixvals' if level == 1 else 'summary_stats'
        if mode == 0:
            return ['bytes_in_packed_' + s]
        return ['bytes_in_compressed_binary_map', 'bytes_in_compressed_' + s, 'bytes_in_packed_' + s]
    return [] if mode == 0 else ['bytes_in_compressed_binary_map']


def build_record(frame_id, level, mode, map_bytes, packed_bytes, compression_level=1):
    """recode_writer.py:482-550 with stock zlib as the codec (recode_compressors.py:84-85)."""
    rec = struct.pack('<I', frame_id)
    if mode == 0:
        if level in (1, 2):
            return rec + struct.pack('<I', len(packed_bytes)) + map_bytes + packed_bytes
        return rec + map_bytes
    cm = zlib.compress(map_bytes, compression_level)
    if level in (1, 2):
        cv = zlib.compress(packed_bytes, compression_level)
        return rec + struct.pack('<III', len(cm), len(cv), len(packed_bytes)) + cm + cv
    return rec + struct.pack('<I', len(cm)) + cm


def parse_part_file(path_or_bytes):
    """Walk an intermediate (part) file -> (header dict, [record dict]).
    Each record: frame_id, metadata dict, map (raw bytes, inflated if mode 1), vals (ditto or None)."""
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, 'rb').read()
    h = parse_header(buf)
    level, mode = h['reduction_level'], h['rc_operation_mode']
    names = metadata_fields(level, mode)
    map_len = (h['nx'] * h['ny'] + 7) // 8
    off = HEADER_LEN + h['source_header_length']
    recs = []
    while off < len(buf):
        fid = struct.unpack_from('<I', buf, off)[0]
        off += 4
        md = {}
        for nme in names:
            md[nme] = struct.unpack_from('<I', buf, off)[0]
            off += 4
        if mode == 1:
            n1 = md['bytes_in_compressed_binary_map']
            cmap = bytes(buf[off:off + n1]); off += n1
            m = zlib.decompress(cmap)
            v, cv = None, None
            if level in (1, 2):
                n2 = md[names[1]]
                cv = bytes(buf[off:off + n2]); off += n2
                v = zlib.decompress(cv)
            recs.append(dict(frame_id=fid, metadata=md, map=m, vals=v, cmap=cmap, cvals=cv))
        else:
            m = bytes(buf[off:off + map_len]); off += map_len
            v = None
            if level in (1, 2):
                n2 = md[names[0]]
                v = bytes(buf[off:off + n2]); off += n2
            recs.append(dict(frame_id=fid, metadata=md, map=m, vals=v, cmap=None, cvals=None))
    return h, recs


def parse_merged_file(path_or_bytes):
    """Walk a merged file: header, nz x metadata table, payloads in frame order."""
    buf = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, 'rb').read()
    h = parse_header(buf)
    level, mode = h['reduction_level'], h['rc_operation_mode']
    names = metadata_fields(level, mode)
    map_len = (h['nx'] * h['ny'] + 7) // 8
    off = HEADER_LEN + h['source_header_length']
    table = np.frombuffer(buf, dtype='<u4', count=h['nz'] * len(names), offset=off).reshape(h['nz'], len(names))
    off += table.nbytes
    recs = []
    for z in range(h['nz']):
        md = {nme: int(table[z, i]) for i, nme in enumerate(names)}
        if mode == 1:
            n1 = md['bytes_in_compressed_binary_map']
            m = zlib.decompress(bytes(buf[off:off + n1])); off += n1
            v = None
            if level in (1, 2):
                n2 = md[names[1]]
                v = zlib.decompress(bytes(buf[off:off + n2])); off += n2
        else:
            m = bytes(buf[off:off + map_len]); off += map_len
            v = None
            if level in (1, 2):
                n2 = md[names[0]]
                v = bytes(buf[off:off + n2]); off += n2
        recs.append(dict(frame_id=z, metadata=md, map=m, vals=v))
    assert off == len(buf), (off, len(buf))
    return h, recs


def partition(n_frames, num_threads, node_id):
    """recode_writer.py:320-322 -> (frame_offset, available_frames)"""
    per = -(-n_frames // num_threads)
    off = node_id * per
    return off, min(per, max(n_frames - off, 0))


# ---------------------------------------------------------------------------
# synthetic frames of SURVEY 8(d) (shared by tests and bench so both sides see
# the same inputs)
# ---------------------------------------------------------------------------
# the synthetic workload generator lives with the product's bench support, not with the checker
from pyrecode_b200.synth import synth_dark, synth_frames  # noqa: E402,F401

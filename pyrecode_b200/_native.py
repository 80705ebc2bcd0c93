"""ctypes binding of librecode_b200.so (include/recode_b200.h) plus thin torch-tensor helpers.

The library is the product's only compute path.  If it is missing or no CUDA device is present every entry
point raises -- there is no CPU fallback (BASELINE.json north_star).
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'librecode_b200.so')
_lib = None

RC_STATUS_RECORDS_OVERFLOW = 1
RC_STATUS_BAD_STREAM = 2
RC_STATUS_OUT_OVERFLOW = 4
RC_STATUS_SIZE_MISMATCH = 8


class RcConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'ny', 'nx', 'itemsize', 'bit_depth', 'reduction_level', 'rc_operation_mode', 'l2_statistics',
        'l4_centroiding', 'compression_level', 'max_frames')]


_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_cfgp = ctypes.POINTER(RcConfig)

_SIGNATURES = {
    'rc_create': (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int]),
    'rc_destroy': (None, [_vp]),
    'rc_last_error': (ctypes.c_char_p, [_vp]),
    'rc_version': (ctypes.c_int, []),
    'rc_sm_count': (ctypes.c_int, [_vp]),
    'rc_profile_enable': (ctypes.c_int, [_vp, ctypes.c_int]),
    'rc_profile_read': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
    'rc_profile_read_detail': (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
    'rc_launch_count': (ctypes.c_ulonglong, [_vp]),
    'rc_set_pipelined': (ctypes.c_int, [_vp, ctypes.c_int]),
    'rc_map_stride_words': (_sz, [_sz]),
    'rc_packed_stride_bytes': (_sz, [_cfgp]),
    'rc_workspace_bytes': (_sz, [_cfgp]),
    'rc_stage_workspace_bytes': (_sz, [_cfgp]),
    'rc_records_capacity': (_sz, [_cfgp]),
    'rc_read_workspace_bytes': (_sz, [_cfgp]),
    'rc_reduce_compress': (ctypes.c_int, [_vp, _cfgp, _vp, ctypes.c_int, _vp, ctypes.c_uint32, _vp, _sz, _vp, _sz,
                                          _vp, _vp, _vp, _vp]),
    'rc_make_threshold': (ctypes.c_int, [_vp, _cfgp, _vp, ctypes.c_uint64, _vp, _vp]),
    'rc_reduce': (ctypes.c_int, [_vp, _cfgp, _vp, ctypes.c_int, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp]),
    'rc_ccl_label': (ctypes.c_int, [_vp, _cfgp, _vp, ctypes.c_int, _vp, _sz, _vp, _vp, _vp]),
    'rc_l4_centroids': (ctypes.c_int, [_vp, _cfgp, _vp, ctypes.c_int, _vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    'rc_deflate_bound': (_sz, [_sz]),
    'rc_deflate_workspace_bytes': (_sz, [ctypes.c_int, _sz]),
    'rc_deflate_zlib': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, _sz, _vp, _sz, _vp, _sz, _vp,
                                       _vp]),
    'rc_inflate_workspace_bytes': (_sz, [ctypes.c_int, _sz]),
    'rc_inflate_zlib': (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _sz, _vp, _sz, _vp, _vp, _vp]),
    'rc_unpack_sparse': (ctypes.c_int, [_vp, _cfgp, _vp, _vp, _sz, ctypes.c_int, _vp, _sz, _vp, _sz, _vp, _vp]),
    'rc_unpack_dense': (ctypes.c_int, [_vp, _cfgp, _vp, _vp, _sz, ctypes.c_int, _vp, _sz, _vp, _vp, _vp, _vp]),
    'rc_bit_unpack': (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint64, _vp, _vp]),
    'rc_recalibrate': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _sz, ctypes.c_int, _vp, _vp]),
    'rc_pixel_thresholds': (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int, _sz, _vp, ctypes.c_int, ctypes.c_int, _vp,
                                           _vp]),
    'rc_median_std_workspace_bytes': (_sz, [_sz]),
    'rc_median_std': (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int, _sz, _vp, _vp, _vp, _sz, _vp]),
    'rc_bit_pack': (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_uint64, _vp, _vp]),
}

EXPORTS = tuple(_SIGNATURES)


def library_path():
    return _SO


def lib():
    """Load the shared library (no compute, works without a GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError('librecode_b200.so is not built: run `python -m pyrecode_b200.build` '
                               '(there is no CPU fallback)')
        L = ctypes.CDLL(_SO)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def make_config(ny, nx, itemsize, bit_depth, reduction_level, rc_operation_mode=1, l2_statistics=0,
                l4_centroiding=0, compression_level=1, max_frames=1):
    return RcConfig(int(ny), int(nx), int(itemsize), int(bit_depth), int(reduction_level), int(rc_operation_mode),
                    int(l2_statistics), int(l4_centroiding), int(compression_level), int(max_frames))


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One rc_ctx bound to a CUDA device.  All tensors passed in must live on that device."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('pyrecode_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self._lib = lib()
        h = ctypes.c_void_p()
        rc = self._lib.rc_create(ctypes.byref(h), self.device.index)
        if rc != 0:
            raise RuntimeError('rc_create failed with code %d (needs an sm_100 class GPU)' % rc)
        self._h = h

    def close(self):
        if self._h:
            self._lib.rc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError('%s failed (%d): %s' % (what, rc, self._lib.rc_last_error(self._h).decode()))

    def empty(self, n, dtype=torch.uint8):
        return torch.empty(int(n), dtype=dtype, device=self.device)

    def zeros(self, n, dtype=torch.uint8):
        return torch.zeros(int(n), dtype=dtype, device=self.device)

    # ---- instrumentation -----------------------------------------------------------------------
    STAGES = ('threshold_pack_compact', 'reduce_rest', 'deflate', 'assemble')

    def set_pipelined(self, on=True):
        self._check(self._lib.rc_set_pipelined(self._h, 1 if on else 0), 'rc_set_pipelined')

    def profile_enable(self, on=True):
        self._check(self._lib.rc_profile_enable(self._h, int(on)), 'rc_profile_enable')

    def profile_read(self):
        buf = (ctypes.c_float * 8)()
        n = self._lib.rc_profile_read(self._h, buf, 8)
        return [float(buf[i]) for i in range(n)]

    def profile_read_detail(self):
        """per-kernel milliseconds of the reduction's second stage (after profile_enable(2))"""
        buf = (ctypes.c_float * 16)()
        n = self._lib.rc_profile_read_detail(self._h, buf, 16)
        return [float(buf[i]) for i in range(n)]

    def launch_count(self):
        return int(self._lib.rc_launch_count(self._h))

    # ---- sizes -------------------------------------------------------------------------------
    def map_stride_words(self, n_pixels):
        return self._lib.rc_map_stride_words(n_pixels)

    def packed_stride_bytes(self, cfg):
        return self._lib.rc_packed_stride_bytes(ctypes.byref(cfg))

    def workspace_bytes(self, cfg):
        return self._lib.rc_workspace_bytes(ctypes.byref(cfg))

    def records_capacity(self, cfg):
        return self._lib.rc_records_capacity(ctypes.byref(cfg))

    def stage_workspace_bytes(self, cfg):
        return self._lib.rc_stage_workspace_bytes(ctypes.byref(cfg))

    def read_workspace_bytes(self, cfg):
        return self._lib.rc_read_workspace_bytes(ctypes.byref(cfg))

    # ---- write side --------------------------------------------------------------------------
    def make_threshold(self, cfg, dark, eps):
        thr = torch.empty_like(dark)
        self._check(self._lib.rc_make_threshold(self._h, ctypes.byref(cfg), _ptr(dark), int(eps) & (2 ** 64 - 1),
                                                _ptr(thr), _stream()), 'rc_make_threshold')
        return thr

    def reduce_compress(self, cfg, frames, n_frames, thr, first_frame_id, ws, records, offsets, counts, status):
        self._check(self._lib.rc_reduce_compress(self._h, ctypes.byref(cfg), _ptr(frames), n_frames, _ptr(thr),
                                                 first_frame_id, _ptr(ws), ws.numel(), _ptr(records), records.numel(),
                                                 _ptr(offsets), _ptr(counts), _ptr(status), _stream()),
                    'rc_reduce_compress')

    def reduce(self, cfg, frames, n_frames, thr, ws, maps, packed, packed_bytes, counts):
        self._check(self._lib.rc_reduce(self._h, ctypes.byref(cfg), _ptr(frames), n_frames, _ptr(thr), _ptr(ws),
                                        ws.numel(), _ptr(maps), _ptr(packed), _ptr(packed_bytes), _ptr(counts),
                                        _stream()), 'rc_reduce')

    def ccl_label(self, cfg, maps, n_frames, ws, labels, counts):
        self._check(self._lib.rc_ccl_label(self._h, ctypes.byref(cfg), _ptr(maps), n_frames, _ptr(ws), ws.numel(),
                                           _ptr(labels), _ptr(counts), _stream()), 'rc_ccl_label')

    def l4_centroids(self, cfg, frames, n_frames, thr, ws, centroids, capacity, counts):
        self._check(self._lib.rc_l4_centroids(self._h, ctypes.byref(cfg), _ptr(frames), n_frames, _ptr(thr), _ptr(ws),
                                              ws.numel(), _ptr(centroids), capacity, _ptr(counts), _stream()),
                    'rc_l4_centroids')

    def deflate_bound(self, n):
        return self._lib.rc_deflate_bound(n)

    def deflate_zlib(self, level, data, in_off, in_bytes, n_streams, max_in_bytes, out, out_stride, out_bytes):
        need = self._lib.rc_deflate_workspace_bytes(n_streams, max_in_bytes)
        ws = self.empty(need)
        self._check(self._lib.rc_deflate_zlib(self._h, level, _ptr(data), _ptr(in_off), _ptr(in_bytes), n_streams,
                                              max_in_bytes, _ptr(ws), need, _ptr(out), out_stride, _ptr(out_bytes),
                                              _stream()), 'rc_deflate_zlib')
        return ws      # keep alive until the stream has run

    # ---- read side ---------------------------------------------------------------------------
    def inflate_zlib(self, data, in_off, in_bytes, n_streams, out, out_stride, out_bytes, status, ws=None):
        need = self._lib.rc_inflate_workspace_bytes(n_streams, out_stride)
        if ws is None or ws.numel() < need:
            ws = self.empty(need)
        self._check(self._lib.rc_inflate_zlib(self._h, _ptr(data), _ptr(in_off), _ptr(in_bytes), n_streams, _ptr(ws),
                                              ws.numel(), _ptr(out), out_stride, _ptr(out_bytes), _ptr(status),
                                              _stream()), 'rc_inflate_zlib')
        return ws

    def unpack_sparse(self, cfg, maps, packed, packed_stride, n_frames, ws, triples, capacity, counts):
        self._check(self._lib.rc_unpack_sparse(self._h, ctypes.byref(cfg), _ptr(maps), _ptr(packed), packed_stride,
                                               n_frames, _ptr(ws), ws.numel(), _ptr(triples), capacity, _ptr(counts),
                                               _stream()), 'rc_unpack_sparse')

    def unpack_dense(self, cfg, maps, packed, packed_stride, n_frames, ws, dense, total, counts):
        self._check(self._lib.rc_unpack_dense(self._h, ctypes.byref(cfg), _ptr(maps), _ptr(packed), packed_stride,
                                              n_frames, _ptr(ws), ws.numel(), _ptr(dense), _ptr(total), _ptr(counts),
                                              _stream()), 'rc_unpack_dense')

    def bit_unpack(self, bit_depth, packed, n_values, out):
        self._check(self._lib.rc_bit_unpack(self._h, bit_depth, _ptr(packed), n_values, _ptr(out), _stream()),
                    'rc_bit_unpack')

    def recalibrate(self, itemsize, frames, diff, n_pixels, n_frames, out):
        self._check(self._lib.rc_recalibrate(self._h, itemsize, _ptr(frames), _ptr(diff), n_pixels, n_frames, _ptr(out),
                                             _stream()), 'rc_recalibrate')

    def median_std(self, itemsize, stack, n_frames, n_pixels, median, std):
        ws = self.empty(self._lib.rc_median_std_workspace_bytes(n_pixels))
        self._check(self._lib.rc_median_std(self._h, itemsize, _ptr(stack), n_frames, n_pixels, _ptr(median), _ptr(std),
                                            _ptr(ws), ws.numel(), _stream()), 'rc_median_std')
        return ws

    def pixel_thresholds(self, itemsize, stack, n_frames, n_pixels, thr, expected_n_events, as_run, out):
        self._check(self._lib.rc_pixel_thresholds(self._h, itemsize, _ptr(stack), n_frames, n_pixels, _ptr(thr),
                                                  expected_n_events, 1 if as_run else 0, _ptr(out), _stream()),
                    'rc_pixel_thresholds')

    def bit_pack(self, bit_depth, vals, n_values, packed):
        self._check(self._lib.rc_bit_pack(self._h, bit_depth, _ptr(vals), n_values, _ptr(packed), _stream()),
                    'rc_bit_pack')


def numpy_dtype(itemsize):
    return np.uint8 if itemsize == 1 else np.uint16


def torch_dtype(itemsize):
    return torch.uint8 if itemsize == 1 else torch.uint16

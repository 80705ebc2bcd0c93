"""recalibrate_l1 / l1_to_l4_converter -- drop-in for pyrecode/utils/converters.py:15-123 on the GPU.

Both take and return the reader's frame dictionaries {frame_id: {'metadata': ..., 'data': coo_matrix}}.  The
per-frame arithmetic of the reference runs in librecode_b200:

  recalibrate_l1      frame.astype(float64) + (original - (new + epsilon)), clamped to the dtype's range and cast
                      back (converters.py:18-45)                      -> rc_recalibrate (k_recalibrate)
  l1_to_l4_converter  scipy.ndimage.label(frame > 0, 3x3) + get_centroids_2D_nb weighted by the dark-subtracted
                      values (converters.py:86-90)                    -> rc_l4_centroids (k_reduce_tiles +
                      k_ccl_tiles<3> + k_l4_open) with a zero threshold frame

The COO containers around them (todense / coo_matrix) are host-side plumbing as in the reference.  There is no CPU
fallback: without the CUDA library these functions raise.

Reference behaviour kept on purpose:
  * l1_to_l4_converter stores a centroid [row_c, col_c] at matrix position (col_c, row_c) -- the transpose
    (converters.py:100); frames that are not square therefore fail in the reference.  `transpose=True` (default)
    reproduces that bit for bit, `transpose=False` gives the intended (row_c, col_c) map and works for any shape.
  * only method 'weighted_average' is reachable in the reference (converters.py:159-164); 'max' and 'unweighted'
    follow _get_centroids_2d_nb_m / _u as intended ("no reference behaviour").
  * recalibrate_l1 adds the calibration difference to EVERY pixel, background included, as the reference does.
"""
import copy
from datetime import datetime

import numpy as np
from scipy.sparse import coo_matrix

_METHODS = {'weighted_average': 0, 'max': 2, 'unweighted': 3}


def _deep_copy_frame_metadata(src, dst, key):
    # converters.py:126-131: everything but the data is deep-copied
    dst[key] = {}
    for k in src[key]:
        if k != 'data':
            dst[key][k] = copy.deepcopy(src[key][k])


def _dense_batch(frames, keys, shape, dtype):
    out = np.zeros((len(keys),) + tuple(shape), dtype=dtype)
    for i, k in enumerate(keys):
        c = frames[k]['data'].tocoo()
        out[i, c.row, c.col] = c.data           # reader-made COO matrices hold every pixel once
    return out


def recalibrate_l1(l1_frames, n_frames=-1, original_calibration_frame=None, new_calibration_frame=None,
                   epsilon=0.0, in_place=False, device=None, batch_frames=16):
    import torch
    from .._native import Context
    if n_frames < 1:
        n_frames = len(l1_frames)
    diff = original_calibration_frame.astype(np.float64) - (new_calibration_frame.astype(np.float64) + epsilon)
    keys = list(l1_frames.keys())
    dtype = np.dtype(l1_frames[keys[0]]['data'].dtype)
    if dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise NotImplementedError('the GPU path handles uint8 / uint16 frames (got %s)' % dtype)
    # converters.py:52-53 breaks when n_frames == frame_count AFTER processing that frame: n_frames + 1 frames
    keys = keys[:n_frames + 1]
    shape = l1_frames[keys[0]]['data'].shape
    P = shape[0] * shape[1]
    if (P * dtype.itemsize) % 16:
        raise NotImplementedError('frame size must be a multiple of 16 bytes')
    ctx = Context(device)
    out = {}
    start = datetime.now()
    tdt = torch.uint8 if dtype.itemsize == 1 else torch.uint16
    with torch.cuda.device(ctx.device):
        d_diff = torch.from_numpy(np.ascontiguousarray(diff.reshape(-1))).to(ctx.device)
        for i0 in range(0, len(keys), batch_frames):
            part = keys[i0:i0 + batch_frames]
            d_in = torch.from_numpy(_dense_batch(l1_frames, part, shape, dtype)).to(ctx.device)
            d_out = torch.empty_like(d_in)
            ctx.recalibrate(dtype.itemsize, d_in, d_diff, P, len(part), d_out)
            res = d_out.view(tdt).cpu().numpy().view(dtype)
            for j, key in enumerate(part):
                if in_place:
                    out[key] = l1_frames[key]
                else:
                    _deep_copy_frame_metadata(l1_frames, out, key)
                out[key]['data'] = coo_matrix(res[j], dtype=dtype)
    print('Total processing time: ' + str(datetime.now() - start))
    return out


def recalibrate_dense(frames, original_calibration_frame, new_calibration_frame, epsilon=0.0, ctx=None):
    """Device-resident form: frames = CUDA tensor [n, ny, nx] (uint8 / uint16, e.g. from
    ReCoDeReader.read_frames_dense) -> recalibrated CUDA tensor of the same shape and dtype."""
    import torch
    from .._native import Context
    ctx = ctx or Context(frames.device.index)
    itemsize = frames.element_size()
    n, ny, nx = frames.shape
    diff = original_calibration_frame.astype(np.float64) - (new_calibration_frame.astype(np.float64) + epsilon)
    with torch.cuda.device(ctx.device):
        d_diff = torch.from_numpy(np.ascontiguousarray(diff.reshape(-1))).to(ctx.device)
        out = torch.empty_like(frames)
        ctx.recalibrate(itemsize, frames.contiguous(), d_diff, ny * nx, n, out)
    return out


def l1_to_l4_converter(l1_frames, frame_shape, n_frames=-1, area_threshold=0, verbosity=0, method='weighted_average',
                       in_place=False, transpose=True, device=None, batch_frames=16):
    from ..engine import WriteEngine
    if area_threshold != 0:
        raise NotImplementedError('area_threshold other than 0 is not supported on the GPU path')
    if method not in _METHODS:
        raise ValueError('unknown centroiding method %r' % (method,))
    ny, nx = int(frame_shape[0]), int(frame_shape[1])
    max_dim = max(ny, nx)
    cdt = None
    for d in (np.uint8, np.uint16, np.uint32, np.uint64):          # converters.py:64-70
        if max_dim < np.iinfo(d).max:
            cdt = d
            break
    if cdt is None:
        raise ValueError("Unable to identify data type for centroids")
    keys = list(l1_frames.keys())
    if n_frames > 0:
        keys = keys[:n_frames + 1]                                  # converters.py:119-120 (break after the frame)
    if not keys:
        return {}
    dtype = np.dtype(l1_frames[keys[0]]['data'].dtype)
    if dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise NotImplementedError('the GPU path handles uint8 / uint16 frames (got %s)' % dtype)
    eng = WriteEngine(ny, nx, dtype.itemsize, 8 * dtype.itemsize, 4, 1, 0, _METHODS[method], 1,
                      max_frames=batch_frames, device=device)
    eng.set_threshold(np.zeros((ny, nx), dtype=dtype), 0)          # foreground = frame > 0
    t = np.ones(ny * nx, dtype=bool)
    cd = {}
    start = datetime.now()
    dose = 0.0
    for i0 in range(0, len(keys), batch_frames):
        part = keys[i0:i0 + batch_frames]
        cents = eng.centroids(_dense_batch(l1_frames, part, (ny, nx), dtype))
        for key, c in zip(part, cents):
            if in_place:
                cd[key] = l1_frames[key]
            else:
                _deep_copy_frame_metadata(l1_frames, cd, key)
            k = len(c)
            dose += k / float(ny * nx)
            if k:
                ci = np.round(c).astype(cdt)                       # [row_c, col_c], half to even
                rows, cols = (ci[:, 1], ci[:, 0]) if transpose else (ci[:, 0], ci[:, 1])
                cd[key]['data'] = coo_matrix((t[:k], (rows, cols)), shape=(ny, nx), dtype=bool)
            else:
                cd[key]['data'] = coo_matrix((ny, nx), dtype=bool)
            if verbosity > 0:
                print(key)
                print('Dose Rate =', k / float(ny * nx))
    if verbosity == 0 and len(keys) > 100:
        print('Avg. Dose Rate = {0:0.4f}'.format(dose / len(keys)))
    print('Total processing time: ' + str(datetime.now() - start))
    return cd

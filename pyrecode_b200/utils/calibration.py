"""make_calibration_frames -- drop-in for pyrecode/utils/calibration.py:87-138 with the per-pixel statistics on the GPU.

A stack of flat-field / dark frames -> per-pixel median and standard deviation (rc_median_std, replacing the numba
loops of _median_std_nb, calibration.py:48-57) -> a Gaussian fit of the zero-centred intensities of the last
`n_stats_frames` frames (numpy / scipy on the host, exactly the reference's _get_fit_params, :64-84) -> the threshold
frames floor(median + i * sigma), i = 0 .. n_sigmas - 1, written as <prefix>_dark_ref_<i>.bin like the reference.
The per-sigma event statistics the reference prints (:114-123) come from the GPU reduction kernels (puddle counts and
foreground fractions of the stats frames against each threshold frame).

The reference reads the stack with pims (a .seq file); pims is not a dependency here: pass the frames as `data`
([nFrames, ny, nx] numpy array) or point `filepath` at a raw binary stack of `dtype` together with `shape=(ny, nx)`.
`use_acc=True` adds the per-pixel thresholds of _get_pixel_thresh_2 (:27-45) for sigma index `sigma_acc`
(rc_pixel_thresholds).  No CPU fallback.
"""
import os
from datetime import datetime

import numpy as np


def _gaussian(x, a, x0, sigma):
    return a * np.exp(-(x - x0) ** 2 / (2 * sigma ** 2))


def _get_fit_params(d, nFrames, n_stats_frames, _m):
    """calibration.py:64-84 (host): sigma of the Gaussian fitted to the histogram of the zero-centred stats frames"""
    from scipy.optimize import curve_fit
    dsd = np.zeros((n_stats_frames,) + d.shape[1:])
    for i, f in enumerate(range(nFrames - n_stats_frames, nFrames)):
        dsd[i] = d[f] - _m
    h, edges = np.histogram(dsd.flatten(), bins=100, density=False)
    c = [(edges[i] + edges[i + 1]) / 2 for i in range(len(edges) - 1)]
    hn = h / np.sum(h)
    mean = np.average(c, weights=hn)
    sigma = np.sqrt(np.average((c - mean) ** 2, weights=hn))
    _p0 = [np.max(hn), mean, sigma]
    popt, pcov = curve_fit(_gaussian, c, hn, p0=_p0)
    print("\n Fit Result \n Init params=", _p0, "\n Optimal params=", popt)
    return popt[2]


def median_std(data, device=None, frames_per_upload=None):
    """[nFrames, ny, nx] uint8 / uint16 (numpy array or CUDA tensor) -> (median, std) float32 [ny, nx] numpy arrays"""
    import torch
    from .._native import Context
    ctx = Context(device)
    if isinstance(data, torch.Tensor):
        stack = data.contiguous()
        itemsize = stack.element_size()
    else:
        a = np.ascontiguousarray(data)
        if a.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
            raise NotImplementedError('the GPU path handles uint8 / uint16 stacks (got %s)' % a.dtype)
        itemsize = a.dtype.itemsize
        stack = torch.from_numpy(a).to(ctx.device)
    n, ny, nx = stack.shape
    with torch.cuda.device(ctx.device):
        med = torch.empty(ny * nx, dtype=torch.float32, device=ctx.device)
        sd = torch.empty(ny * nx, dtype=torch.float32, device=ctx.device)
        ws = ctx.median_std(itemsize, stack, n, ny * nx, med, sd)
        torch.cuda.synchronize()
        del ws
        return med.cpu().numpy().reshape(ny, nx), sd.cpu().numpy().reshape(ny, nx)


def pixel_thresholds(data, thr, expected_n_events, device=None, as_run=True):
    """_get_pixel_thresh_2 (calibration.py:27-45) -> float32 [ny, nx].  As written: per pixel, the mean of the (k + 1)-th
    and k-th largest of the stack values above thr[pixel] (k = expected_n_events; missing values count as the float32
    minimum) -- as_run=False.  As it runs (numba 0.65; as_run=True, the default, bit-exact with the live reference): the
    "remove the maximum" store of a float32 minimum into a list of unsigned integers changes nothing, so the result is
    the largest value above thr[pixel] whatever k."""
    import torch
    from .._native import Context
    ctx = Context(device)
    a = np.ascontiguousarray(data)
    if a.dtype not in (np.dtype(np.uint8), np.dtype(np.uint16)):
        raise NotImplementedError('the GPU path handles uint8 / uint16 stacks (got %s)' % a.dtype)
    n, ny, nx = a.shape
    with torch.cuda.device(ctx.device):
        stack = torch.from_numpy(a).to(ctx.device)
        t = torch.from_numpy(np.ascontiguousarray(thr, dtype=np.float32).reshape(-1)).to(ctx.device)
        out = torch.empty(ny * nx, dtype=torch.float32, device=ctx.device)
        ctx.pixel_thresholds(a.dtype.itemsize, stack, n, ny * nx, t, int(expected_n_events), as_run, out)
        torch.cuda.synchronize()
        return out.cpu().numpy().reshape(ny, nx)


def make_calibration_frames(filepath, dtype, nFrames, n_stats_frames, n_sigmas, savepath='', filename_prefix='',
                            use_acc=False, sigma_acc=-1, data=None, shape=None, device=None, acc_as_run=True):
    from ..engine import WriteEngine
    if not filename_prefix.endswith('_'):
        filename_prefix += '_'
    start = datetime.now()
    if data is None:
        if shape is None:
            raise NotImplementedError('reading .seq stacks needs pims; pass data= or a raw binary stack with shape=(ny, nx)')
        d = np.fromfile(filepath, dtype=dtype, count=nFrames * shape[0] * shape[1]).reshape(nFrames, shape[0], shape[1])
    else:
        d = np.asarray(data)[:nFrames].astype(dtype, copy=False)
    ny, nx = d.shape[1], d.shape[2]

    _m, _stds = median_std(d, device=device)
    _fit_std = _get_fit_params(d, nFrames, n_stats_frames, _m)
    print('\nAvg. std.dev. per pixel:', np.average(_stds))
    print('Global intensity std. dev.:', _fit_std)
    print("Calibration time:", datetime.now() - start, "\n")

    n_pixels_in_frame = nx * ny
    itemsize = np.dtype(dtype).itemsize
    stats = d[nFrames - n_stats_frames:nFrames]
    eng = None
    out = []
    acc = None
    for i in range(n_sigmas):
        t = np.floor(_m + _fit_std * i).astype(dtype)
        t.tofile(os.path.join(savepath, filename_prefix + "_dark_ref_" + str(i) + ".bin"))
        out.append(t)
        # events (8-connected puddles) and foreground pixels of the stats frames above this threshold frame
        if eng is None:
            eng = WriteEngine(ny, nx, itemsize, 8 * itemsize, 2, 1, 0, 0, 1, max_frames=min(16, max(1, n_stats_frames)),
                              device=device)
        eng.set_threshold(t, 0)
        n_events = 0
        n_fg = 0
        for b0 in range(0, n_stats_frames, eng.max_frames):
            maps, _, counts = eng.reduce(stats[b0:b0 + eng.max_frames])
            n_events += int(np.sum(counts))
            n_fg += sum(int(np.unpackbits(np.frombuffer(m, dtype=np.uint8)).sum()) for m in maps)
        avg_n_events = n_events / n_stats_frames
        avg_p_foreground_pixels = n_fg / n_pixels_in_frame / n_stats_frames
        print("Avg. prop. foreground pixels for sigma=" + str(i) + " is: " + str(avg_p_foreground_pixels))
        print("Avg. electron count for sigma=" + str(i) + " is: " + str(avg_n_events))
        print("Avg. dose rate for sigma=" + str(i) + " is: " + str(avg_n_events / n_pixels_in_frame))
        print("")
        if use_acc and i == sigma_acc:
            expected_n_events = int(np.ceil(nFrames * (avg_n_events / n_pixels_in_frame)))
            print(expected_n_events)
            if expected_n_events < 2:
                print("Unable to compute accurate thresholds: too few events in dataset")
            else:
                acc_t = pixel_thresholds(d, _m, expected_n_events, device=device, as_run=acc_as_run)
                with np.errstate(invalid='ignore'):
                    acc_t.astype(dtype).tofile(os.path.join(savepath, filename_prefix + "_dark_ref_" + str(i) + "A.bin"))
                print(acc_t)
                acc = acc_t
    return {'median': _m, 'std': _stds, 'sigma': _fit_std, 'thresholds': out, 'accurate_thresholds': acc}

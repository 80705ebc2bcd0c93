#!/usr/bin/env python
"""bench.py -- headline benchmark of the ReCoDe reduce-and-compress hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--level L] [--frames F] [--launches M]

Metric (BASELINE.json): frames/s (and input GB/s) of L2 reduce + deflate on synthetic 4096 x 4096 uint16
frames, 12-bit, zlib level 1.  A "step" is one pass of the hot path over one batch of 256 device-resident frames per
GPU, issued as M = 8 launches of rc_reduce_compress over F = 32 frames each, round-robin on the engine's
pipelined slots.  One process per GPU; under torchrun every rank processes its own frame range (frames are
independent: weak scaling, no data-path collective on the write path) and rank 0 prints ONE JSON line.

  value          whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e            same metric through the host-buffer API: pinned host frames -> H2D -> kernels -> D2H of the
                 records, all inside the timed region; h2d_ceiling_gbs = plain pinned cudaMemcpyAsync on all ranks at once
  roofline       dominant kernel (k_reduce_tiles_bulk: the only kernel that reads the raw frames): algorithmic
                 bytes per launch = F * ny*nx*itemsize, over its CUDA-event duration, against the measured
                 HBM copy bandwidth in MEASURED_PEAKS.json
  parity_checked the records of the last launch on every slot, inflated by stock zlib, equal the CPU oracle's streams
                 (checked after the timed region)
  read           BASELINE config 5 beside it: this rank's L2 part file -> live-view sum + NCCL all-reduce, and -> dense
                 frames, with its own roofline (dense bytes written) and cpu_baseline (zlib inflate + oracle unpack)
  cpu_baseline   the CPU oracle port (oracle/, kind "port": the reference's own L2 writer does not execute,
                 SURVEY 0.1) with scipy-equivalent labelling + stock zlib level 1 on all host cores, on a
                 bounded sample of the same frames
  --impl reference   times that CPU path as the measured arm (rank 0 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NY = NX = 4096
BIT_DEPTH = 12
EPS = 20
KIND = {1: 'l1', 2: 'l2', 3: 'l1', 4: 'l4'}


def ncu_traffic(level, frames):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed
    `ncu --set full` capture of this configuration (profiles/ncu_traffic.json), else None"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
            return json.load(f).get('L%d_F%d' % (level, frames))
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def src_dtype():
    """the source dtype misc.map_dtype selects for the bit depth (pyrecode/misc.py:41-71): bytes up to 8 bits"""
    return (np.uint8, 1) if BIT_DEPTH <= 8 else (np.uint16, 2)


def make_inputs(level, n_distinct, seed=1234):
    from pyrecode_b200.synth import synth_dark, synth_frames
    dt, _ = src_dtype()
    dark = synth_dark(NY, NX)
    if dt == np.uint8:
        # an 8-bit detector: the 12-bit model (dark level, read noise, event amplitudes) seen through a converter that
        # drops the 4 low bits -- the same events and the same occupancy, one byte per pixel (eps_of() scales too)
        frames = synth_frames(KIND[level], n_distinct, NY, NX, dark, seed=seed, bit_depth=12)
        sh = 12 - min(BIT_DEPTH, 8)
        return (dark >> 4).astype(np.uint8), ((frames >> 4) >> (4 - sh if sh > 4 else 0)).astype(np.uint8)
    frames = synth_frames(KIND[level], n_distinct, NY, NX, dark, seed=seed, bit_depth=BIT_DEPTH)
    return dark, frames


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port on all host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    from oracle import oracle as orc
    level, frames, thr, reps = args
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(reps):
        for f in frames:
            m, v, n = orc.reduce_frame(f, thr, level, BIT_DEPTH)
            rec = orc.build_record(0, level, 1, m, v, 1)
            nbytes += len(rec)
    return time.perf_counter() - t0, nbytes


def cpu_run(level, dark, frames, frames_per_core, cores=None):
    """-> (frames/s, cores, total frames, seconds)."""
    import multiprocessing as mp
    from oracle import oracle as orc
    orc.lib()
    cores = cores or os.cpu_count() or 1
    thr = orc.make_threshold(dark.astype(np.uint16), eps_of()).astype(np.uint16)
    frames = [np.ascontiguousarray(f, dtype=np.uint16) for f in frames]
    jobs = []
    for c in range(cores):
        sel = [frames[(c + i) % len(frames)] for i in range(frames_per_core)]
        jobs.append((level, sel, thr, 1))
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(level, [frames[0]], thr, 1)] * cores)      # warm the workers
        res = pool.map(_cpu_worker, jobs)
    slowest = max(r[0] for r in res)
    total = cores * frames_per_core
    return total / slowest, cores, total, slowest


def eps_of():
    return EPS if BIT_DEPTH > 8 else 2


def _cpu_read_worker(args):
    import zlib
    from oracle import oracle as orc
    recs, reps = args
    t0 = time.perf_counter()
    acc = np.zeros((NY, NX), dtype=np.uint32)
    for _ in range(reps):
        for cm, cv in recs:
            m = zlib.decompress(cm)
            v = zlib.decompress(cv)
            d = orc.unpack_dense(NY, NX, BIT_DEPTH, m, v, 2)
            acc += d
    return time.perf_counter() - t0


def cpu_read_run(records, frames_per_core, cores=None):
    """CPU baseline of the read path: stock zlib inflate of both streams + the oracle's dense unpack
    + the live-view accumulation, one process per core.  -> (frames/s, cores, frames, seconds)"""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [([records[(c + i) % len(records)] for i in range(frames_per_core)], 1) for c in range(cores)]
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_read_worker, [([records[0]], 1)] * cores)
        res = pool.map(_cpu_read_worker, jobs)
    slowest = max(res)
    return cores * frames_per_core / slowest, cores, cores * frames_per_core, slowest


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop_evt = threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(',')]
                self.samples.append((float(p[0]), float(p[1])))
                for nme, v in zip(names, p[2:6]):
                    if v.lower().startswith('active'):
                        self.reasons.add(nme)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': sorted(self.reasons)}
        return {'sm_mhz': float(np.median([s[0] for s in self.samples])), 'sm_max_mhz': self.samples[0][1],
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def bind_numa(local_rank):
    """run this rank's host threads (and so its pinned allocations, first touch) on the CPUs next to its GPU"""
    try:
        from pyrecode_b200.distributed import bind_to_gpu_cpus
        return bind_to_gpu_cpus(local_rank)
    except Exception:
        return None


def parse_record(r, level):
    hdr = np.frombuffer(r[:16 if level <= 2 else 8], dtype='<u4')
    if level <= 2:
        return int(hdr[0]), bytes(r[16:16 + hdr[1]]), bytes(r[16 + hdr[1]:16 + hdr[1] + hdr[2]])
    return int(hdr[0]), bytes(r[8:8 + hdr[1]]), None


def parity_check(eng, level, dark, frames, F, ids_of_slot):
    """the records every slot holds from its last launch vs the CPU oracle (the checker; outside any timed region)"""
    import zlib
    from oracle import oracle as orc
    thr = orc.make_threshold(dark.astype(np.uint16), eps_of()).astype(np.uint16)
    expect = [orc.reduce_frame(np.ascontiguousarray(f, dtype=np.uint16), thr, level, BIT_DEPTH) for f in frames]
    checked = 0
    ours_bytes = zlib1_bytes = 0
    z1 = [len(zlib.compress(m, 1)) + (len(zlib.compress(v, 1)) if level <= 2 else 0) for m, v, _ in expect]
    for k, sl in enumerate(eng.slots):
        if ids_of_slot[k] is None:
            continue
        if int(sl.status.cpu()[0]) != 0:
            return False, 'slot %d status' % k
        offs = sl.offsets.cpu().numpy()
        counts = sl.counts.cpu().numpy()
        rec = memoryview(sl.records[:int(offs[F])].cpu().numpy())
        for i in range(F):
            fid, cm, cv = parse_record(rec[int(offs[i]):int(offs[i + 1])], level)
            m, v, n = expect[i % len(expect)]
            if fid != ids_of_slot[k] + i or zlib.decompress(cm) != m or int(counts[i]) != n:
                return False, 'slot %d frame %d: id / map / count' % (k, i)
            if level <= 2 and zlib.decompress(cv) != v:
                return False, 'slot %d frame %d: value stream' % (k, i)
            checked += 1
            ours_bytes += len(cm) + (len(cv) if cv is not None else 0)
            zlib1_bytes += z1[i % len(z1)]
    return True, '%d records of %d slots' % (checked, len(eng.slots)), ours_bytes / max(zlib1_bytes, 1)


# ---------------------------------------------------------------------------------------------
# read path (BASELINE config 5)
# ---------------------------------------------------------------------------------------------
def read_leg(args, rank, world, local_rank, dev, dark, frames, want_cpu):
    """Every rank decodes its own L2 part file (frame-sharded stream) to the summed live-view image, the images are
    all-reduced over NCCL; then the same file to dense frames.  File reads are inside the timed regions.
    -> dict (rank 0) / None"""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from pyrecode_b200.params import InputParams
    from pyrecode_b200.recode_reader import ReCoDeReader
    from pyrecode_b200.recode_writer import ReCoDeWriter
    from pyrecode_b200 import distributed as rcd
    nz = args.read_frames
    steps = max(1, args.read_steps)
    _, isz = src_dtype()
    frame_bytes = NY * NX * isz
    tmp = tempfile.mkdtemp(prefix='recode_bench_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    try:
        ip = InputParams()
        nz = max(256, nz // 256 * 256)
        for k, v in dict(l4_centroiding=0, source_file_type=0, num_frames=256, source_header_length=0,
                         calibration_frame_offset=0, compression_scheme=0, calibration_file_type=0, compression_level=1,
                         l2_statistics=0, calibration_threshold_epsilon=eps_of(), frame_offset=0, num_threads=1,
                         rc_operation_mode=1, num_calibration_frames=1, reduction_level=2, keep_calibration_data=1,
                         source_bit_depth=BIT_DEPTH, target_bit_depth=BIT_DEPTH, keep_part_files=0, num_rows=NY,
                         num_cols=NX, source_data_type=0, target_data_type=0).items():
            ip._param_map[k] = v
        w = ReCoDeWriter('rb', dark_data=dark[None], output_directory=tmp, input_params=ip, mode='batch', node_id=0,
                         device=local_rank)
        w.start()
        per_run = ip._param_map['num_frames']
        chunk = np.stack([frames[i % len(frames)] for i in range(per_run)])
        for _ in range(nz // per_run):                # the part file grows by one chunk of frames per run()
            w.run(chunk)
        w.close()
        del w, chunk
        path = os.path.join(tmp, 'rb.rc2_part000')
        fsize = os.path.getsize(path)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # one reader for the whole run, rewound every step: its engines (contexts, device buffers, pinned staging)
        # are set up once, as for a long acquisition decoded batch after batch
        r = ReCoDeReader(path, is_intermediate=True, device=local_rank, bulk_frames=args.read_batch,
                         bulk_inflight=args.read_inflight)
        r.open(print_header=False)
        view = torch.zeros(NY * NX, dtype=torch.int32, device=dev)
        ar0, ar1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def one(what):
            r.rewind()
            if what == 'sum':
                view.zero_()
                ids, total = r.sum_frames(nz, total=view)
                ar0.record()
                rcd.allreduce_view(total)
                ar1.record()
                out = int(total[:1024].sum().item())          # device -> host read of the result
            else:
                ids, dense = r.read_frames_dense(min(nz, args.read_dense_frames))
                out = int(dense[0, 0, :8].sum().item())
                del dense
            return len(ids), dict(r.bulk_stats)

        res = {}
        ar_ms = []
        for what in ('sum', 'dense'):
            for _ in range(2):
                one(what)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            n = 0
            for _ in range(steps):
                k, st = one(what)
                n += k
                if what == 'sum':
                    torch.cuda.synchronize()
                    ar_ms.append(ar0.elapsed_time(ar1))
            e1.record()
            barrier()
            dt = max(time.perf_counter() - t0, e0.elapsed_time(e1) / 1e3)
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[what] = (world * n / float(t[0]), float(t[0]) / steps, st)
        # the collective alone: all ranks enter together (inside a view its events also hold the wait for the slowest rank)
        ar_iso = []
        if world > 1:
            for _ in range(6):
                barrier()
                ar0.record()
                rcd.allreduce_view(view)
                ar1.record()
                torch.cuda.synchronize()
                ar_iso.append(ar0.elapsed_time(ar1))
        # parity of the read leg (outside the timed regions; the oracle is the checker): this rank's live-view sum before
        # the all-reduce, and the first dense frames, against stock zlib + the oracle's unpack of the file's own records
        read_parity, read_parity_err = None, None
        if rank == 0 and not args.no_parity:
            try:
                import zlib
                from oracle import oracle as orc
                r.rewind()
                view.zero_()
                r.sum_frames(nz, total=view)
                got = view.cpu().numpy().astype(np.int64).reshape(NY, NX)
                r.rewind()
                _, dn = r.read_frames_dense(len(frames))
                dn = dn.cpu().numpy()
                rr = ReCoDeReader(path, is_intermediate=True, device=local_rank)
                rr.open(print_header=False)
                want = np.zeros((NY, NX), dtype=np.int64)
                ok = nz % len(frames) == 0 and per_run % len(frames) == 0
                for j in range(len(frames)):                       # the file cycles through the distinct frames
                    fr = next(iter(rr.get_next_frame_raw().values()))['data']
                    d = orc.unpack_dense(NY, NX, BIT_DEPTH, zlib.decompress(bytes(fr['binary_map'])),
                                         zlib.decompress(bytes(fr['pixvals'])), 2)
                    ok = ok and np.array_equal(dn[j], d)
                    want += d.astype(np.int64) * (nz // len(frames))
                rr.close()
                read_parity = bool(ok and np.array_equal(got, want))
                del dn
            except Exception as e:                              # a checker problem must not lose the bench line
                read_parity, read_parity_err = False, repr(e)[:200]
        gpu = r.decode_stage_ms()
        # the per-frame API (get_next_frame -> scipy COO + summary statistics), as a user loop would call it
        seq_fps = None
        if rank == 0:
            rr = ReCoDeReader(path, is_intermediate=True, device=local_rank)
            rr.open(print_header=False)
            rr.get_next_frame()
            t0 = time.perf_counter()
            k = 0
            for _ in range(48):
                if rr.get_next_frame() is None:
                    break
                k += 1
            seq_fps = k / (time.perf_counter() - t0)
            rr.close()
        # records for the CPU baseline before the file goes away
        recs = None
        if want_cpu and rank == 0:
            rr = ReCoDeReader(path, is_intermediate=True, device=local_rank)
            rr.open(print_header=False)
            recs = []
            for _ in range(min(4, nz)):
                d = rr.get_next_frame_raw()
                fr = next(iter(d.values()))['data']
                recs.append((bytes(fr['binary_map']), bytes(fr['pixvals'])))
            rr.close()
        r.close()
        if rank != 0:
            return None
        # reference-written files hold ONE block sequence per stream (stock zlib, matches at any distance): no chunk
        # parallelism, one decoding lane per stream through a 32 KiB shared-memory history ring
        foreign = None
        if recs:
            import zlib
            from pyrecode_b200._native import Context
            from pyrecode_b200.engine import _stage_streams
            maps = [zlib.decompress(cm) for cm, _ in recs[:4]]
            comp = [zlib.compress(m, 1) for m in maps]
            t0 = time.perf_counter()
            for c_ in comp:
                zlib.decompress(c_)
            core_ms = 1e3 * (time.perf_counter() - t0) / len(comp)
            fctx = Context(local_rank)
            cap = len(maps[0])
            fstride = (cap + 15) // 16 * 16 + 16
            foreign = {'inflated_bytes_each': cap, 'stock_zlib_one_core_ms_per_stream': core_ms, 'runs': [], 'ok': True,
                       'note': 'one decoding lane per stream (k_inflate_serial, 32 KiB shared-memory history ring): the '
                               'latency of one stream is that of the whole batch, so throughput grows with the streams '
                               'in flight (a 64-frame batch of a reference-written L1/L2 file holds 128); compressed '
                               'input and inflated output resident in HBM, CUDA events'}
            for n_st in (32, 512):
                streams = [comp[i % len(comp)] for i in range(n_st)]
                d_in, d_off, d_sz, _ = _stage_streams(fctx, streams)
                fout = fctx.empty(n_st * fstride + 64)
                fob = fctx.zeros(n_st, torch.int32)
                fst = fctx.zeros(n_st, torch.int32)
                best = None
                for rep in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fws = fctx.inflate_zlib(d_in, d_off, d_sz, n_st, fout, fstride, fob, fst)
                    e1.record()
                    torch.cuda.synchronize()
                    ms_ = e0.elapsed_time(e1)
                    best = ms_ if best is None else min(best, ms_)
                    del fws
                oh = fout[(n_st - 1) * fstride:(n_st - 1) * fstride + cap].cpu().numpy().tobytes()
                ok = bool(not fst.cpu().numpy().any()) and oh == maps[(n_st - 1) % len(maps)]
                foreign['ok'] = foreign['ok'] and ok
                foreign['runs'].append({'streams': n_st, 'gpu_ms': best, 'streams_per_s': n_st / best * 1e3,
                                        'inflated_gb_s': n_st * cap / best / 1e6,
                                        'host_cores_equivalent': n_st / best * core_ms})
                del d_in, d_off, d_sz, fout, fob, fst
            fctx.close()
        peak, peak_src = measured_peak()
        out = {'workload': 'read path (BASELINE config 5): %d-frame L2 part file per GPU on tmpfs, %d-bit, file reads '
                           'inside the timed region; live-view sum all-reduced (NCCL, 64 MiB int32) once per view' % (nz, BIT_DEPTH),
               'live_view_frames_per_s': res['sum'][0], 'ms_per_view': 1e3 * res['sum'][1],
               'allreduce_ms': float(np.median(ar_ms)) if ar_ms else None,
               'allreduce_ms_isolated': float(np.median(ar_iso[1:])) if len(ar_iso) > 1 else None,
               'dense_frames_per_s': res['dense'][0], 'dense_output_gb_s': res['dense'][0] * frame_bytes / 1e9,
               'file_bytes': fsize, 'steps': steps, 'frames_per_view': nz,
               'get_next_frame_fps': seq_fps, 'foreign_zlib_streams': foreign, 'parity_checked': read_parity, 'parity_error': read_parity_err,
               'host_time_split_ms_last_view': {k: (1e3 * v if k.endswith('_s') else v) for k, v in res['sum'][2].items()},
               'e2e': {'value': res['sum'][0], 'unit': 'frames/s', 'h2d_bytes_per_step': fsize, 'd2h_bytes_per_step': 8}}
        if gpu and gpu.get('frames'):
            # dominant kernel of the decode, timed alone on one batch: k_inflate_lanes over the batch's map streams;
            # algorithmic bytes = the dense frames the batch reconstructs
            kframes, kms = gpu['frames'], max(gpu['lanes_ms'], 1e-6)
            ach = kframes * frame_bytes / (kms / 1e3) / 1e9
            out['roofline'] = {'bound': 'hbm', 'kernel': 'k_inflate_lanes', 'achieved': ach, 'peak': peak, 'unit': 'GB/s',
                               'frac': ach / peak, 'traffic': None, 'peak_source': peak_src, 'kernel_ms': kms,
                               'algorithmic_bytes_per_launch': kframes * frame_bytes,
                               'note': 'a serial chain per 16 KiB chunk (latency-bound, a few per cent of the warp slots): '
                                       'batches in flight overlap, so the pipeline runs faster than this kernel alone'}
            out['gpu_stage_ms_per_batch'] = gpu
            ud = max(gpu['unpack_dense_ms'], 1e-6)
            out['unpack_dense_gb_s'] = kframes * frame_bytes / (ud / 1e3) / 1e9
        out['hbm_roofline_frac_whole_path'] = res['dense'][0] / world * frame_bytes / 1e9 / peak
        if recs:
            fps, cores, total, secs = cpu_read_run(recs, args.cpu_read_frames_per_core)
            out['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                   'sample': '%d frames (%d per core): stock zlib inflate of both streams + oracle dense '
                                             'unpack + uint32 accumulation, %.1f s' % (total, args.cpu_read_frames_per_core, secs)}
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--level', type=int, default=2)
    ap.add_argument('--frames', type=int, default=32, help='frames per launch of rc_reduce_compress')
    ap.add_argument('--launches', type=int, default=8, help='launches per step (a step = frames x launches per GPU)')
    ap.add_argument('--distinct', type=int, default=4, help='distinct synthetic frames generated on the host')
    ap.add_argument('--cpu-frames-per-core', type=int, default=96,
                    help='frames per host core of the cpu_baseline sample (about 10 s of CPU work)')
    ap.add_argument('--cpu-read-frames-per-core', type=int, default=48)
    ap.add_argument('--ref-frames-per-core', type=int, default=4, help='frames per core per step of --impl reference')
    ap.add_argument('--slots', type=int, default=3, help='batches in flight (each on its own CUDA stream)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-read', action='store_true', help='skip the read-path sub-record (BASELINE config 5)')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--also-levels', default='1,4', help='other reduction levels timed briefly beside the headline one')
    ap.add_argument('--bit-depth', type=int, default=12, help='source / target bit depth (SURVEY 8d: 8, 12, 16)')
    ap.add_argument('--mode', default='write', choices=['write', 'read'],
                    help="read: only BASELINE config 5 -- an L2 part file per GPU -> live-view sum (+ NCCL all-reduce) and "
                         "dense frames through ReCoDeReader's bulk calls")
    ap.add_argument('--read-frames', type=int, default=1024, help='frames in the part file of the read leg (a multiple of 256)')
    ap.add_argument('--read-steps', type=int, default=5)
    ap.add_argument('--read-dense-frames', type=int, default=1024,
                    help='frames per dense read (32 MiB of device memory each at 4096 x 4096 x 16 bit)')
    ap.add_argument('--read-batch', type=int, default=64, help='frames per decode batch of the read leg')
    ap.add_argument('--read-inflight', type=int, default=6, help='decode batches in flight')
    args = ap.parse_args()
    global BIT_DEPTH
    BIT_DEPTH = args.bit_depth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    level = args.level
    dt, isz = src_dtype()
    frame_bytes = NY * NX * isz
    F, M = args.frames, args.launches
    workload = ('L%d reduce + deflate, 4096x4096 %s, %d-bit, zlib level 1, synthetic %s frames (SURVEY 8d), '
                '%d frames per step per GPU (%d launches of %d), frame-sharded'
                % (level, np.dtype(dt).name, BIT_DEPTH, KIND[level], F * M, M, F))
    config = {'workload': workload, 'reduction_level': level, 'frames_per_step_per_gpu': F * M,
              'frames_per_launch': F, 'launches_per_step': M,
              'frame_shape': [NY, NX], 'bit_depth': BIT_DEPTH, 'compression_level': 1,
              'cache': 'inputs larger than L2 (%d MiB per launch)' % (F * frame_bytes >> 20),
              'batches_in_flight': args.slots,
              'clocks_sampled': 'timed region plus 1.5 s of the same steps, untimed'}
    metric = 'frames/s, 4096x4096 L%d reduce+deflate' % level

    # ----------------------------------------------------------------------------------- reference arm
    if args.impl == 'reference':
        if rank != 0:
            return
        dark, frames = make_inputs(level, args.distinct)
        per_step = []
        for s in range(args.warmup + args.steps):
            fps, cores, total, secs = cpu_run(level, dark, frames, args.ref_frames_per_core)
            if s >= args.warmup:
                per_step.append((total, secs))
        tot = sum(p[0] for p in per_step)
        sec = sum(p[1] for p in per_step)
        value = tot / sec
        line = {'impl': 'reference', 'metric': metric, 'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sec / max(len(per_step), 1),
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u%d' % (8 * isz),
                'data': 'synthetic', 'config': config, 'input_gb_s': value * frame_bytes / 1e9,
                'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                 'sample': '%d frames per step (%d per core), oracle C port + stock zlib level 1; '
                                           'the reference cannot execute L2/L4 (SURVEY 0.1)' % (total, args.ref_frames_per_core)},
                'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    # ----------------------------------------------------------------------------------- our arm
    cpus = bind_numa(local_rank)
    import torch
    import torch.distributed as dist
    from pyrecode_b200.engine import WriteEngine

    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if args.mode == 'read':
        dark, frames = make_inputs(2, args.distinct, seed=1234 + rank)
        rd = read_leg(args, rank, world, local_rank, dev, dark, frames, not args.no_cpu)
        if rank == 0:
            line = {'metric': 'frames/s, 4096x4096 L2 part file -> live-view sum (NCCL all-reduce across ranks)',
                    'value': rd['live_view_frames_per_s'], 'unit': 'frames/s', 'n_gpus': world, 'steps': rd['steps'],
                    'warmup': 2, 'ms_per_step': rd['ms_per_view'], 'higher_is_better': True, 'scaling': 'weak',
                    'vs_baseline': None, 'dtype': 'u%d' % (8 * isz), 'data': 'synthetic',
                    'config': {'workload': rd['workload'], 'frames_per_step_per_gpu': args.read_frames}}
            line.update({k: v for k, v in rd.items() if k not in ('workload', 'steps')})
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    # every rank generates its own frame range of the stream (seed offset = rank: different frames per GPU)
    dark, frames = make_inputs(level, args.distinct, seed=1234 + rank)
    eng = WriteEngine(NY, NX, isz, BIT_DEPTH, level, 1, 0, 0, 1, max_frames=F, device=local_rank,
                      records_capacity=F * (frame_bytes // 4), n_slots=args.slots)
    eng.set_threshold(dark, eps_of())
    host = torch.empty((F, NY, NX), dtype=torch.uint16 if isz == 2 else torch.uint8).pin_memory()
    hv = host.numpy()
    for i in range(F):
        hv[i] = frames[i % len(frames)]
    d_frames = host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first_id = rank * F * M * (args.warmup + args.steps)

    # ---- device-resident timing (value) + dominant-kernel timing (roofline)
    # Launches are issued round-robin on the engine's slots (one CUDA stream each), the way the writer keeps
    # batches in flight; the timed region is bracketed by events on the default stream that all slot streams
    # fork from / join to.
    nsl = len(eng.slots)
    cur = torch.cuda.current_stream()
    last_ids = [None] * nsl

    def run_steps(n_steps, id0):
        for sl in eng.slots:
            sl.stream.wait_stream(cur)
        for s in range(n_steps * M):
            sl = eng.slots[s % nsl]
            with torch.cuda.stream(sl.stream):
                eng.launch(d_frames, F, id0 + s * F, s % nsl)
            last_ids[s % nsl] = id0 + s * F
        for sl in eng.slots:
            cur.wait_stream(sl.stream)

    eng.ctx.profile_enable(True)
    run_steps(args.warmup, first_id)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stage_ms = np.zeros(4)
    launches0 = sum(sl.ctx.launch_count() for sl in eng.slots)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps, first_id + args.warmup * F * M)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sum(sl.ctx.launch_count() for sl in eng.slots) - launches0
    # parity of what the timed region left in the slots' buffers, before anything else touches them
    parity = None
    if not args.no_parity and rank == 0:
        parity = parity_check(eng, level, dark, frames, F, last_ids)
    # the timed region lasts a fraction of a second, one nvidia-smi query ~0.1 s: keep the same steps running
    # (untimed) for about 1.5 s so that the clock sampler sees the GPU under this load
    if rank == 0:
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            run_steps(2, first_id)
            torch.cuda.synchronize()
    # stage split: re-run a few launches one at a time on slot 0 with per-launch readback of the stage marks
    # the same stages as the batches in flight see them (events between the stages of every slot's last launch):
    # how long the streaming kernel takes while the other batches' labelling / encoder kernels share the SMs
    stage_pipe = None
    if nsl > 1:
        for sl in eng.slots:
            sl.ctx.profile_enable(1)
        run_steps(2, first_id)
        torch.cuda.synchronize()
        stage_pipe = [float(x) for x in np.mean([sl.ctx.profile_read()[:4] for sl in eng.slots], axis=0)]
        for sl in eng.slots[1:]:
            sl.ctx.profile_enable(0)
    # (outside the timed region; the dominant kernel is timed alone here, which is what `roofline` reports)
    nprof = 5
    eng.ctx.set_pipelined(False)              # one batch at a time: every kernel gets the whole GPU
    eng.ctx.profile_enable(2)
    detail = None
    for s in range(nprof):
        eng.launch(d_frames, F, 0)
        stage_ms += np.array(eng.ctx.profile_read()[:4])
        d = np.array(eng.ctx.profile_read_detail())
        detail = d if detail is None else detail + d
    stage_ms /= nprof
    detail = [float(x) / nprof for x in detail] if detail is not None else None
    eng.ctx.profile_enable(1)
    eng.ctx.set_pipelined(nsl > 1)
    st = int(eng.status.cpu()[0])
    offs = eng.offsets.cpu().numpy()
    rec_bytes = int(offs[F])
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    value = world * F * M * args.steps / (ms_total / 1e3)

    # ---- end to end through host buffers
    e2e = None
    if not args.no_e2e:
        # through the host-buffer API the writer uses: pinned host frames -> H2D -> kernels -> D2H of the records,
        # every launch, with `slots` batches in flight
        def e2e_steps(n_steps, id0):
            h2d = d2h = 0
            pending = []
            for s in range(n_steps * M):
                pending.append(eng.submit(host, first_frame_id=id0 + s * F))
                if len(pending) == nsl:
                    _, _, _, a, b = eng.collect(pending.pop(0))
                    h2d += a
                    d2h += b
            while pending:
                _, _, _, a, b = eng.collect(pending.pop(0))
                h2d += a
                d2h += b
            return h2d // n_steps, d2h // n_steps

        e2e_n = max(2, min(args.steps, 6))
        e2e_steps(1, 0)
        barrier()
        e0.record()
        h2d, d2h = e2e_steps(e2e_n, first_id)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        # the box's ceiling for this traffic: the same pinned buffer copied by plain cudaMemcpyAsync on every rank at once
        dst = eng.slots[0].frames_dev
        for _ in range(2):
            dst.copy_(host, non_blocking=True)
        barrier()
        e0.record()
        for _ in range(8):
            dst.copy_(host, non_blocking=True)
        e1.record()
        barrier()
        ms_cp = e0.elapsed_time(e1)
        t = torch.tensor([ms_e2e, ms_cp], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ceil_gbs = 8 * F * frame_bytes / (float(t[1]) / 1e3) / 1e9
        e2e = {'value': world * F * M * e2e_n / (float(t[0]) / 1e3), 'unit': 'frames/s', 'steps': e2e_n,
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'h2d_ceiling_gbs': ceil_gbs, 'h2d_ceiling_frames_per_s': world * ceil_gbs * 1e9 / frame_bytes,
               'h2d_ceiling_how': 'per rank, all ranks at once: 8 x cudaMemcpyAsync of the %d MiB pinned batch, CUDA events, '
                                  'max over ranks' % (F * frame_bytes >> 20),
               'cpu_affinity': cpus}

    del d_frames
    eng = None
    torch.cuda.empty_cache()

    # ---- the other reduction levels of BASELINE.json's configs (L1: config 2, L4: config 4), same harness, short
    others = {}
    for lv in [int(x) for x in args.also_levels.replace("'", '').replace('"', '').split(',') if x.strip()]:
        if lv == level:
            continue
        dk, fr = make_inputs(lv, args.distinct, seed=1234 + rank)
        e2 = WriteEngine(NY, NX, isz, BIT_DEPTH, lv, 1, 0, 0, 1, max_frames=F, device=local_rank,
                         records_capacity=F * (frame_bytes // 4), n_slots=args.slots)
        e2.set_threshold(dk, eps_of())
        for i in range(F):
            hv[i] = fr[i % len(fr)]
        d2 = host.to(dev)
        ids2 = [None] * len(e2.slots)

        def steps2(n_steps, id0):
            for sl in e2.slots:
                sl.stream.wait_stream(cur)
            for q in range(n_steps * M):
                sl = e2.slots[q % len(e2.slots)]
                with torch.cuda.stream(sl.stream):
                    e2.launch(d2, F, id0 + q * F, q % len(e2.slots))
                ids2[q % len(e2.slots)] = id0 + q * F
            for sl in e2.slots:
                cur.wait_stream(sl.stream)

        steps2(2, 0)
        barrier()
        e0.record()
        steps2(5, 7000)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        v2 = world * F * M * 5 / (float(t2[0]) / 1e3)
        par2 = parity_check(e2, lv, dk, fr, F, ids2) if (rank == 0 and not args.no_parity) else None
        others['L%d' % lv] = {'value': v2, 'unit': 'frames/s', 'steps': 5, 'ms_per_launch': float(t2[0]) / (5 * M),
                              'input_gb_s': v2 * frame_bytes / 1e9,
                              'hbm_roofline_frac_whole_path': v2 / world * frame_bytes / 1e9 / measured_peak()[0],
                              'parity_checked': bool(par2[0]) if par2 else False,
                              'workload': 'L%d reduce + deflate, synthetic %s frames, same geometry / batching' % (lv, KIND[lv])}
        del e2, d2
        torch.cuda.empty_cache()

    rd = None
    if not args.no_read:
        dark2, frames2 = (dark, frames) if level == 2 else make_inputs(2, args.distinct, seed=1234 + rank)
        rd = read_leg(args, rank, world, local_rank, dev, dark2, frames2, not args.no_cpu)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    k1_ms = float(stage_ms[0])
    achieved = F * frame_bytes / (k1_ms / 1e3) / 1e9
    line = {'metric': metric, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u%d' % (8 * isz), 'data': 'synthetic', 'config': config,
            'input_gb_s': value * frame_bytes / 1e9,
            'hbm_roofline_frac_whole_path': value / world * frame_bytes / 1e9 / peak,
            'ms_per_launch': ms_total / (args.steps * M),
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'parity_checked': bool(parity[0]) if parity else False, 'parity_detail': parity[1] if parity else 'skipped',
            'compressed_bytes_vs_zlib1': parity[2] if parity and len(parity) > 2 else None,
            'roofline': {'bound': 'hbm', 'kernel': 'k_reduce_tiles_bulk', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': ncu_traffic(level, F), 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': F * frame_bytes, 'kernel_ms': k1_ms},
            'stage_ms_per_launch': dict(zip(['threshold_pack_compact', 'reduce_rest', 'deflate', 'assemble'],
                                            [float(x) for x in stage_ms])),
            'stage2_kernel_ms_per_launch': detail,
            'stage_ms_batches_in_flight': dict(zip(['threshold_pack_compact', 'reduce_rest', 'deflate', 'assemble'],
                                                   stage_pipe)) if stage_pipe else None,
            'record_bytes_per_frame': rec_bytes / F, 'status': st}
    if others:
        line['other_levels'] = others
    if rd is not None:
        line['read'] = rd
    if not args.no_cpu:
        fps, cores, total, secs = cpu_run(level, dark, frames, args.cpu_frames_per_core)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                'sample': '%d frames (%d per core) of the same synthetic workload, oracle C port + '
                                          'stock zlib level 1, %.1f s' % (total, args.cpu_frames_per_core, secs)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the ReCoDe reduce-and-compress hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--level L] [--frames F]

Metric (BASELINE.json): frames/s (and input GB/s) of L2 reduce + deflate on synthetic 4096 x 4096 uint16
frames, 12-bit, zlib level 1.  A "step" is one pass of the hot path (rc_reduce_compress) over one batch of F
device-resident frames.  One process per GPU; under torchrun every rank processes its own frame range
(frames are independent: weak scaling, no data-path collective) and rank 0 prints ONE JSON line.

  value          whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e            same metric through the host-buffer API: pinned host frames -> H2D -> kernels -> D2H of the
                 records, all inside the timed region
  roofline       dominant kernel (k_reduce_tiles: the only kernel that reads the raw frames): algorithmic
                 bytes per launch = F * ny*nx*itemsize, over its CUDA-event duration, against the measured
                 HBM copy bandwidth in MEASURED_PEAKS.json
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


# dram__bytes_read.sum + dram__bytes_write.sum of k_reduce_tiles per launch from the committed `ncu --set full`
# capture (profiles/r01_ncu_k_reduce_tiles_bulk_v6.txt): keyed by (level, frames per step)
NCU_TRAFFIC = {(2, 32): 1107738000 + 131561216}


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def make_inputs(level, n_distinct, seed=1234):
    from pyrecode_b200.synth import synth_dark, synth_frames
    dark = synth_dark(NY, NX)
    frames = synth_frames(KIND[level], n_distinct, NY, NX, dark, seed=seed, bit_depth=BIT_DEPTH)
    return dark, frames


# ---------------------------------------------------------------------------------------------
# CPU arm: oracle port on all host cores
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import zlib
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
    thr = orc.make_threshold(dark, EPS)
    jobs = []
    for c in range(cores):
        sel = [frames[(c + i) % len(frames)] for i in range(frames_per_core)]
        jobs.append((level, sel, thr, 1))
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [(level, [frames[0]], thr, 1)] * cores)      # warm the workers
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        wall = time.perf_counter() - t0
    slowest = max(r[0] for r in res)
    total = cores * frames_per_core
    return total / max(slowest, wall * 0 + slowest), cores, total, slowest


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


def run_read(args, rank, world, local_rank, dev):
    """BASELINE config 5: every rank decodes its own L2 part file (frame-sharded stream) to the summed live-view image,
    the images are all-reduced over NCCL; then the same file to dense frames.  File reads are inside the timed region."""
    import shutil
    import tempfile
    import torch
    import torch.distributed as dist
    from pyrecode_b200.params import InputParams
    from pyrecode_b200.recode_reader import ReCoDeReader
    from pyrecode_b200.recode_writer import ReCoDeWriter
    from pyrecode_b200 import distributed as rcd
    nz = args.read_frames
    dark, frames = make_inputs(2, args.distinct, seed=1234 + rank)
    tmp = tempfile.mkdtemp(prefix='recode_bench_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    try:
        ip = InputParams()
        for k, v in dict(l4_centroiding=0, source_file_type=0, num_frames=nz, source_header_length=0,
                         calibration_frame_offset=0, compression_scheme=0, calibration_file_type=0, compression_level=1,
                         l2_statistics=0, calibration_threshold_epsilon=EPS, frame_offset=0, num_threads=1,
                         rc_operation_mode=1, num_calibration_frames=1, reduction_level=2, keep_calibration_data=1,
                         source_bit_depth=BIT_DEPTH, target_bit_depth=BIT_DEPTH, keep_part_files=0, num_rows=NY,
                         num_cols=NX, source_data_type=0, target_data_type=0).items():
            ip._param_map[k] = v
        w = ReCoDeWriter('rb', dark_data=dark[None], output_directory=tmp, input_params=ip, mode='batch', node_id=0,
                         device=local_rank)
        w.start()
        w.run(np.stack([frames[i % len(frames)] for i in range(nz)]))
        w.close()
        path = os.path.join(tmp, 'rb.rc2_part000')
        fsize = os.path.getsize(path)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # one reader for the whole run, rewound every step: its engines (contexts, device buffers, pinned staging)
        # are set up once, as for a long acquisition decoded batch after batch
        r = ReCoDeReader(path, is_intermediate=True, device=local_rank)
        r.open(print_header=False)

        def one(what):
            r.rewind()
            if what == 'sum':
                ids, total = r.sum_frames(nz)
                rcd.allreduce_view(total)
                out = int(total[:1024].sum().item())          # device -> host read of the result
            else:
                ids, dense = r.read_frames_dense(min(nz, 128))
                out = int(dense[0, 0, :8].sum().item())
                del dense
            return len(ids), dict(r.bulk_stats)

        res = {}
        for what in ('sum', 'dense'):
            for _ in range(max(1, args.warmup)):
                one(what)
            barrier()
            t0 = time.perf_counter()
            n = 0
            for _ in range(args.steps):
                k, st = one(what)
                n += k
            barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[what] = (world * n / float(t[0]), float(t[0]) / args.steps, st)
        if rank == 0:
            line = {'metric': 'frames/s, 4096x4096 L2 part file -> live-view sum (NCCL all-reduce across ranks)',
                    'value': res['sum'][0], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
                    'warmup': max(1, args.warmup), 'ms_per_step': 1e3 * res['sum'][1], 'higher_is_better': True,
                    'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u16', 'data': 'synthetic',
                    'config': {'workload': 'read path (BASELINE config 5): %d-frame L2 part file per GPU on tmpfs, '
                                           '%d-bit, file reads inside the timed region' % (nz, BIT_DEPTH),
                               'file_bytes': fsize, 'frames_per_step_per_gpu': nz},
                    'dense_frames_per_s': res['dense'][0],
                    'dense_output_gb_s': res['dense'][0] * NY * NX * 2 / 1e9,
                    'host_time_split_ms_last_step': {k: (1e3 * v if k.endswith('_s') else v)
                                                     for k, v in res['sum'][2].items()},
                    'e2e': {'value': res['sum'][0], 'unit': 'frames/s', 'h2d_bytes_per_step': fsize,
                            'd2h_bytes_per_step': 8},
                    'roofline': None, 'cpu_baseline': None}
            print(json.dumps(line))
        r.close()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--level', type=int, default=2)
    ap.add_argument('--frames', type=int, default=32, help='frames per step per GPU')
    ap.add_argument('--distinct', type=int, default=4, help='distinct synthetic frames generated on the host')
    ap.add_argument('--cpu-frames-per-core', type=int, default=96,
                    help='frames per host core of the cpu_baseline sample (about 10 s of CPU work)')
    ap.add_argument('--ref-frames-per-core', type=int, default=4, help='frames per core per step of --impl reference')
    ap.add_argument('--slots', type=int, default=3, help='batches in flight (each on its own CUDA stream)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--bit-depth', type=int, default=12, help='source / target bit depth (SURVEY 8d: 8, 12, 16)')
    ap.add_argument('--mode', default='write', choices=['write', 'read'],
                    help="read: BASELINE config 5 -- an L2 part file per GPU -> live-view sum (+ NCCL all-reduce) and "
                         "dense frames through ReCoDeReader's bulk calls; an auxiliary line, not the headline")
    ap.add_argument('--read-frames', type=int, default=256, help='frames in the part file of --mode read')
    args = ap.parse_args()
    global BIT_DEPTH
    BIT_DEPTH = args.bit_depth

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    level = args.level
    frame_bytes = NY * NX * 2
    workload = ('L%d reduce + deflate, 4096x4096 uint16, %d-bit, zlib level 1, synthetic %s frames '
                '(SURVEY 8d), %d frames per step per GPU, frame-sharded' % (level, BIT_DEPTH, KIND[level], args.frames))
    config = {'workload': workload, 'reduction_level': level, 'frames_per_step_per_gpu': args.frames,
              'frame_shape': [NY, NX], 'bit_depth': BIT_DEPTH, 'compression_level': 1,
              'cache': 'inputs larger than L2 (%d MiB per step)' % (args.frames * frame_bytes >> 20),
              'batches_in_flight': args.slots,
              'clocks_sampled': 'timed region plus 1.5 s of the same steps, untimed'}
    metric = 'frames/s, 4096x4096 L%d reduce+deflate' % level

    # ----------------------------------------------------------------------------------- reference arm
    if args.impl == 'reference':
        if rank != 0:
            return
        dark, frames = make_inputs(level, args.distinct)
        best = None
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
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u16', 'data': 'synthetic',
                'config': config, 'input_gb_s': value * frame_bytes / 1e9,
                'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                 'sample': '%d frames per step (%d per core), oracle C port + stock zlib level 1; '
                                           'the reference cannot execute L2/L4 (SURVEY 0.1)' % (total, args.ref_frames_per_core)},
                'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
                'gpu_launches': 0}
        print(json.dumps(line))
        return

    # ----------------------------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from pyrecode_b200.engine import WriteEngine

    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if args.mode == 'read':
        run_read(args, rank, world, local_rank, dev)
        if world > 1:
            dist.destroy_process_group()
        return

    # every rank generates its own frame range of the stream (seed offset = rank: different frames per GPU)
    dark, frames = make_inputs(level, args.distinct, seed=1234 + rank)
    F = args.frames
    eng = WriteEngine(NY, NX, 2, BIT_DEPTH, level, 1, 0, 0, 1, max_frames=F, device=local_rank,
                      records_capacity=F * (frame_bytes // 4), n_slots=args.slots)
    eng.set_threshold(dark, EPS)
    host = torch.empty((F, NY, NX), dtype=torch.uint16).pin_memory()
    hv = host.numpy()
    for i in range(F):
        hv[i] = frames[i % len(frames)]
    d_frames = host.to(dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first_id = rank * F * (args.warmup + args.steps)

    # ---- device-resident timing (value) + dominant-kernel timing (roofline)
    # Steps are issued round-robin on the engine's slots (one CUDA stream each), the way the writer keeps
    # batches in flight; the timed region is bracketed by events on the default stream that all slot streams
    # fork from / join to.
    nsl = len(eng.slots)
    cur = torch.cuda.current_stream()

    def run_steps(n_steps, id0):
        for sl in eng.slots:
            sl.stream.wait_stream(cur)
        for s in range(n_steps):
            sl = eng.slots[s % nsl]
            with torch.cuda.stream(sl.stream):
                eng.launch(d_frames, F, id0 + s * F, s % nsl)
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
    run_steps(args.steps, first_id + args.warmup * F)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = sum(sl.ctx.launch_count() for sl in eng.slots) - launches0
    # the timed region lasts milliseconds, one nvidia-smi query ~0.1 s: keep the same steps running (untimed) for
    # about 1.5 s so that the clock sampler sees the GPU under this load
    if rank == 0:
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            run_steps(8, first_id)
            torch.cuda.synchronize()
    # stage split: re-run a few steps one at a time on slot 0 with per-step readback of the stage marks
    # (outside the timed region; the dominant kernel is timed alone here, which is what `roofline` reports)
    nprof = min(args.steps, 5)
    eng.ctx.set_pipelined(False)              # one batch at a time: every kernel gets the whole GPU
    for s in range(nprof):
        eng.launch(d_frames, F, 0)
        stage_ms += np.array(eng.ctx.profile_read()[:4])
    stage_ms /= nprof
    st = int(eng.status.cpu()[0])
    offs = eng.offsets.cpu().numpy()
    rec_bytes = int(offs[F])
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t[0])
    value = world * F * args.steps / (ms_total / 1e3)

    # ---- end to end through host buffers
    e2e = None
    if not args.no_e2e:
        # through the host-buffer API the writer uses: pinned host frames -> H2D -> kernels -> D2H of the records,
        # every step, with `slots` batches in flight
        def e2e_steps(n_steps, id0):
            h2d = d2h = 0
            pending = []
            for s in range(n_steps):
                pending.append(eng.submit(host, first_frame_id=id0 + s * F))
                if len(pending) == nsl:
                    _, _, _, h2d, d2h = eng.collect(pending.pop(0))
            while pending:
                _, _, _, h2d, d2h = eng.collect(pending.pop(0))
            return h2d, d2h

        e2e_steps(2, 0)
        barrier()
        e0.record()
        h2d, d2h = e2e_steps(args.steps, first_id)
        e1.record()
        barrier()
        ms_e2e = e0.elapsed_time(e1)
        t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {'value': world * F * args.steps / (float(t[0]) / 1e3), 'unit': 'frames/s',
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    k1_ms = float(stage_ms[0])
    achieved = F * frame_bytes / (k1_ms / 1e3) / 1e9
    line = {'metric': metric, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u16', 'data': 'synthetic', 'config': config,
            'input_gb_s': value * frame_bytes / 1e9,
            'hbm_roofline_frac_whole_path': value / world * frame_bytes / 1e9 / peak,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
            'roofline': {'bound': 'hbm', 'kernel': 'k_reduce_tiles_bulk', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': NCU_TRAFFIC.get((level, F)), 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': F * frame_bytes, 'kernel_ms': k1_ms},
            'stage_ms_per_step': dict(zip(['threshold_pack_compact', 'reduce_rest', 'deflate', 'assemble'],
                                          [float(x) for x in stage_ms])),
            'record_bytes_per_frame': rec_bytes / F, 'status': st}
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

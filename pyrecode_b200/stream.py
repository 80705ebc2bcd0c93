"""Streaming ingest: the acquisition drops `.seq` chunks into a watched (RAM-disk) directory, every
chunk is renamed to `Next_Stream.seq`, reduced and compressed by the writer(s) in `mode='stream'`, and deleted.

This is the data path of `ReCoDeServer._recode_queue_manager` + `ReCoDeNode._process_file`
(pyrecode/recode_server.py:463-564, 720-730) without its ZMQ control plane (out of scope): the "broadcast
process_file / wait for all nodes" step is a function call -- and, with one process per GPU, a barrier on either side.
The queueing rules are the reference's:

  * the directory is emptied when the session starts (:468-469);
  * a chunk is taken up only when a NEWER file is queued behind it, i.e. when the acquisition has finished writing it
    (:487); the last chunk is taken after waiting chunk_time_in_sec + 1 seconds (:540-541);
  * the first chunk that shows up is dropped from the queue unprocessed (:491-493; it stays in the directory);
  * max_count chunks are processed in all (:476, :536-563).

Chunks are read by pyrecode_b200.em_reader.SEQReader straight into the write engine's pinned staging buffers
(ReCoDeWriter.run with data=None and source_file_type = 2).
"""
import os
import time

NEXT_STREAM = 'Next_Stream.seq'


def _seq_files(source_dir):
    """`.seq` files of the directory, oldest first (modification time, then name: chunk names carry a counter)"""
    out = []
    for f in os.listdir(source_dir):
        p = os.path.join(source_dir, f)
        if f.endswith('.seq') and f != NEXT_STREAM and os.path.isfile(p):
            try:
                out.append((os.stat(p).st_mtime_ns, f))
            except FileNotFoundError:
                pass
    return [f for _, f in sorted(out)]


def recode_queue_manager(source_dir, max_count, chunk_time_in_sec, process_file, poll_s=0.02, clear=True,
                         should_stop=None, idle_timeout_s=None, log=None):
    """Watch `source_dir` and call process_file(path_of_Next_Stream.seq, chunk_name) for max_count chunks.
    Returns the list of chunk names processed, in order.  should_stop() -> True interrupts the wait between chunks;
    idle_timeout_s bounds the wait for a new file (None = wait for ever, as the reference does)."""
    if not os.path.isdir(source_dir):
        raise ValueError('Directory ' + source_dir + ' not found.')
    if clear:
        for f in os.listdir(source_dir):
            os.remove(os.path.join(source_dir, f))
    say = log or (lambda *a: None)
    queue, queued = [], set()
    is_first = True
    done = []
    target = os.path.join(source_dir, NEXT_STREAM)

    def scan():
        for f in _seq_files(source_dir):
            if f not in queued:
                queued.add(f)
                queue.append(f)

    def take(fname):
        os.rename(os.path.join(source_dir, fname), target)
        t0 = time.time()
        process_file(target, fname)
        say('Processed chunk %s in %.3f seconds.' % (fname, time.time() - t0))
        os.remove(target)
        done.append(fname)

    last_new = time.time()
    while len(done) < max_count - 1:
        if should_stop is not None and should_stop():
            return done
        n0 = len(queue)
        scan()
        if len(queue) != n0:
            last_new = time.time()
        if len(queue) > 1:
            fname = queue.pop(0)
            if is_first:
                is_first = False
                continue
            take(fname)
        else:
            if idle_timeout_s is not None and time.time() - last_new > idle_timeout_s:
                return done
            time.sleep(poll_s)
    # the last chunk: nothing will be queued behind it, so give the acquisition time to finish writing it
    t_end = None if idle_timeout_s is None else time.time() + idle_timeout_s
    while not queue:
        if (should_stop is not None and should_stop()) or (t_end is not None and time.time() > t_end):
            return done
        scan()
        if not queue:
            time.sleep(poll_s)
    fname = queue.pop(0)
    if is_first and max_count > 0:
        # a one-chunk session: the reference would still take the only file here (:536-545)
        is_first = False
    time.sleep(chunk_time_in_sec + 1 if chunk_time_in_sec >= 0 else 0)
    take(fname)
    return done


def run_stream(source_dir, writer, max_count, chunk_time_in_sec=0, rank=0, world=1, barrier=None, **kw):
    """One streaming session of a started ReCoDeWriter(mode='stream', image_filename=<source_dir>/Next_Stream.seq):
    every processed chunk is one writer.run().  With world > 1 (one process per GPU) rank 0 owns the directory --
    it renames and deletes -- and the ranks meet at `barrier()` before and after every chunk, which stands for the
    reference's 'process_file' broadcast and its wait for all nodes (:497-510).  -> list of run_metrics (this rank)."""
    metrics = []
    if world <= 1:
        recode_queue_manager(source_dir, max_count, chunk_time_in_sec,
                             lambda path, name: metrics.append(writer.run()), **kw)
        return metrics
    if barrier is None:
        from .distributed import barrier as _b
        barrier = _b
    import torch
    import torch.distributed as dist
    flag = torch.zeros(1, dtype=torch.int64)
    dev = None
    if dist.get_backend() == 'nccl':
        dev = torch.device('cuda', torch.cuda.current_device())
        flag = flag.to(dev)

    def announce(v):
        flag.fill_(v)
        dist.broadcast(flag, src=0)
        return int(flag.item())

    if rank == 0:
        def proc(path, name):
            announce(1)
            metrics.append(writer.run())
            barrier()                              # every rank has closed the chunk: it may be deleted
        recode_queue_manager(source_dir, max_count, chunk_time_in_sec, proc, **kw)
        announce(0)
    else:
        while announce(0) == 1:
            metrics.append(writer.run())
            barrier()
    return metrics

import sys, zlib
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from pyrecode_b200._native import Context
from pyrecode_b200 import engine
from pyrecode_b200.engine import deflate_batch, _stage_streams
from test_gpu_stages import payload_cases
ctx = Context()
cases = payload_cases()
names = list(cases)

def ru(x, m=256): return (x + m - 1)//m*m
def inflate_dbg(streams, cap, tag):
    n = len(streams)
    stride = (int(cap) + 15)//16*16 + 16
    d_in, d_off, d_sz, _ = _stage_streams(ctx, streams)
    out = ctx.empty(n*stride + 64); out_bytes = ctx.zeros(n, torch.int32); status = ctx.zeros(n, torch.int32)
    ws = ctx.inflate_zlib(d_in, d_off, d_sz, n, out, stride, out_bytes, status)
    torch.cuda.synchronize()
    st = status.cpu().numpy(); ob = out_bytes.cpu().numpy()
    bad = np.nonzero(st)[0]
    print(tag, 'n', n, 'bad', bad[:10], 'count', len(bad))
    if len(bad):
        w = ws.cpu().numpy()
        cmax = stride//16384 + 3
        off = 0
        cand = w[off:off+n*cmax*4].view(np.uint32).reshape(n, cmax); off = ru(off + n*cmax*4)
        ncand = w[off:off+(n+1)*4].view(np.uint32); off = ru(off + (n+1)*4)
        tb = w[off:off+(n+1)*4].view(np.uint32); off = ru(off + (n+1)*4)
        cnt = w[off:off+32].view(np.uint32); off = ru(off + 32)
        tasks = w[off:off+(n*cmax+1)*20].view(np.int32).reshape(-1, 5); off = ru(off + (n*cmax+1)*20)
        ns = w[off:off+(n+1)*4].view(np.uint32)
        for s in bad[:4]:
            print(' stream', s, 'clen', len(streams[s]), 'hex', streams[s][:16].hex(), 'ncand', ncand[s], 'cand', cand[s][:4], 'tb', tb[s], tb[s+1], 'need_serial', ns[s], 'out_bytes', ob[s])
            for j in range(min(int(ncand[s]), 4)):
                print('   task', j, tasks[tb[s]+j])
        print(' counters', cnt[:4], 'total tasks', tb[n])
    return st

def poison(val):
    ts = [torch.full((sz,), val, dtype=torch.uint8, device='cuda') for sz in (1<<20, 1<<22, 1<<23, 1<<24, 1<<26, 1<<10, 1<<14, 1<<16, 1<<18)]
    torch.cuda.synchronize(); del ts

for rep in range(3):
    poison([0xFF, 0x01, 0x80][rep])
    for level in (1, 0, 9):
        comp = deflate_batch(ctx, [cases[k] for k in names], level)
        for k, c in zip(names, comp):
            assert zlib.decompress(c) == cases[k]
    streams = []
    for k, d in cases.items():
        for lvl in (0, 1, 6, 9):
            streams.append(zlib.compress(d, lvl))
    inflate_dbg(streams, max(len(v) for v in cases.values()), 'stock rep%d' % rep)
    comp = deflate_batch(ctx, [cases[k] for k in names], 1)
    poison([0xFF, 0x01, 0x80][rep])
    inflate_dbg(comp, max(len(v) for v in cases.values()), 'own rep%d' % rep)

"""The reference arm of bench.py (`--impl reference`: the CPU restatement of the reference path on all host cores) needs
no GPU; this checks its JSON contract here: one line, the ours-arm's metric / unit / config, `impl`, `cpu_baseline`, and an
`e2e` that repeats the line's own value with no copies."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--ref-frames-per-core', '1'], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'].startswith('frames/s') and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['steps'] == 1 and d['warmup'] == 0 and d['n_gpus'] == 1
    assert d['value'] > 0 and d['config']['reduction_level'] == 2 and d['config']['frame_shape'] == [4096, 4096]
    cb = d['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert d.get('gpu_launches', 0) == 0

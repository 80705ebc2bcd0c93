"""Run ON THE GPU BOX: writes small ReCoDe files with the GPU writer from the committed golden inputs; the files are
brought back through gpurun_out/ours_golden/ and committed as tests/golden/ours_* so that the CPU-side test
tests/test_reference_reads_ours.py can open them with the UNMODIFIED reference reader (reverse interop, SURVEY
Appendix A).

    python tools/make_ours_golden.py [out_dir]
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main(out_dir):
    from test_gpu_api import make_params
    from pyrecode_b200.recode_reader import merge_parts
    from pyrecode_b200.recode_writer import ReCoDeWriter
    os.makedirs(out_dir, exist_ok=True)
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'gold_a_input.npz'))
    data, dark, eps = g['data'], g['dark'], int(g['eps'])
    nz, ny, nx = data.shape

    def write(name, level, nodes, mode=1, merged=False, clevel=1):
        ip = make_params(ny, nx, nz, level=level, mode=mode, b=12, eps=eps, threads=nodes, clevel=clevel)
        for node in range(nodes):
            w = ReCoDeWriter(name, dark_data=dark[None], output_directory=out_dir, input_params=ip, mode='batch',
                             node_id=node, batch_frames=3, merged=merged)
            w.start()
            w.run(data)
            w.close()
        if nodes > 1 or (not merged and name.endswith('_mg')):
            merge_parts(out_dir, '%s.rc%d' % (name, level), nodes)

    write('ours_a', 1, 3)                       # 3 part files + merge_parts -> ours_a.rc1
    write('ours_l3', 3, 1)                      # one part file, level 3
    write('ours_l3_mg', 3, 1)                   # ... and merged by merge_parts
    write('ours_m', 1, 1, merged=True)          # the merged layout written directly
    write('ours_m0', 1, 1, mode=0)              # reduce only (no deflate)
    write('ours_c9', 1, 1, clevel=9)            # per-stream Huffman codes (levels 6..9)
    print(sorted(os.listdir(out_dir)))


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'ours_golden'))

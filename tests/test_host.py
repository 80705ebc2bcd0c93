"""CPU: host-side drop-in layer (params, header, structures, merge) and the C ABI surface."""
import ctypes
import os
import re
import shutil
import zlib

import numpy as np
import pytest

from oracle import oracle as orc
from pyrecode_b200.misc import map_dtype, get_dtype_code, get_dtype_string
from pyrecode_b200.params import InitParams, InputParams
from pyrecode_b200.recode_header import ReCoDeHeader
from pyrecode_b200.recode_reader import ReCoDeReader, merge_parts
from pyrecode_b200.structures import ReCoDeStructures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PARAMS = dict(l4_centroiding=0, source_file_type=0, num_frames=8, source_header_length=0, calibration_frame_offset=0,
              compression_scheme=0, calibration_file_type=0, compression_level=1, l2_statistics=0,
              calibration_threshold_epsilon=6, frame_offset=0, num_threads=3, rc_operation_mode=1,
              num_calibration_frames=1, reduction_level=1, keep_calibration_data=1, source_bit_depth=12,
              target_bit_depth=12, keep_part_files=0, num_rows=37, num_cols=53, source_data_type=0, target_data_type=0)


def write_params(path, **kw):
    d = dict(PARAMS)
    d.update(kw)
    with open(path, 'w') as f:
        f.write('# comment line\n\n')
        for k, v in d.items():
            f.write('%s = %d\n' % (k, v))
    return d


def test_input_params_load_validate(tmp_path):
    p = str(tmp_path / 'p.txt')
    write_params(p)
    ip = InputParams()
    ip.load(p)
    assert ip.validate()
    assert (ip.nx, ip.ny, ip.nz) == (53, 37, 8) and ip.num_cols == 53
    assert ip.source_numpy_dtype == np.uint16 and ip.target_numpy_dtype == np.uint16
    assert ip.L2_statistics == ip.l2_statistics == 0
    ip.nx = 64
    assert ip.num_cols == 64
    ip.serialize(str(tmp_path / 'out.txt'))
    assert 'num_cols = 64' in open(str(tmp_path / 'out.txt')).read()
    with open(p, 'a') as f:
        f.write('bogus_key = 1\n')
    with pytest.raises(AssertionError):
        InputParams().load(p)


@pytest.mark.parametrize('key,val', [('reduction_level', 5), ('rc_operation_mode', 2), ('l2_statistics', 3),
                                     ('l4_centroiding', 4), ('compression_scheme', 12), ('compression_level', 23),
                                     ('source_data_type', 3), ('keep_part_files', 2)])
def test_input_params_rejects(tmp_path, key, val, capsys):
    p = str(tmp_path / 'p.txt')
    write_params(p, **{key: val})
    ip = InputParams()
    ip.load(p)
    assert not ip.validate()


def test_target_depth_defaults_to_source(tmp_path):
    p = str(tmp_path / 'p.txt')
    write_params(p, target_bit_depth=-1, source_bit_depth=8)
    ip = InputParams()
    ip.load(p)
    assert ip.validate() and ip.target_bit_depth == 8 and ip.source_numpy_dtype == np.uint8


def test_init_params():
    ip = InitParams('Batch ', '/tmp', image_filename='x', verbosity=7)
    assert ip.mode == 'batch' and ip.verbosity == 2
    with pytest.raises(ValueError):
        InitParams('batch', '', image_filename='x')
    with pytest.raises(ValueError):
        InitParams('online', '/tmp', image_filename='x')
    with pytest.raises(ValueError):
        InitParams('batch', '/tmp')


def test_map_dtype():
    assert map_dtype(0, 8) == np.uint8 and map_dtype(0, 9) == np.uint16 and map_dtype(0, 33) == np.uint64
    assert map_dtype(1, 16) == np.int16 and map_dtype(2, 32) == np.float32 and map_dtype(2, 33) == np.float64
    with pytest.raises(ValueError):
        map_dtype(0, 65)
    assert get_dtype_string(get_dtype_code(np.uint16)) == 'uint16'


def test_header_bytes_match_reference(gold_dir, tmp_path):
    """our serializer reproduces the 512 header bytes the reference wrote, and our parser reads them"""
    ref = open(os.path.join(gold_dir, 'gold_a.rc1_part001'), 'rb').read()[:512]
    p = str(tmp_path / 'p.txt')
    write_params(p)
    ip = InputParams()
    ip.load(p)
    assert ip.validate()
    init = InitParams('batch', str(tmp_path), image_filename='gold_a')
    h = ReCoDeHeader()
    h.create(init, ip, True)
    h.set('source_header_length', 0)
    h.update('nz', 3)                                  # part 001 holds 3 frames after close()
    assert h.validate() and h.recode_header_length == 512
    assert h.to_bytes() == ref
    h2 = ReCoDeHeader()
    h2.load(os.path.join(gold_dir, 'gold_a.rc1_part001'))
    d = h2.as_dict()
    assert d['nx'] == 53 and d['ny'] == 37 and d['nz'] == 3 and d['reduction_level'] == 1
    assert d['calibration_threshold_epsilon'] == 6 and d['source_file_name'].strip() == 'gold_a'
    assert h2.to_bytes() == ref
    assert orc.parse_header(ref)['nz'] == 3 and orc.build_header(orc.parse_header(ref)) == ref
    assert h2.get_field_position_in_bytes('nz') == 23 and h2.get_definition('nz')['bytes'] == 4
    assert h2.get_frame_data_offset(True, 12) == 512 and h2.get_frame_data_offset(False, 12) == 512 + 36


def test_structures():
    s = ReCoDeStructures({'nx': 53, 'ny': 37})
    assert s.binary_image_sz_bytes == 246
    assert s.get_standard_frame_metadata_size(1, 1) == 12 and s.get_standard_frame_metadata_size(3, 1) == 4
    assert s.get_standard_frame_metadata_size(2, 0) == 4 and s.get_standard_frame_metadata_size(4, 0) == 0
    md = {'bytes_in_compressed_binary_map': 10, 'bytes_in_compressed_pixvals': 20, 'bytes_in_packed_pixvals': 99}
    assert s.get_frame_data_size(1, 1, md) == 30
    assert s.get_frame_data_size(1, 0, {'bytes_in_packed_pixvals': 99}) == 246 + 99
    assert s.get_frame_data_size(4, 0, {}) == 246
    assert [f['name'] for f in s.standard_frame_metadata_structure_for(2, 1)] == \
        ['bytes_in_compressed_binary_map', 'bytes_in_compressed_summary_stats', 'bytes_in_packed_summary_stats']
    assert [f['name'] for f in s.standard_frame_metadata_structure_for(1, 1)] == orc.metadata_fields(1, 1)


def test_merge_parts_reproduces_reference_file(gold_dir, tmp_path):
    """merge_parts on the reference's part files == the reference's own merged file, byte for byte"""
    for node in range(3):
        shutil.copy(os.path.join(gold_dir, 'gold_a.rc1_part%03d' % node), str(tmp_path))
    merge_parts(str(tmp_path), 'gold_a.rc1', 3)
    assert open(str(tmp_path / 'gold_a.rc1'), 'rb').read() == open(os.path.join(gold_dir, 'gold_a.rc1'), 'rb').read()


def test_reader_metadata_and_raw_access(gold_dir):
    """header / seek table / raw record access need no GPU"""
    r = ReCoDeReader(os.path.join(gold_dir, 'gold_a.rc1'), is_intermediate=False)
    r.open(print_header=False)
    assert r.get_shape() == (8, 37, 53) and r.sz_frame_metadata == 12
    h, recs = orc.parse_merged_file(os.path.join(gold_dir, 'gold_a.rc1'))
    import zlib
    for z in range(8):
        f = r.get_next_frame_raw()
        assert list(f.keys()) == [z]
        assert zlib.decompress(f[z]['data']['binary_map']) == recs[z]['map']
        assert zlib.decompress(f[z]['data']['pixvals']) == recs[z]['vals']
        assert {k: int(v) for k, v in f[z]['metadata'].items()} == recs[z]['metadata']
    with pytest.raises(ValueError):
        r.get_frame(99)
    r.close()
    p = ReCoDeReader(os.path.join(gold_dir, 'gold_a.rc1_part002'), is_intermediate=True)
    p.open(print_header=False)
    ids = []
    while True:
        f = p.get_next_frame_raw(read_data=False)
        if f is None:
            break
        ids += list(f.keys())
    assert ids == [6, 7]
    with pytest.raises(ValueError):
        p.get_frame(0)
    p.close()


def test_c_abi_exports_every_declared_symbol():
    from pyrecode_b200 import _native
    hdr = open(os.path.join(ROOT, 'include', 'recode_b200.h')).read()
    declared = set(re.findall(r'\b(rc_[a-z0-9_]+)\s*\(', hdr))
    declared -= {'rc_ctx', 'rc_config'}
    assert declared, 'no declarations found'
    lib = ctypes.CDLL(_native.library_path())
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(_native.EXPORTS) == declared
    assert lib.rc_version() >= 100
    lib.rc_map_stride_words.restype = ctypes.c_size_t
    lib.rc_map_stride_words.argtypes = [ctypes.c_size_t]
    assert lib.rc_map_stride_words(4096 * 4096) == 4096 * 4096 // 32
    lib.rc_deflate_bound.restype = ctypes.c_size_t
    lib.rc_deflate_bound.argtypes = [ctypes.c_size_t]
    assert lib.rc_deflate_bound(0) == 8 and lib.rc_deflate_bound(16385) == 16385 + 20 + 8


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under pyrecode_b200/ may reference it"""
    pkg = os.path.join(ROOT, 'pyrecode_b200')
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh')):
                src = open(os.path.join(dp, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'liboracle' not in src, fn


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from pyrecode_b200._native import Context
    with pytest.raises(RuntimeError):
        Context()
    from pyrecode_b200.engine import WriteEngine
    with pytest.raises(RuntimeError):
        WriteEngine(64, 64, 2, 12, 1)


class _FakeEngine:
    """stands in for a ReadEngine in the host-side staging code: a plain buffer instead of pinned memory"""
    max_frames = 64

    def __init__(self):
        self.buf = np.zeros(1 << 16, dtype=np.uint8)

    def block_buffer(self, nbytes, keep=0):
        if self.buf.size < nbytes:
            new = np.zeros(max(nbytes, 2 * self.buf.size), dtype=np.uint8)
            new[:keep] = self.buf[:keep]
            self.buf = new
        return self.buf


@pytest.mark.parametrize('name,inter', [('gold_a.rc1_part001', True), ('gold_a.rc1', False), ('gold_c_l3m1.rc3_part000', True)])
def test_bulk_read_staging_offsets(gold_dir, name, inter):
    """ReCoDeReader._read_block (the host half of read_frames_dense / sum_frames): the records of a batch land in the
    staging block and the stream offsets / sizes it reports address exactly the reference-written streams"""
    path = os.path.join(gold_dir, name)
    if inter:
        _, recs = orc.parse_part_file(path)
    else:
        _, recs = orc.parse_merged_file(path)
    r = ReCoDeReader(path, is_intermediate=inter)
    r.open(print_header=False)
    eng = _FakeEngine()
    got = 0
    for n in (2, 64):                                    # two calls: the second continues where the first stopped
        ids, nbytes, moff, msz, voff, vsz = r._read_block(n, eng)
        for k, fid in enumerate(ids):
            want = recs[got + k]
            assert fid == want['frame_id']
            # (the oracle's merged-file walker keeps the inflated payloads only)
            assert zlib.decompress(bytes(eng.buf[moff[k]:moff[k] + msz[k]])) == want['map']
            if 'cmap' in want:
                assert bytes(eng.buf[moff[k]:moff[k] + msz[k]]) == want['cmap']
            if voff is not None:
                assert zlib.decompress(bytes(eng.buf[voff[k]:voff[k] + vsz[k]])) == want['vals']
            assert moff[k] + msz[k] <= nbytes
        got += len(ids)
    assert got == len(recs)
    assert r._read_block(4, eng)[0] == []                # end of file
    r.rewind()
    assert r._read_block(1, eng)[0] == [recs[0]['frame_id']]
    r.close()


def test_offline_utilities_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from scipy.sparse import coo_matrix
    from pyrecode_b200.utils.calibration import median_std
    from pyrecode_b200.utils.converters import recalibrate_l1
    with pytest.raises(RuntimeError):
        median_std(np.zeros((3, 8, 8), np.uint16))
    frames = {0: {'data': coo_matrix(np.ones((8, 8), np.uint16))}}
    with pytest.raises(RuntimeError):
        recalibrate_l1(frames, original_calibration_frame=np.zeros((8, 8), np.uint16),
                       new_calibration_frame=np.zeros((8, 8), np.uint16))

"""Per-frame metadata layout of ReCoDe files (drop-in for pyrecode/structures.py:18-91).

Every metadata field is a little-endian uint32.  `is_frame_size` marks the fields whose sum is the size of the
frame's data in the file.
"""
import math

import numpy as np


def _field(name, is_frame_size):
    return {'name': name, 'bytes': 4, 'dtype': np.uint32, 'is_frame_size': is_frame_size}


class ReCoDeStructures:

    def __init__(self, recode_header):
        self._recode_header = recode_header
        self._binary_image_sz_bytes = int(math.ceil(float(recode_header['nx']) * float(recode_header['ny']) / 8.))
        s = {}
        for level, stream in ((1, 'pixvals'), (2, 'summary_stats')):
            s[(level, 0)] = [_field('bytes_in_packed_' + stream, True)]
            s[(level, 1)] = [_field('bytes_in_compressed_binary_map', True),
                             _field('bytes_in_compressed_' + stream, True),
                             _field('bytes_in_packed_' + stream, False)]
        for level in (3, 4):
            s[(level, 0)] = []
            s[(level, 1)] = [_field('bytes_in_compressed_binary_map', True)]
        self._standard_frame_metadata_structure = s

    def get_standard_frame_metadata_size(self, reduction_level, rc_operation_mode):
        return sum(np.dtype(f['dtype']).itemsize
                   for f in self._standard_frame_metadata_structure[(reduction_level, rc_operation_mode)])

    def get_frame_data_size(self, reduction_level, rc_operation_mode, metadata):
        """bytes of one frame's data (without metadata) given its metadata dict (structures.py:60-91)."""
        fields = self._standard_frame_metadata_structure[(reduction_level, rc_operation_mode)]
        size = sum(int(metadata[f['name']]) for f in fields if f['is_frame_size'])
        if rc_operation_mode == 0:
            size += self._binary_image_sz_bytes
        return size

    @property
    def binary_image_sz_bytes(self):
        return self._binary_image_sz_bytes

    @property
    def standard_frame_metadata_structure(self):
        return self._standard_frame_metadata_structure

    def standard_frame_metadata_structure_for(self, reduction_level, rc_operation_mode):
        return self._standard_frame_metadata_structure[(reduction_level, rc_operation_mode)]

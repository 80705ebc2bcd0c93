"""Constants and dtype maps of the ReCoDe API (drop-in for pyrecode/misc.py:4-95)."""
import numpy as np


class rc_cfg:
    REQ_TYPE_QUERY = 0
    REQ_TYPE_COMMAND = 1

    FILE_TYPE_BINARY = 0
    FILE_TYPE_MRC = 1
    FILE_TYPE_SEQ = 2
    FILE_TYPE_OTHER = 255

    STATUS_CODE_BUSY = 0
    STATUS_CODE_AVAILABLE = 1
    STATUS_CODE_ERROR = -1
    STATUS_CODE_NOT_READY = -2
    STATUS_CODE_IS_CLOSED = -3
    STATUS_CODES = {'STATUS_CODE_BUSY': 0, 'STATUS_CODE_AVAILABLE': 1, 'STATUS_CODE_ERROR': -1,
                    'STATUS_CODE_NOT_READY': -2, 'STATUS_CODE_IS_CLOSED': -3}

    MESSAGE_TYPE_INFO = 0
    MESSAGE_TYPE_ERROR = -1
    MESSAGE_TYPE_STATUS = 1
    MESSAGE_TYPE_ACK = 2
    MESSAGE_TYPES = {'MESSAGE_TYPE_INFO': 0, 'MESSAGE_TYPE_ERROR': -1, 'MESSAGE_TYPE_STATUS': 1, 'MESSAGE_TYPE_ACK': 2}


_UNSIGNED = ((8, np.uint8), (16, np.uint16), (32, np.uint32), (64, np.uint64))
_SIGNED = ((8, np.int8), (16, np.int16), (32, np.int32), (64, np.int64))
_FLOAT = ((32, np.float32), (64, np.float64))


def map_dtype(type, bit_depth):
    """(type code 0 unsigned / 1 signed / 2 float, bit depth) -> smallest numpy dtype that holds it
    (pyrecode/misc.py:41-71)."""
    table = {0: _UNSIGNED, 1: _SIGNED, 2: _FLOAT}.get(type)
    if table is not None:
        for bits, dt in table:
            if bit_depth <= bits:
                return dt
    raise ValueError('Unable to match a numpy dtype for type = ' + str(type) +
                     ' (0=unsigned int, 1=signed int, 2=float) with bit depth = ' + str(bit_depth))


_DTYPE_CODES = [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64,
                np.float32, np.float64]


def get_dtype_code(dtype):
    for code, dt in enumerate(_DTYPE_CODES):
        if dtype == dt:
            return code
    raise ValueError('Unknown dtype')


def get_dtype_string(dtype):
    code = int(dtype)
    if 0 <= code < len(_DTYPE_CODES):
        return np.dtype(_DTYPE_CODES[code]).name
    raise ValueError('Unknown dtype')

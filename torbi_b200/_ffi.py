"""ctypes binding of libtorbi_b200.so (include/torbi_b200.h).

Replaces the `torch.ops.load_library(_C*.so)` discovery of
/root/reference/torbi/__init__.py:24-59.  There is exactly one library, built
in-tree by `make -C torbi_b200/csrc` (or __graft_entry__.build()); if it is
missing every decode fails loudly -- there is no Python or CPU fallback.
"""
import ctypes
import os
import pathlib

# TORBI_B200_LIBRARY: load another build of the same ABI (instrumented builds of tools/)
LIBRARY = pathlib.Path(os.environ.get('TORBI_B200_LIBRARY') or
                       pathlib.Path(__file__).resolve().parent / 'libtorbi_b200.so')

OK = 0
ERR_INVALID_ARGUMENT = -1
ERR_WORKSPACE = -2
ERR_CUDA = -3
ERR_OUT_OF_MEMORY = -4

OBS_IS_PROBABILITY = 1 << 0
EPS_FLOOR = 1 << 1
OBS_F16 = 1 << 4
OBS_BF16 = 1 << 5
FORCE_GENERIC = 1 << 8
FORCE_TILED = 1 << 9
TIME_KERNELS = 1 << 10
FORCE_WIDE = 1 << 11
TILE_ROWS_16 = 1 << 12
TILE_ROWS_32 = 2 << 12
TILE_ROWS_64 = 3 << 12
BANDED = 1 << 15
BACKPOINTERS = 1 << 16
FORCE_SPLIT = 1 << 17
SPLIT_STREAMED = 1 << 18
FORCE_CLUSTER = 1 << 19
FORCE_FLOW = 1 << 20

ABI_VERSION = 3

_p = ctypes.c_void_p
_i64 = ctypes.c_int64

SIGNATURES = {
    'torbi_b200_abi_version': (ctypes.c_int, []),
    'torbi_b200_last_error': (ctypes.c_char_p, []),
    'torbi_b200_device_count': (ctypes.c_int, []),
    'torbi_b200_workspace_bytes': (ctypes.c_size_t, [_i64, _i64, _i64, ctypes.c_uint32]),
    'torbi_b200_decode': (ctypes.c_int, [
        _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, _p, ctypes.c_size_t, _p]),
    'torbi_b200_workspace_bytes_planned': (ctypes.c_size_t, [_p, _i64, _i64, _i64, ctypes.c_uint32, ctypes.c_int]),
    'torbi_b200_decode_checkpointed': (ctypes.c_int, [
        _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, _p, ctypes.c_size_t, _p, _p, _i64]),
    'torbi_b200_decode_packed': (ctypes.c_int, [
        _p, _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, _p, ctypes.c_size_t, _p]),
    'torbi_b200_workspace_bytes_packed': (ctypes.c_size_t, [_p, _p, _i64, _i64, _i64, ctypes.c_uint32, ctypes.c_int]),
    'torbi_b200_decode_host_packed': (ctypes.c_int, [
        _p, _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, ctypes.c_int]),
    'torbi_b200_decode_host': (ctypes.c_int, [
        _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, ctypes.c_int]),
    'torbi_b200_decode_host_multi': (ctypes.c_int, [
        _p, _p, _p, _p, _i64, _i64, _i64, ctypes.c_uint32, _p, _p, _p, ctypes.c_int]),
    'torbi_b200_entropy': (ctypes.c_int, [_p, _i64, _i64, ctypes.c_uint32, _p, _p]),
    'torbi_b200_release': (None, []),
    'torbi_b200_launch_count': (ctypes.c_uint64, []),
    'torbi_b200_last_kernel': (ctypes.c_char_p, []),
    'torbi_b200_last_backtrace': (ctypes.c_char_p, []),
    'torbi_b200_last_timing': (ctypes.c_int, [_p, _p]),
    'torbi_b200_selftest_conversion': (ctypes.c_int, [_p]),
    'torbi_b200_plan': (ctypes.c_char_p, [_p, _i64, _i64, _i64, ctypes.c_uint32, ctypes.c_int]),
}



class PtInfo(ctypes.Structure):
    """torbi_b200_pt_info of include/torbi_b200.h"""
    _fields_ = [('data_offset', ctypes.c_int64), ('nbytes', ctypes.c_int64), ('shape', ctypes.c_int64 * 4),
                ('ndim', ctypes.c_int32), ('dtype', ctypes.c_int32), ('status', ctypes.c_int32),
                ('reserved', ctypes.c_int32)]


PT_OK = 0
PT_F32, PT_F16, PT_BF16, PT_F64, PT_I32, PT_I64 = range(6)

SIGNATURES.update({
    'torbi_b200_pt_inspect': (ctypes.c_int, [_p, _i64, _p, ctypes.c_int]),
    'torbi_b200_pt_read': (ctypes.c_int, [_p, _p, _i64, _p, _p, ctypes.c_int]),
    'torbi_b200_pt_write_i32': (ctypes.c_int, [_p, _p, _p, _p, _i64, ctypes.c_int]),
})

_lib = None


def library():
    """Load (once) and return the ctypes handle; raise if the build is absent."""
    global _lib
    if _lib is None:
        if not LIBRARY.exists():
            raise RuntimeError(
                f'{LIBRARY} not found: build it with `make -C torbi_b200/csrc` '
                '(or `python -c "import __graft_entry__ as g; g.build()"`). '
                'torbi_b200 has no CPU or pure-Python decode path.')
        lib = ctypes.CDLL(str(LIBRARY))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.torbi_b200_abi_version() != ABI_VERSION:
            raise RuntimeError('libtorbi_b200.so ABI version mismatch; rebuild')
        _lib = lib
    return _lib


def last_error() -> str:
    return library().torbi_b200_last_error().decode()


def check(status: int):
    """Map a C status to the exception the reference would raise: bad shapes /
    dtypes are ValueError, everything else RuntimeError (the reference raises
    RuntimeError from TORCH_INTERNAL_ASSERT, viterbi.cpp:187-190)."""
    if status == OK:
        return
    message = last_error()
    if status == ERR_INVALID_ARGUMENT:
        raise ValueError(f'torbi_b200: {message}')
    if status == ERR_OUT_OF_MEMORY:
        raise MemoryError(f'torbi_b200: {message}')
    raise RuntimeError(f'torbi_b200: {message} (status {status})')

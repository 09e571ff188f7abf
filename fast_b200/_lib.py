"""Loader of the CUDA C-ABI library (include/fastb200.h).  Fails loudly: there is no CPU fallback."""
import ctypes, os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FAST_B200_LIB: another build of the same library (kernel-configuration A/B runs, fast_b200/csrc/Makefile VAR=)
_PATH = os.environ.get('FAST_B200_LIB') or os.path.join(_HERE, 'libfastb200.so')
_lib = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)

# every symbol include/fastb200.h declares (checked by tests/test_host_logic.py::test_abi_exports_every_declared_symbol)
SYMBOLS = ['fb_version', 'fb_last_error', 'fb_device_count', 'fb_solver_create', 'fb_solver_destroy',
           'fb_zone_set_params', 'fb_zone_set_coords', 'fb_init_metric', 'fb_zone_get_metric',
           'fb_zone_upload_state', 'fb_zone_download_state', 'fb_zone_upload_aux', 'fb_zone_download_aux',
           'fb_zone_download_work', 'fb_zone_download_wig', 'fb_set_joins', 'fb_set_joins_tc', 'fb_compute', 'fb_apply_bc', 'fb_transfers', 'fb_compute_mut',
           'fb_zone_get_residuals', 'fb_zone_get_cfl', 'fb_zone_isfinite', 'fb_compute_host', 'fb_timing_enable', 'fb_timing_reset',
           'fb_timing_get', 'fb_sync', 'fb_zone_device_ptr', 'fb_halo_pack', 'fb_halo_unpack', 'fb_set_modulo_verif',
           'fb_stream', 'fb_halo_sizes', 'fb_set_exchange_cb', 'fb_halo_buffers', 'fb_counter', 'fb_exchange_stream',
           'fb_debug_tile_order', 'fb_nccl_unique_id', 'fb_nccl_init', 'fb_set_peers']


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError("fast_b200: %s is missing -- build it with `make -C fast_b200/csrc` "
                              "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % _PATH)
        L = ctypes.CDLL(_PATH)
        L.fb_last_error.restype = ctypes.c_char_p
        L.fb_zone_device_ptr.restype = ctypes.c_ulonglong
        L.fb_stream.restype = ctypes.c_ulonglong
        L.fb_exchange_stream.restype = ctypes.c_ulonglong
        L.fb_halo_pack.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_longlong]
        L.fb_halo_unpack.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_longlong]
        for n in SYMBOLS: getattr(L, n)         # AttributeError if the build lacks a declared entry point
        _lib = L
    return _lib


class FastError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise FastError("fast_b200 C-ABI error %d: %s" % (rc, lib().fb_last_error().decode()))

"""ctypes binding of libhdlz.so (include/hdlz.h).  No CPU fallback: a missing library,
a missing symbol or a missing GPU raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhdlz.so")

c_u8p = ctypes.c_void_p
c_u32p = ctypes.c_void_p
c_u64p = ctypes.c_void_p
u32, u64, cint, vp = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/hdlz.h declares
SIGNATURES = {
    "hdlz_version": (cint, []),
    "hdlz_last_error": (ctypes.c_char_p, []),
    "hdlz_status_name": (ctypes.c_char_p, [u32]),
    "hdlz_device_count": (cint, []),
    "hdlz_create": (cint, [cint, ctypes.POINTER(vp)]),
    "hdlz_destroy": (cint, [vp]),
    "hdlz_compress_bound": (u32, [u32]),
    "hdlz_set_match10": (cint, [vp, cint]),
    "hdlz_get_match10": (cint, [vp]),
    "hdlz_set_fast": (cint, [vp, cint]),
    "hdlz_get_fast": (cint, [vp]),
    "hdlz_set_container": (cint, [vp, cint]),
    "hdlz_get_container": (cint, [vp]),
    "hdlz_compress_bound_ex": (u32, [u32, cint]),
    "hdlz_set_tree": (cint, [vp, c_u8p, c_u8p]),
    "hdlz_get_tree": (cint, [vp, c_u8p, c_u8p]),
    "hdlz_train_tree": (cint, [vp, c_u8p, u64, c_u32p, u32, u64, vp]),
    "hdlz_compress_bound_tree": (u32, [vp, u32]),
    "hdlz_compress_stream_dyn": (cint, [vp, c_u8p, u32, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_tree_lengths": (cint, [c_u64p, cint, cint, c_u8p]),
    "hdlz_tree_header": (cint, [c_u8p, c_u8p, cint, c_u8p, u32, ctypes.POINTER(u32)]),
    "hdlz_compress_batch": (cint, [vp, c_u8p, u64, c_u32p, u32, c_u8p, u64, c_u32p, c_u32p, u64, vp]),
    "hdlz_decompress_batch": (cint, [vp, c_u8p, c_u64p, u64, c_u32p, c_u8p, u64, u32, c_u32p, c_u32p, u64, u32, vp]),
    "hdlz_compress_host": (cint, [vp, c_u8p, u64, c_u32p, u32, c_u8p, u64, c_u32p, c_u32p, u64]),
    "hdlz_decompress_host": (cint, [vp, c_u8p, c_u64p, u64, c_u32p, c_u8p, u64, u32, c_u32p, c_u32p, u64, u32]),
    "hdlz_pack_batch": (cint, [vp, c_u8p, u64, c_u32p, c_u8p, c_u64p, c_u64p, u64, vp]),
    "hdlz_compress_host_packed": (cint, [vp, c_u8p, u64, c_u32p, u32, c_u8p, u64, c_u64p, c_u32p, c_u32p, u64,
                                         ctypes.POINTER(u64)]),
    "hdlz_compress_stream": (cint, [vp, c_u8p, u32, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_decompress_stream": (cint, [vp, c_u8p, u32, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), u32]),
    "hdlz_cstream_begin": (cint, [vp, ctypes.POINTER(vp)]),
    "hdlz_cstream_feed": (cint, [vp, c_u8p, u32, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_cstream_finish": (cint, [vp, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_cstream_end": (cint, [vp]),
    "hdlz_dstream_begin": (cint, [vp, u32, u32, ctypes.POINTER(vp)]),
    "hdlz_dstream_feed": (cint, [vp, c_u8p, u32, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_dstream_finish": (cint, [vp, c_u8p, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u32)]),
    "hdlz_dstream_end": (cint, [vp]),
    "hdlz_dev_alloc": (cint, [vp, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "hdlz_dev_free": (cint, [vp, vp]),
    "hdlz_host_alloc_pinned": (cint, [vp, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "hdlz_host_free_pinned": (cint, [vp, vp]),
    "hdlz_copy_h2d": (cint, [vp, vp, vp, ctypes.c_size_t, vp]),
    "hdlz_copy_d2h": (cint, [vp, vp, vp, ctypes.c_size_t, vp]),
    "hdlz_stream_sync": (cint, [vp, vp]),
    "hdlz_generate_blocks": (cint, [vp, c_u8p, u64, u32, u64, u64, u64, vp]),
    "hdlz_launch_count": (u64, [vp]),
}

_lib = None


def load():
    """Load libhdlz.so and type every entry point.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libhdlz.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C hdl-deflate_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

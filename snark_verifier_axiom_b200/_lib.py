"""ctypes binding of libsvk.so (include/svk.h).  Loading fails loudly when the library is missing;
`Context()` fails loudly when no sm_100 GPU is present -- there is no CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvk.so")


class SvkError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise SvkError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (make -C csrc)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, sz, i32, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p
    lib.svk_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.svk_destroy.argtypes = [vp]
    lib.svk_destroy.restype = None
    lib.svk_last_error.argtypes = [vp]
    lib.svk_last_error.restype = ctypes.c_char_p
    lib.svk_set_stream.argtypes = [vp, vp]
    lib.svk_sync.argtypes = [vp]
    lib.svk_launch_count.argtypes = [vp]
    lib.svk_launch_count.restype = ctypes.c_uint64
    lib.svk_profile_enable.argtypes = [vp, i32]
    lib.svk_profile_report.argtypes = [vp, ctypes.c_char_p, sz]
    lib.svk_profile_timeline.argtypes = [vp, ctypes.c_char_p, sz]
    lib.svk_dk_load.argtypes = [vp, u8p]
    lib.svk_kzg_decide_batch.argtypes = [vp, i32, sz, vp, vp]
    lib.svk_kzg_decide_batch_dev.argtypes = [vp, i32, sz, vp, vp]
    u32p = ctypes.POINTER(ctypes.c_uint32)
    lib.svk_protocol_compile.argtypes = [vp, u8p, sz, i32, i32]
    lib.svk_protocol_compile_ex.argtypes = [vp, u8p, sz, i32, i32, i32]
    lib.svk_protocol_compile_bincode.argtypes = [vp, u8p, sz, i32, i32, i32, i32, ctypes.POINTER(sz), ctypes.POINTER(i32)]
    lib.svk_protocol_info.argtypes = [vp, i32, u32p]
    lib.svk_plonk_instance_shape_ok.argtypes = [vp, i32, ctypes.c_uint32, vp]
    lib.svk_plonk_succinct_verify_batch.argtypes = [vp, i32, sz, vp, ctypes.c_uint32, vp, sz, vp, vp, vp, vp]
    lib.svk_plonk_succinct_verify_batch_dev.argtypes = [vp, i32, sz, vp, ctypes.c_uint32, vp, sz, vp, vp, vp, vp]
    lib.svk_kzg_as_fold.argtypes = [vp, sz, vp, sz, vp, vp, vp]
    lib.svk_kzg_as_fold_zk.argtypes = [vp, sz, vp, vp, sz, vp, vp, vp]
    lib.svk_kzg_as_fold_dev.argtypes = [vp, sz, vp, sz, vp, vp, vp]
    lib.svk_plonk_verify_batch.argtypes = [vp, i32, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, i32, vp, vp, vp]
    lib.svk_plonk_verify_batch_dev.argtypes = [vp, i32, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, vp, vp, vp]
    lib.svk_msm_g1.argtypes = [vp, sz, vp, vp, vp, vp]
    lib.svk_msm_g1_dev.argtypes = [vp, sz, vp, vp, vp, vp]
    lib.svk_g1_mul_batch.argtypes = [vp, sz, vp, vp, sz, vp]
    lib.svk_g1_mul_batch_dev.argtypes = [vp, sz, vp, vp, sz, vp]
    lib.svk_plonk_verify_multi.argtypes = [vp, i32, sz, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, i32, vp, vp]
    lib.svk_plonk_verify_multi_dev.argtypes = [vp, i32, sz, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, vp, vp, vp]
    lib.svk_plonk_fold_multi_dev.argtypes = [vp, i32, sz, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, vp, vp, vp]
    lib.svk_kzg_as_fold_multi_dev.argtypes = [vp, sz, sz, vp, sz, vp]
    lib.svk_kzg_decide_records_dev.argtypes = [vp, i32, sz, vp]
    lib.svk_protocol_msm_terms.argtypes = [vp, i32, i32, vp, sz]
    lib.svk_plonk_msm_scalars_batch.argtypes = [vp, i32, sz, vp, ctypes.c_uint32, vp, sz, vp, vp, vp, vp]
    lib.svk_msm_curve.argtypes = [vp, i32, sz, vp, vp, vp, vp]
    lib.svk_msm_curve_dev.argtypes = [vp, i32, sz, vp, vp, vp, vp]
    lib.svk_ipa_decide_batch.argtypes = [vp, i32, ctypes.c_uint32, vp, sz, vp, vp, vp, vp]
    lib.svk_ipa_decide_batch_dev.argtypes = [vp, i32, ctypes.c_uint32, vp, sz, vp, vp, vp, vp]
    lib.svk_poseidon_squeeze.argtypes = [vp, sz, vp, ctypes.c_uint32, i32, vp]
    lib.svk_nccl_unique_id.argtypes = [vp]
    lib.svk_nccl_init.argtypes = [vp, i32, i32, vp]
    lib.svk_nccl_attach.argtypes = [vp, vp, i32, i32]
    lib.svk_plonk_verify_sharded_dev.argtypes = [vp, i32, sz, sz, vp, ctypes.c_uint32, vp, sz, vp, sz, vp, vp, vp, vp, vp]
    lib.svk_bench_modmul_peak.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _load()
    return _LIB

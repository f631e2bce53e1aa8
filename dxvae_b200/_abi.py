"""ctypes signatures of include/dxvae_b200.h (one table, shared by the product
loader in _lib.py and by the CPU-emulation harness under tests/emu)."""
import ctypes as C

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F = C.c_float
SZ = C.c_size_t


class ParamEntry(C.Structure):
    _fields_ = [("name", C.c_char_p), ("offset", I64), ("rows", I32), ("cols", I32)]


SIGNATURES = {
    "dxvae_abi_version": (C.c_int, []),
    "dxvae_last_error": (C.c_char_p, []),
    "dxvae_launch_count": (C.c_longlong, []),
    "dxvae_prof_begin": (None, [C.c_int]),
    "dxvae_prof_end": (None, [P, P, P]),
    "dxvae_param_blob_floats": (I64, []),
    "dxvae_param_count": (I64, []),
    "dxvae_param_entry": (C.c_int, [C.c_int, C.POINTER(ParamEntry)]),
    "dxvae_batch_build_host": (C.c_int, [I64, P, P, P, P, P, P, P, P, P, P, C.POINTER(I32)]),
    "dxvae_batch_schedule": (C.c_int, [I64, P, P, P, P, P, P, SZ, P]),
    "dxvae_batch_steps": (C.c_int, [I64, P, P, P, P, P, SZ, P]),
    "dxvae_batch_steps_host": (C.c_int, [I64, P, P, P]),
    "dxvae_pack_graphs": (C.c_int, [I64, P, P, P, P, P]),
    "dxvae_pack_graphs_indexed": (C.c_int, [I64, P, P, P, P, P, P, P, P]),
    "dxvae_unpack_graphs": (C.c_int, [I64, P, P, P, P, P]),
    "dxvae_voices_to_graphs": (C.c_int, [I64, P, P, P, P, P, P, P]),
    "dxvae_pack_syx": (C.c_int, [I64, P, P, P]),
    "dxvae_workspace_bytes": (SZ, [C.c_int, I64]),
    "dxvae_workspace_bytes_sched": (SZ, [C.c_int, I64, I32, P, P]),
    "dxvae_encode_fwd": (C.c_int, [P, I64, P, P, I32, P, P, P, P, P, P, SZ, C.c_int, C.c_int, P]),
    "dxvae_reparameterize": (C.c_int, [I64, P, P, P, P, P]),
    "dxvae_decode_greedy": (C.c_int, [P, I64, P, P, P, P, P, P, SZ, C.c_int, P]),
    "dxvae_elbo_step": (C.c_int, [P, I64, P, P, P, I32, P, P, P, P, F, F, F, F, P, P, P, P, P, SZ, C.c_int, P, P, P, P]),
    "dxvae_loss_step": (C.c_int, [P, I64, P, P, P, P, P, P, F, F, F, F, P, P, P, P, P, SZ, C.c_int, P, P, P]),
    "dxvae_encode_bwd": (C.c_int, [P, I64, P, P, I32, P, P, P, P, P, P, P, P, SZ, C.c_int, P]),
    "dxvae_adamw_step": (C.c_int, [I64, P, P, P, P, F, F, F, F, F, I64, F, P]),
    "dxvae_test_gemm": (C.c_int, [C.c_int, I64, I64, I64, P, I64, P, I64, P, I64, P, C.c_int, C.c_int, P]),
}

PREC_FP32, PREC_TF32, PREC_3XTF32 = 0, 1, 2
OP_ENCODE, OP_DECODE, OP_TRAIN, OP_SCHEDULE, OP_ENCODE_TRAIN, OP_LOSS = 0, 1, 2, 3, 4, 5


def bind(lib):
    """Attach restype/argtypes; raises AttributeError if a declared symbol is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class DxError(RuntimeError):
    pass


def check(lib, rc, what):
    if rc != 0:
        raise DxError("%s failed: %s" % (what, lib.dxvae_last_error().decode()))

"""ctypes binding of libcpf.so (include/cpf.h).  There is no fallback: if the CUDA library is missing
or no device is present, every entry point of the package raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPF_LIB", os.path.join(_HERE, "libcpf.so"))  # CPF_LIB: experiment builds only


class CpfConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int), ("interp", C.c_int), ("locator", C.c_int), ("integrator", C.c_int), ("rng", C.c_int),
        ("reflect_wall", C.c_int), ("path", C.c_int), ("sort_interval", C.c_int), ("fuse_substeps", C.c_int),
        ("dt", C.c_double), ("diffusion_coeff", C.c_double), ("seed", C.c_ulonglong), ("save_interval", C.c_int),
        ("reserved", C.c_int * 7),
    ]


class CpfStats(C.Structure):
    _fields_ = [
        ("n_particles", C.c_longlong), ("n_active", C.c_longlong), ("n_negative_tet", C.c_longlong),
        ("n_escaped", C.c_longlong), ("n_reflections", C.c_longlong), ("n_exact", C.c_longlong),
        ("n_hops", C.c_longlong), ("n_substeps", C.c_longlong), ("kinetic_energy", C.c_double),
        ("reserved", C.c_double * 3),
    ]


# every symbol include/cpf.h declares: (name, restype, argtypes)
_vp, _ip, _dp, _ll = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_longlong
SYMBOLS = {
    "cpf_abi_version": (C.c_int, []),
    "cpf_device_count": (C.c_int, [_ip]),
    "cpf_default_config": (None, [C.POINTER(CpfConfig)]),
    "cpf_create": (C.c_int, [C.POINTER(CpfConfig), C.POINTER(_vp)]),
    "cpf_destroy": (C.c_int, [_vp]),
    "cpf_last_error": (C.c_char_p, [_vp]),
    "cpf_sync": (C.c_int, [_vp]),
    "cpf_set_stream": (C.c_int, [_vp, _vp]),
    "cpf_profile_enable": (C.c_int, [_vp, C.c_int]),
    "cpf_profile_read": (C.c_int, [_vp, _ip, _dp, _dp]),
    "cpf_set_config": (C.c_int, [_vp, C.POINTER(CpfConfig)]),
    "cpf_mesh_upload_poly": (C.c_int, [_vp, C.c_int, _dp, C.c_int, _ip, _ip, _ip, C.c_int, _ip, C.c_int, _dp, _ip, C.c_int, _ip, _ip]),
    "cpf_set_patch_restitution": (C.c_int, [_vp, C.c_int, _dp]),
    "cpf_mesh_upload_tets": (C.c_int, [_vp, C.c_int, _dp, _ll, _ip, _ip, C.c_int]),
    "cpf_mesh_info": (C.c_int, [_vp, C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_ll), C.POINTER(_ll)]),
    "cpf_mesh_download_tets": (C.c_int, [_vp, _ip, _ip]),
    "cpf_mesh_download_neighbours": (C.c_int, [_vp, _ip]),
    "cpf_update_velocity": (C.c_int, [_vp, _vp, C.c_int]),
    "cpf_update_vertex_velocity": (C.c_int, [_vp, _vp, C.c_int]),
    "cpf_seed_box": (C.c_int, [_vp, _ll, _dp, _dp, C.c_ulonglong]),
    "cpf_seed_box_slice": (C.c_int, [_vp, _ll, _ll, _dp, _dp, C.c_ulonglong]),
    "cpf_set_particles": (C.c_int, [_vp, _ll, _dp]),
    "cpf_set_tets": (C.c_int, [_vp, _ip]),
    "cpf_locate_initial": (C.c_int, [_vp]),
    "cpf_init_rng": (C.c_int, [_vp]),
    "cpf_reseed_inactive": (C.c_int, [_vp, _dp, _dp, C.c_ulonglong, C.POINTER(_ll)]),
    "cpf_relocate_lost": (C.c_int, [_vp]),
    "cpf_advect": (C.c_int, [_vp, C.c_double, _ip]),
    "cpf_substeps": (C.c_int, [_vp, C.c_int, C.c_double]),
    "cpf_initial_advect": (C.c_int, [_vp]),
    "cpf_sort_particles": (C.c_int, [_vp]),
    "cpf_last_step_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "cpf_download": (C.c_int, [_vp, _dp, _dp, _ip]),
    "cpf_download_cells": (C.c_int, [_vp, _ip]),
    "cpf_stats_get": (C.c_int, [_vp, C.POINTER(CpfStats)]),
    "cpf_stats_request": (C.c_int, [_vp, C.c_int]),
    "cpf_stats_collect": (C.c_int, [_vp, C.POINTER(CpfStats)]),
    "cpf_comm_unique_id": (C.c_int, [_vp, C.c_size_t]),
    "cpf_comm_init": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, C.c_int]),
    "cpf_comm_info": (C.c_int, [_vp, _ip, _ip, _ip]),
    "cpf_update_velocity_bcast": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "cpf_update_velocity_slices": (C.c_int, [_vp, _ll, _ll, _vp, C.c_int]),
    "cpf_set_particle_id_base": (C.c_int, [_vp, _ll]),
    "cpf_write_vtu": (C.c_int, [_vp, C.c_char_p, C.c_uint]),
    "cpf_write_vtu_async": (C.c_int, [_vp, C.c_char_p, C.c_uint, C.c_int]),
    "cpf_output_wait": (C.c_int, [_vp]),
    "cpf_checkpoint_save": (C.c_int, [_vp, C.c_char_p]),
    "cpf_checkpoint_load": (C.c_int, [_vp, C.c_char_p]),
    "cpf_step_index": (C.c_ulonglong, [_vp]),
    "cpf_num_particles": (_ll, [_vp]),
    "cpf_device_pointers": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "cpf_debug_next_normals": (C.c_int, [_vp, _dp]),
    "cpf_debug_normals": (C.c_int, [_vp, C.c_int, _dp]),
    "cpf_launch_count": (_ll, [_vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libcpf.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"], stdout=out)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  cudaparticlesfoam_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib

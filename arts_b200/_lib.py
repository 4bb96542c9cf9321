"""Loader + ctypes prototypes of the in-tree CUDA library ``arts_b200/libarts_b200.so``.

Fails loudly: a missing library raises ``ImportError`` at first use, a failing call
raises ``Ab200Error`` with the library's message.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _abi as abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, os.environ.get("AB200_LIB", "libarts_b200.so"))  # AB200_LIB: experimental builds only
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "arts_b200.h")

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class Ab200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"arts_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C arts_b200/csrc`). arts_b200 has no CPU fallback."
        )
    L = C.CDLL(SO_PATH)
    L.ab200_last_error.restype = C.c_char_p
    L.ab200_device_count.restype = C.c_int
    L.ab200_set_device.argtypes = [C.c_int]
    L.ab200_set_thread_stream.argtypes = [_vp]
    L.ab200_path_set_timing.argtypes = [_vp, C.c_int]
    L.ab200_path_get_timings.argtypes = [_vp, _dp, C.POINTER(C.c_int64)]
    L.ab200_path_region_histogram.argtypes = [_vp, C.c_int64, C.c_uint64, _dp]
    L.ab200_launch_count.restype = C.c_int64
    L.ab200_launch_count.argtypes = [C.c_int]
    L.ab200_catalog_create.argtypes = [C.POINTER(abi.CatalogDesc), C.POINTER(_vp)]
    L.ab200_catalog_destroy.argtypes = [_vp]
    L.ab200_catalog_destroy.restype = None
    L.ab200_catalog_counts.argtypes = [_vp, C.POINTER(C.c_int64)]
    L.ab200_propmat_levels.argtypes = [_vp] + abi.SIG_PROPMAT_LEVELS_CORE + [C.c_uint32, _dp, _dp]
    L.ab200_tramat.argtypes = abi.SIG_TRAMAT
    L.ab200_srcvec.argtypes = abi.SIG_SRCVEC
    L.ab200_rte_emission.argtypes = abi.SIG_RTE
    L.ab200_rte_transmission.argtypes = abi.SIG_TRANSMISSION
    L.ab200_clearsky_emission.argtypes = [_vp] + abi.SIG_CLEARSKY_CORE
    L.ab200_planck_tb.argtypes = [C.c_int64, _dp, _dp]
    L.ab200_set_thread_grid_bounds.argtypes = [C.c_int32, _dp]
    L.ab200_atm_path_from_profile.argtypes = [C.POINTER(abi.AtmProfileDesc), C.c_int32, C.c_int32, C.c_int32, _dp, C.POINTER(C.c_uint8)] + [_dp] * 8
    L.ab200_multi_create.argtypes = [C.POINTER(abi.CatalogDesc), C.c_int32, C.POINTER(C.c_int32), C.POINTER(_vp)]
    L.ab200_multi_destroy.argtypes = [_vp]
    L.ab200_multi_destroy.restype = None
    L.ab200_multi_device_count.argtypes = [_vp]
    L.ab200_multi_device_count.restype = C.c_int32
    L.ab200_multi_clearsky_emission.argtypes = [_vp] + abi.SIG_CLEARSKY_CORE
    L.ab200_multi_propmat_levels.argtypes = [_vp] + abi.SIG_PROPMAT_LEVELS_CORE + [C.c_uint32, _dp, _dp]
    L.ab200_path_create.argtypes = [_vp, C.c_int64, C.c_int32, C.c_int32, C.POINTER(_vp)]
    L.ab200_path_create_stage2.argtypes = [_vp, C.c_int64, C.c_int32, C.POINTER(_vp)]
    L.ab200_path_adopt_K.argtypes = [_vp]
    L.ab200_path_destroy.argtypes = [_vp]
    L.ab200_path_destroy.restype = None
    L.ab200_path_set_stream.argtypes = [_vp, _vp]
    L.ab200_path_upload.argtypes = [
        _vp, _dp, C.c_int64, C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.POINTER(abi.Target), _dp, C.c_int32,
        C.c_int32, _dp, C.c_uint32,
    ]
    L.ab200_path_set_grid_bounds.argtypes = [_vp, _dp]
    L.ab200_path_run_propmat.argtypes = [_vp]
    L.ab200_path_run_stokes.argtypes = [_vp]
    L.ab200_path_download.argtypes = [_vp, _dp, _dp, _dp, _dp]
    L.ab200_path_sync.argtypes = [_vp]
    L.ab200_predef_levels.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.POINTER(abi.PredefSpecies), C.c_int64, _dp, C.c_int64,
                                      C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, C.POINTER(abi.Target), _dp, _dp, _dp]
    L.ab200_path_add_predefined.argtypes = [_vp, C.POINTER(C.c_int32), C.c_int32, C.POINTER(abi.PredefSpecies), _dp]
    L.ab200_predef_data_create.argtypes = [C.POINTER(abi.MtckdWater), C.POINTER(abi.MtckdWater), C.c_int32, C.POINTER(_vp)]
    L.ab200_predef_data_destroy.argtypes = [_vp]
    L.ab200_predef_data_destroy.restype = None
    L.ab200_predef_levels_data.argtypes = L.ab200_predef_levels.argtypes + [_vp]
    L.ab200_path_add_predefined_data.argtypes = L.ab200_path_add_predefined.argtypes + [_vp]
    L.ab200_lookup_create.argtypes = [C.POINTER(abi.LookupTableDesc), C.c_int32, C.POINTER(_vp)]
    L.ab200_lookup_destroy.argtypes = [_vp]
    L.ab200_lookup_destroy.restype = None
    L.ab200_lookup_levels.argtypes = [_vp, C.c_int64, _dp, C.c_int64, C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.POINTER(abi.Target), _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp, _dp]
    L.ab200_lookup_precompute.argtypes = [_vp, C.c_int64, _dp, C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.c_int32, _dp, C.c_int32,
                                          _dp, C.POINTER(abi.PartfunTable), _dp]
    L.ab200_path_add_lookup.argtypes = [_vp, _vp, C.c_int32, _dp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32]
    L.ab200_partfun_eval.argtypes = [C.POINTER(abi.PartfunTable), C.c_int32, C.c_int32, _dp, _dp, _dp]
    L.ab200_cia_create.argtypes = [C.POINTER(abi.CiaRecordDesc), C.c_int32, C.POINTER(_vp)]
    L.ab200_cia_destroy.argtypes = [_vp]
    L.ab200_cia_destroy.restype = None
    L.ab200_path_add_cia.argtypes = [_vp, _vp, C.c_double, C.c_int32, C.c_double]
    L.ab200_cia_levels.argtypes = [_vp, C.c_int64, _dp, C.c_int64, C.POINTER(abi.AtmPathDesc), C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(abi.Target), C.c_double, C.c_double, C.c_int32, _dp, _dp]
    L.ab200_hitran_read_par.argtypes = [C.c_char_p, C.c_int64, C.c_double, C.c_double, C.c_int32, C.POINTER(abi.HitranIsotopologue),
                                        C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]
    L.ab200_hitran_read_par_file.argtypes = [C.c_char_p, C.c_double, C.c_double, C.c_int32, C.POINTER(abi.HitranIsotopologue),
                                             C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]
    L.ab200_xml_read_bands.argtypes = [C.c_char_p, C.c_int64, C.POINTER(abi.XmlIsotopologue), C.c_int32, C.POINTER(abi.XmlSpecies),
                                       C.c_int32, C.c_int32, C.POINTER(_vp)]
    L.ab200_xml_read_bands_file.argtypes = [C.c_char_p, C.POINTER(abi.XmlIsotopologue), C.c_int32, C.POINTER(abi.XmlSpecies),
                                            C.c_int32, C.c_int32, C.POINTER(_vp)]
    L.ab200_xml_desc.argtypes = [_vp]
    L.ab200_xml_desc.restype = C.POINTER(abi.CatalogDesc)
    L.ab200_xml_destroy.argtypes = [_vp]
    L.ab200_xml_destroy.restype = None
    L.ab200_hitran_desc.argtypes = [_vp]
    L.ab200_hitran_desc.restype = C.POINTER(abi.CatalogDesc)
    L.ab200_hitran_destroy.argtypes = [_vp]
    L.ab200_hitran_destroy.restype = None
    L.ab200_path_run_observer.argtypes = [_vp, C.POINTER(abi.ObserverDesc)]
    L.ab200_path_download_observer.argtypes = [_vp, _dp, _dp, _dp, _dp]
    L.ab200_path_device_ptr.argtypes = [_vp, C.c_int]
    L.ab200_path_device_ptr.restype = _vp
    L.ab200_release_thread_cache.argtypes = []
    L.ab200_measure_dfma_peak.argtypes = [C.c_int, _dp, _dp]
    L.ab200_measure_dfma_mix.argtypes = [C.c_int, _dp, _dp]
    L.ab200_faddeeva_w.argtypes = [C.c_int64, _dp, _dp, _dp, _dp]
    L.ab200_dawson.argtypes = [C.c_int64, _dp, _dp, _dp, _dp]
    L.ab200_zeeman_components.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp]
    L.ab200_norm_view.argtypes = [C.c_int, _dp, _dp, _dp]
    _lib = L
    return L


def check(rc: int):
    if rc:
        raise Ab200Error(rc, lib().ab200_last_error().decode(errors="replace"))


def declared_symbols() -> list[str]:
    """Every function name declared in include/arts_b200.h."""
    import re

    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", src)))

"""ctypes binding of the CPU oracle (oracle/pm_oracle.h).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

from polymutt_b200.capi import (PEEL_STEP_DTYPE, PERSON_RESULT_DTYPE, PERSON_SITE_DTYPE, SITE_HDR_DTYPE,
                                SITE_RESULT_DTYPE, Params, PedigreeArrays, _PmParams, _PmPedigree)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libpm_oracle.so"))
        lib.pmo_create.restype = C.c_void_p
        lib.pmo_create.argtypes = [C.POINTER(_PmPedigree), C.POINTER(_PmParams), C.c_void_p]
        lib.pmo_destroy.argtypes = [C.c_void_p]
        lib.pmo_call_glf_sites.restype = C.c_int
        lib.pmo_call_glf_sites.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pmo_call_vcf_records.restype = C.c_int
        lib.pmo_call_vcf_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.pmo_build_peel_order.restype = C.c_int
        lib.pmo_build_peel_order.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pmo_load_site.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pmo_family_loglik.restype = C.c_double
        lib.pmo_family_loglik.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
        lib.pmo_all_family_loglik.restype = C.c_double
        lib.pmo_all_family_loglik.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]
        lib.pmo_optimize.restype = C.c_double
        lib.pmo_optimize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        lib.pmo_fill_lut.argtypes = [C.c_void_p]
        lib.pmo_genotype_mutation_matrix.argtypes = [C.c_double, C.c_double, C.c_void_p]
        lib.pmo_last_error.restype = C.c_char_p
        _LIB = lib
    return _LIB


class OracleEngine:
    def __init__(self, ped: PedigreeArrays, params: Params, lut=None):
        self.lib = load()
        self.ped, self.params = ped, params
        self._cped, self._cpar = ped.to_c(), params.to_c()
        self._lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        self.ctx = self.lib.pmo_create(C.byref(self._cped), C.byref(self._cpar), None if self._lut is None else self._lut.ctypes.data)
        if not self.ctx:
            raise RuntimeError("pmo_create failed: " + self.lib.pmo_last_error().decode())

    def close(self):
        if self.ctx:
            self.lib.pmo_destroy(self.ctx)
            self.ctx = None

    def call_glf_sites(self, hdr, recs):
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        recs = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE)
        n, npers = len(hdr), self.ped.n_person
        status = np.zeros(n, dtype=np.uint16)
        res = np.zeros(n, dtype=SITE_RESULT_DTYPE)
        per = np.zeros((n, npers), dtype=PERSON_RESULT_DTYPE)
        rc = self.lib.pmo_call_glf_sites(self.ctx, hdr.ctypes.data, recs.ctypes.data, n, status.ctypes.data, res.ctypes.data, per.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle error {rc}: " + self.lib.pmo_last_error().decode())
        return status, res, per

    def call_vcf_records(self, hdr, recs, mono):
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        recs = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE)
        mono = np.ascontiguousarray(mono, dtype=np.float64)
        n, npers = len(hdr), self.ped.n_person
        res = np.zeros(n, dtype=SITE_RESULT_DTYPE)
        per = np.zeros((n, npers), dtype=PERSON_RESULT_DTYPE)
        rc = self.lib.pmo_call_vcf_records(self.ctx, hdr.ctypes.data, recs.ctypes.data, mono.ctypes.data, n, res.ctypes.data, per.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle error {rc}: " + self.lib.pmo_last_error().decode())
        return res, per

    def load_site(self, hdr1, recs1):
        self._h = np.ascontiguousarray(hdr1, dtype=SITE_HDR_DTYPE)
        self._r = np.ascontiguousarray(recs1, dtype=PERSON_SITE_DTYPE)
        self.lib.pmo_load_site(self.ctx, self._h.ctypes.data, self._r.ctypes.data)

    def optimize(self, a1, a2, denovo=False):
        f, n = C.c_double(0), C.c_int(0)
        ll = self.lib.pmo_optimize(self.ctx, a1, a2, int(denovo), C.byref(f), C.byref(n))
        return ll, f.value, n.value


def oracle_peel_order(father, mother, sex):
    lib = load()
    n = len(father)
    fa = np.ascontiguousarray(father, dtype=np.int32)
    mo = np.ascontiguousarray(mother, dtype=np.int32)
    sx = np.ascontiguousarray(sex, dtype=np.uint8)
    steps = np.zeros(n, dtype=PEEL_STEP_DTYPE)
    ns = lib.pmo_build_peel_order(n, fa.ctypes.data, mo.ctypes.data, sx.ctypes.data, steps.ctypes.data)
    if ns < 0:
        raise RuntimeError(lib.pmo_last_error().decode())
    return steps[:ns]

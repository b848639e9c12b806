"""ctypes binding of include/polymutt_b200.h.

Struct layouts mirror the header byte for byte (numpy structured dtypes for the bulk arrays,
ctypes.Structure for the small descriptors).  ``Engine`` wraps pm_create / pm_call_glf_sites /
pm_destroy; ``OracleEngine`` in tests/ wraps the CPU oracle with the same call signature.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class LibraryNotBuilt(RuntimeError):
    pass


def lib_path() -> str:
    # PM_LIB: an alternative build of the same library (scripts/gpu_phase_timing.py uses a -DPM_PHASE_TIMING one)
    return os.environ.get("PM_LIB") or os.path.join(_HERE, "lib", "libpolymutt_b200.so")


# ---- bulk record dtypes -------------------------------------------------------------------------
SITE_HDR_DTYPE = np.dtype([("pos", "<u4"), ("ref_base", "u1"), ("chr_class", "u1"), ("reserved", "<u2")])
PERSON_SITE_DTYPE = np.dtype([("lk", "u1", (10,)), ("depth", "u1", (3,)), ("map_quality", "u1"), ("pad", "u1", (2,))])
SITE_RESULT_DTYPE = np.dtype(
    [
        ("site", "<u4"), ("status", "u1"), ("maxidx", "i1"), ("allele1", "u1"), ("allele2", "u1"),
        ("n_hyp", "u1"), ("flags", "u1"), ("reserved", "<u2"),
        ("total_depth", "<i4"), ("num_samp", "<i4"),
        ("perc_samp", "<f8"), ("avg_map_qual", "<f8"), ("var_post_prob", "<f8"), ("poly_qual", "<f8"),
        ("freq", "<f8"), ("denovo_lr", "<f8"), ("ab", "<f8"),
        ("varllk", "<f8", (7,)), ("varllk_noprior", "<f8", (7,)), ("varfreq", "<f8", (7,)),
        ("refit_llk", "<f8"),
    ],
    align=True,
)
PERSON_RESULT_DTYPE = np.dtype(
    [("post", "<f8", (10,)), ("dosage", "<f8"), ("best", "<i4"), ("gq", "u1"), ("ten_state", "u1"), ("reserved", "u1", (2,))],
    align=True,
)
assert SITE_HDR_DTYPE.itemsize == 8 and PERSON_SITE_DTYPE.itemsize == 16
assert SITE_RESULT_DTYPE.itemsize == 256, SITE_RESULT_DTYPE.itemsize
assert PERSON_RESULT_DTYPE.itemsize == 96, PERSON_RESULT_DTYPE.itemsize

PEEL_STEP_DTYPE = np.dtype([("type", "<i4"), ("from0", "<i4"), ("from1", "<i4"), ("to0", "<i4"), ("to1", "<i4")])

# status codes / flags (include/polymutt_b200.h)
PM_SITE_EMITTED, PM_SITE_BAD_REF, PM_SITE_MIN_DEPTH, PM_SITE_MAX_DEPTH, PM_SITE_MIN_PS = 0, 1, 2, 3, 4
PM_SITE_MIN_MAPQ, PM_SITE_NOCALL, PM_SITE_MONO, PM_SITE_DENOVO_LOW_LR, PM_SITE_QUICK_SKIP = 5, 6, 7, 8, 9
PM_SITE_DENOVO_DROPPED = 10
PM_FLAG_NOCALL, PM_FLAG_ROW_DROPPED, PM_FLAG_MONO = 1, 2, 4
PM_OUT_EMITTED, PM_OUT_ALL = 0, 1
PM_OK, PM_EINVAL, PM_ECUDA, PM_ENOMEM, PM_EUNSUPPORTED = 0, -1, -2, -3, -4


class _PmPedigree(C.Structure):
    _fields_ = [
        ("n_fam", C.c_int32), ("n_person", C.c_int32),
        ("fam_size", C.c_void_p), ("fam_founders", C.c_void_p), ("fam_generations", C.c_void_p),
        ("sex", C.c_void_p), ("father", C.c_void_p), ("mother", C.c_void_p),
        ("peel_first", C.c_void_p), ("peel", C.c_void_p),
    ]


class _PmParams(C.Structure):
    _fields_ = [
        ("theta", C.c_double), ("theta_indel", C.c_double), ("poly_tstv", C.c_double),
        ("posterior_cutoff", C.c_double), ("precision", C.c_double), ("denovo_mut_rate", C.c_double),
        ("denovo_tstv", C.c_double), ("denovo_min_llr", C.c_double), ("min_ps", C.c_double),
        ("min_map_quality", C.c_int32), ("min_total_depth", C.c_int32), ("max_total_depth", C.c_int32),
        ("denovo", C.c_int32), ("force_call", C.c_int32), ("out_all_sites", C.c_int32),
        ("quick_call", C.c_int32), ("vcf_input", C.c_int32),
    ]


@dataclass
class Params:
    """pm_params with the reference's defaults (src/main.cpp:59-85)."""
    theta: float = 0.001
    theta_indel: float = 0.0001
    poly_tstv: float = 2.0
    posterior_cutoff: float = 0.5
    precision: float = 0.0001
    denovo_mut_rate: float = 1.5e-08
    denovo_tstv: float = 2.0
    denovo_min_llr: float = 0.01
    min_ps: float = 0.0
    min_map_quality: int = 0
    min_total_depth: int = 0
    max_total_depth: int = 0
    denovo: bool = False
    force_call: bool = False
    out_all_sites: bool = False
    quick_call: bool = False
    vcf_input: bool = False

    def to_c(self) -> _PmParams:
        p = _PmParams()
        for name, _ in _PmParams._fields_:
            v = getattr(self, name)
            setattr(p, name, int(v) if isinstance(v, bool) else v)
        return p


@dataclass
class PedigreeArrays:
    """Flat pedigree topology in VCF column order (pm_pedigree)."""
    fam_size: np.ndarray
    fam_founders: np.ndarray
    fam_generations: np.ndarray
    sex: np.ndarray
    father: np.ndarray
    mother: np.ndarray
    peel_first: Optional[np.ndarray] = None
    peel: Optional[np.ndarray] = None
    _keep: list = field(default_factory=list, repr=False)

    def __post_init__(self):
        self.fam_size = np.ascontiguousarray(self.fam_size, dtype=np.int32)
        self.fam_founders = np.ascontiguousarray(self.fam_founders, dtype=np.int32)
        self.fam_generations = np.ascontiguousarray(self.fam_generations, dtype=np.int32)
        self.sex = np.ascontiguousarray(self.sex, dtype=np.uint8)
        self.father = np.ascontiguousarray(self.father, dtype=np.int32)
        self.mother = np.ascontiguousarray(self.mother, dtype=np.int32)
        if self.peel_first is None:
            self.peel_first = np.zeros(len(self.fam_size) + 1, dtype=np.int32)
        self.peel_first = np.ascontiguousarray(self.peel_first, dtype=np.int32)
        if self.peel is None:
            self.peel = np.zeros(0, dtype=PEEL_STEP_DTYPE)
        self.peel = np.ascontiguousarray(self.peel, dtype=PEEL_STEP_DTYPE)

    @property
    def n_fam(self) -> int:
        return int(len(self.fam_size))

    @property
    def n_person(self) -> int:
        return int(len(self.sex))

    def family_first(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.fam_size)[:-1]]).astype(np.int32)

    def with_peel_orders(self, lib, all_families: bool = False) -> "PedigreeArrays":
        """Fills peel_first/peel for every extended family (all_families: also nuclear ones, which the VCF
        mode peels when the pedigree has a single family) with pm_build_peel_order."""
        firsts = self.family_first()
        steps_all, pf = [], [0]
        for f in range(self.n_fam):
            n, nf = int(self.fam_size[f]), int(self.fam_founders[f])
            nuclear = int(self.fam_generations[f]) == 2 and nf == 2 and not all_families
            if n != nf and not nuclear:
                a = int(firsts[f])
                steps = np.zeros(n, dtype=PEEL_STEP_DTYPE)
                fa = np.ascontiguousarray(self.father[a:a + n])
                mo = np.ascontiguousarray(self.mother[a:a + n])
                sx = np.ascontiguousarray(self.sex[a:a + n])
                ns = lib.pm_build_peel_order(n, fa.ctypes.data, mo.ctypes.data, sx.ctypes.data, steps.ctypes.data)
                if ns < 0:
                    raise RuntimeError(lib.pm_last_error().decode())
                steps_all.append(steps[:ns])
            pf.append(pf[-1] + (len(steps_all[-1]) if (n != nf and not nuclear) else 0))
        peel = np.concatenate(steps_all) if steps_all else np.zeros(0, dtype=PEEL_STEP_DTYPE)
        return PedigreeArrays(self.fam_size, self.fam_founders, self.fam_generations, self.sex, self.father,
                              self.mother, np.asarray(pf, dtype=np.int32), peel)

    def to_c(self) -> _PmPedigree:
        p = _PmPedigree()
        p.n_fam, p.n_person = self.n_fam, self.n_person
        p.fam_size = self.fam_size.ctypes.data
        p.fam_founders = self.fam_founders.ctypes.data
        p.fam_generations = self.fam_generations.ctypes.data
        p.sex = self.sex.ctypes.data
        p.father = self.father.ctypes.data
        p.mother = self.mother.ctypes.data
        p.peel_first = self.peel_first.ctypes.data
        p.peel = self.peel.ctypes.data if len(self.peel) else None
        return p


_LIB = None


def to_wire(recs: np.ndarray) -> np.ndarray:
    """16-byte pm_person_site records -> 14-byte pm_person_site_wire records (uint8 [..., 14])."""
    raw = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE).view(np.uint8).reshape(-1, 16)
    return np.ascontiguousarray(raw[:, :14])


def geno_index(b1: int, b2: int) -> int:
    """core/glfHandler.h:102-106 (bases 1..4)."""
    if b1 > b2:
        b1, b2 = b2, b1
    return (b1 - 1) * (10 - b1) // 2 + (b2 - b1)


def to_pl3(hdr: np.ndarray, recs: np.ndarray, n_person: int) -> np.ndarray:
    """VCF-input records -> the three PL bytes per sample (a1a1, a1a2, a2a2) pm_call_vcf_records_pl takes."""
    hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
    raw = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE).view(np.uint8).reshape(len(hdr), n_person, 16)
    out = np.zeros((len(hdr), n_person, 3), dtype=np.uint8)
    a1 = hdr["ref_base"].astype(np.int64)
    a2 = (hdr["reserved"] & 0xff).astype(np.int64)
    gi = np.vectorize(geno_index)
    if len(hdr):
        for k, g in enumerate((gi(a1, a1), gi(a1, a2), gi(a2, a2))):
            out[:, :, k] = np.take_along_axis(raw[:, :, :10], g[:, None, None].astype(np.int64), axis=2)[:, :, 0]
    return out


def _declare(lib):
    lib.pm_create.restype = C.c_void_p
    lib.pm_create.argtypes = [C.POINTER(_PmPedigree), C.POINTER(_PmParams), C.c_void_p, C.c_int]
    lib.pm_destroy.restype = None
    lib.pm_destroy.argtypes = [C.c_void_p]
    lib.pm_call_glf_sites.restype = C.c_int
    lib.pm_call_glf_sites.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.pm_call_glf_sites_wire.restype = C.c_int
    lib.pm_call_glf_sites_wire.argtypes = lib.pm_call_glf_sites.argtypes
    lib.pm_call_vcf_records_pl.restype = C.c_int
    lib.pm_call_vcf_records_pl.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.pm_call_vcf_records_calls_device.restype = C.c_int
    lib.pm_call_vcf_records_calls_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_call_vcf_records.restype = C.c_int
    lib.pm_call_vcf_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.pm_call_glf_sites_device.restype = C.c_int
    lib.pm_call_glf_sites_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.pm_call_vcf_records_calls.restype = C.c_int
    lib.pm_call_vcf_records_calls.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.pm_call_vcf_records_device.restype = C.c_int
    lib.pm_call_vcf_records_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_host_alloc.restype = C.c_void_p
    lib.pm_host_alloc.argtypes = [C.c_size_t]
    lib.pm_host_free.restype = None
    lib.pm_host_free.argtypes = [C.c_void_p]
    lib.pm_sync.restype = C.c_int
    lib.pm_sync.argtypes = [C.c_void_p]
    lib.pm_last_timing.restype = C.c_int
    lib.pm_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.pm_measure_fp64_peak.restype = C.c_int
    lib.pm_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pm_measure_copy_bw.restype = C.c_int
    lib.pm_measure_copy_bw.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.pm_describe_plan.restype = C.c_int
    lib.pm_describe_plan.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.pm_force_wide_plan.restype = C.c_int
    lib.pm_force_wide_plan.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.pm_timer_start.restype = C.c_int
    lib.pm_timer_start.argtypes = [C.c_void_p]
    lib.pm_timer_stop.restype = C.c_int
    lib.pm_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    lib.pm_get_counters.restype = C.c_int
    lib.pm_get_counters.argtypes = [C.c_void_p, C.c_void_p]
    lib.pm_reset_counters.restype = C.c_int
    lib.pm_reset_counters.argtypes = [C.c_void_p]
    lib.pm_last_error.restype = C.c_char_p
    lib.pm_last_error.argtypes = []
    lib.pm_abi_version.restype = C.c_int
    lib.pm_build_peel_order.restype = C.c_int
    lib.pm_build_peel_order.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pm_fill_lut.restype = None
    lib.pm_fill_lut.argtypes = [C.c_void_p]
    lib.pm_genotype_mutation_matrix.restype = None
    lib.pm_genotype_mutation_matrix.argtypes = [C.c_double, C.c_double, C.c_void_p]
    return lib


def load_library():
    """Loads the in-tree CUDA C-ABI library.  There is no fallback: a missing library is an error."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise LibraryNotBuilt(
                f"{path} not found: build it with `make lib` (or python -c 'import __graft_entry__ as g; g.build()'). "
                "polymutt_b200 has no CPU implementation of the likelihood path.")
        _LIB = _declare(C.CDLL(path))
    return _LIB


class Engine:
    """pm_ctx bound to one CUDA device."""

    def __init__(self, ped: PedigreeArrays, params: Params, device: int = 0, lut: Optional[np.ndarray] = None):
        self.lib = load_library()
        if len(ped.peel) == 0:
            ped = ped.with_peel_orders(self.lib, all_families=params.vcf_input)
        self.ped = ped
        self.params = params
        self._cped, self._cpar = ped.to_c(), params.to_c()
        self._lut = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
        self.ctx = self.lib.pm_create(C.byref(self._cped), C.byref(self._cpar), None if self._lut is None else self._lut.ctypes.data, device)
        if not self.ctx:
            raise RuntimeError("pm_create failed: " + self.lib.pm_last_error().decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.pm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise RuntimeError(f"pm error {rc}: " + self.lib.pm_last_error().decode())

    def call_glf_sites(self, hdr: np.ndarray, recs: np.ndarray, out_mode: int = PM_OUT_ALL, res_cap: Optional[int] = None):
        """Host-buffer entry point.  Returns (status[n_sites], results[n_res], persons[n_res, n_person])."""
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        recs = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE)
        n, npers = len(hdr), self.ped.n_person
        assert recs.size == n * npers, (recs.size, n, npers)
        cap = n if res_cap is None else res_cap
        status = np.zeros(n, dtype=np.uint16)
        res = np.zeros(max(cap, 1), dtype=SITE_RESULT_DTYPE)
        per = np.zeros((max(cap, 1), npers), dtype=PERSON_RESULT_DTYPE)
        n_res = C.c_size_t(0)
        self._check(self.lib.pm_call_glf_sites(self.ctx, hdr.ctypes.data, recs.ctypes.data, n, out_mode, status.ctypes.data,
                                               res.ctypes.data, per.ctypes.data, cap, C.byref(n_res)))
        k = n_res.value
        return status, res[:k], per[:k]

    def call_glf_sites_wire(self, hdr: np.ndarray, recs: np.ndarray, out_mode: int = PM_OUT_ALL, res_cap: Optional[int] = None):
        """pm_call_glf_sites_wire: the same call on 14-byte records (`recs` as for call_glf_sites; packed here)."""
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        wire = to_wire(recs)
        n, npers = len(hdr), self.ped.n_person
        assert wire.size == n * npers * 14, (wire.size, n, npers)
        cap = n if res_cap is None else res_cap
        status = np.zeros(n, dtype=np.uint16)
        res = np.zeros(max(cap, 1), dtype=SITE_RESULT_DTYPE)
        per = np.zeros((max(cap, 1), npers), dtype=PERSON_RESULT_DTYPE)
        n_res = C.c_size_t(0)
        self._check(self.lib.pm_call_glf_sites_wire(self.ctx, hdr.ctypes.data, wire.ctypes.data, n, out_mode, status.ctypes.data,
                                                    res.ctypes.data, per.ctypes.data, cap, C.byref(n_res)))
        k = n_res.value
        return status, res[:k], per[:k]

    def call_vcf_records_pl(self, hdr: np.ndarray, recs: np.ndarray, mono: np.ndarray):
        """pm_call_vcf_records_pl: VCF-input entry point on three PL bytes per sample (`recs` as for call_vcf_records;
        the triplets are taken from them here).  Returns (results[n], calls[n, n_person] = best | gq << 8)."""
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        mono = np.ascontiguousarray(mono, dtype=np.float64)
        n, npers = len(hdr), self.ped.n_person
        pl3 = to_pl3(hdr, recs, npers)
        res = np.zeros(max(n, 1), dtype=SITE_RESULT_DTYPE)
        calls = np.zeros((max(n, 1), npers), dtype=np.uint16)
        self._check(self.lib.pm_call_vcf_records_pl(self.ctx, hdr.ctypes.data, pl3.ctypes.data, mono.ctypes.data, n, res.ctypes.data, calls.ctypes.data))
        return res[:n], calls[:n]

    def call_vcf_records_calls_device(self, d_hdr: int, d_recs: int, d_mono: int, n: int, has_nonauto: bool, d_status: int, d_res: int, d_calls: int):
        """Device-buffer VCF entry point with 2 bytes per sample out (raw device pointers); asynchronous."""
        self._check(self.lib.pm_call_vcf_records_calls_device(self.ctx, d_hdr, d_recs, d_mono, n, 1 if has_nonauto else 0, d_status, d_res, d_calls))

    def call_vcf_records(self, hdr: np.ndarray, recs: np.ndarray, mono: np.ndarray):
        """VCF-input entry point.  Returns (results[n], persons[n, n_person])."""
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        recs = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE)
        mono = np.ascontiguousarray(mono, dtype=np.float64)
        n, npers = len(hdr), self.ped.n_person
        res = np.zeros(max(n, 1), dtype=SITE_RESULT_DTYPE)
        per = np.zeros((max(n, 1), npers), dtype=PERSON_RESULT_DTYPE)
        self._check(self.lib.pm_call_vcf_records(self.ctx, hdr.ctypes.data, recs.ctypes.data, mono.ctypes.data, n, res.ctypes.data, per.ctypes.data))
        return res[:n], per[:n]

    def call_glf_sites_device(self, d_hdr: int, d_recs: int, n_sites: int, out_mode: int, d_status: int, d_res: int,
                              d_person: int, res_cap: int, d_n_res: int):
        """Device-buffer entry point (raw device pointers, e.g. torch tensors' data_ptr()); asynchronous."""
        self._check(self.lib.pm_call_glf_sites_device(self.ctx, d_hdr, d_recs, n_sites, out_mode, d_status, d_res, d_person,
                                                      res_cap, d_n_res))

    def call_vcf_records_calls(self, hdr: np.ndarray, recs: np.ndarray, mono: np.ndarray):
        """VCF-input entry point with compact per-sample output.  Returns (results[n], calls[n, n_person] = best | gq << 8)."""
        hdr = np.ascontiguousarray(hdr, dtype=SITE_HDR_DTYPE)
        recs = np.ascontiguousarray(recs, dtype=PERSON_SITE_DTYPE)
        mono = np.ascontiguousarray(mono, dtype=np.float64)
        n, npers = len(hdr), self.ped.n_person
        res = np.zeros(max(n, 1), dtype=SITE_RESULT_DTYPE)
        calls = np.zeros((max(n, 1), npers), dtype=np.uint16)
        self._check(self.lib.pm_call_vcf_records_calls(self.ctx, hdr.ctypes.data, recs.ctypes.data, mono.ctypes.data, n, res.ctypes.data, calls.ctypes.data))
        return res[:n], calls[:n]

    def call_vcf_records_device(self, d_hdr: int, d_recs: int, d_mono: int, n: int, has_nonauto: bool, d_status: int, d_res: int, d_person: int):
        """Device-buffer VCF entry point (raw device pointers); asynchronous."""
        self._check(self.lib.pm_call_vcf_records_device(self.ctx, d_hdr, d_recs, d_mono, n, 1 if has_nonauto else 0, d_status, d_res, d_person))

    def sync(self):
        self._check(self.lib.pm_sync(self.ctx))

    def last_timing(self):
        a, b, n = C.c_float(0), C.c_float(0), C.c_int(0)
        self._check(self.lib.pm_last_timing(self.ctx, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def describe_plan(self) -> str:
        buf = C.create_string_buffer(512)
        self._check(self.lib.pm_describe_plan(self.ctx, buf, 512))
        return buf.value.decode()

    def force_wide_plan(self, variant: int, threads: int):
        """Test / tuning hook: run the main pass on instantiation `variant` of the block-per-site kernel."""
        self._check(self.lib.pm_force_wide_plan(self.ctx, variant, threads))

    def timer_start(self):
        self._check(self.lib.pm_timer_start(self.ctx))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.pm_timer_stop(self.ctx, C.byref(ms)))
        return ms.value

    def counters(self) -> dict:
        a = np.zeros(4, dtype=np.uint64)
        self._check(self.lib.pm_get_counters(self.ctx, a.ctypes.data))
        return dict(hypotheses=int(a[0]), evaluations=int(a[1]), sites_evaluated=int(a[2]), sites_emitted=int(a[3]))

    def reset_counters(self):
        self._check(self.lib.pm_reset_counters(self.ctx))

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0)
        self._check(self.lib.pm_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def measure_copy_bw(self) -> float:
        v = C.c_double(0)
        self._check(self.lib.pm_measure_copy_bw(self.ctx, C.byref(v)))
        return v.value

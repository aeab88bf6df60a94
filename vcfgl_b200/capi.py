"""ctypes binding of the C ABI in include/vgl.h (libvgl.so).

This is the only way Python reaches the CUDA kernels; there is no fallback: if
the shared library is missing or no CUDA device is present, loading or
`vgl_create` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import args as vargs

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvgl.so")
if os.environ.get("VGL_LIB"):      # development only: A/B builds of the same ABI (tools/gpu_variants.sh)
    LIB_PATH = os.environ["VGL_LIB"]

VGL_OK, VGL_EINVAL, VGL_ENOMEM, VGL_ECUDA, VGL_ESTATE, VGL_ERANGE, VGL_ENODEV, VGL_EOVERFLOW, VGL_EMISSING = 0, -1, -2, -3, -4, -5, -6, -7, -8
ABI_VERSION = 9
HOST_NONE, HOST_I32, HOST_NARROW, HOST_BCF, HOST_BGZF = 0, 1, 2, 3, 4
T_H2D, T_SIM, T_SITE, T_SCAN, T_EMIT, T_D2H, T_TOTAL, T_COUNT = range(8)
SUBMIT_GT_ON_DEVICE = 1
F32_MISSING_BITS = 0x7F800001
I32_MISSING = -(2 ** 31)

EXPORTS = ["vgl_set_gvcf_dps", "vgl_gvcf_flush", "vgl_create", "vgl_destroy", "vgl_input_buffer", "vgl_bcf_input_buffer", "vgl_submit", "vgl_wait", "vgl_set_stream",
           "vgl_slot_timing", "vgl_copy_sites", "vgl_native_draws", "vgl_selftest", "vgl_launch_count", "vgl_algorithmic_bytes", "vgl_strerror",
           "vgl_last_error", "vgl_abi_version", "vgl_native_kernels",
           "vgl_gvcf_merge", "vgl_discordance", "vgl_parser_create", "vgl_parser_destroy", "vgl_parser_text_buffer", "vgl_parse_vcf", "vgl_parse_bcf", "vgl_parser_rows", "vgl_place_rows"]

SOURCE_BINARY, SOURCE_ACGT = 0, 1
PARSE_FINAL, PARSE_TEXT_ON_DEVICE = 1, 2
(IN_OK, IN_ENCOLS, IN_EPOS, IN_ENALLELE, IN_EALLELE, IN_ENOGT, IN_ENSAMPLES, IN_EGTCHAR, IN_EPLOIDY, IN_EALLELEIDX,
 IN_ESYMBOLIC) = range(11)
IN_SITE_DTYPE = np.dtype([("status", "<i4"), ("skip_code", "<i4"), ("pos", "<i8"), ("allele_sum", "<i8"), ("line_off", "<u8"),
                          ("line_len", "<u4"), ("n_allele", "<i4"), ("allele_acgt", "i1", (8,)), ("id_off", "<u4"),
                          ("fmt_off", "<u4"), ("samples_off", "<u4"), ("_pad", "<u4")])
assert IN_SITE_DTYPE.itemsize == 64


GVCF_SITE_IN_DTYPE = np.dtype([("rid", "<i4"), ("pos", "<i4")])
GVCF_REC_DTYPE = np.dtype([("first_site", "<i4"), ("last_site", "<i4"), ("n_members", "<i4"), ("min_dp", "<i4"), ("dp_range", "<i4"),
                           ("plane", "<i4")])


class VglGvcfOut(C.Structure):
    _fields_ = [("n_recs", C.c_int32), ("n_blocks", C.c_int32), ("recs", C.c_void_p), ("dp", C.c_void_p), ("pl", C.c_void_p),
                ("ms_kernels", C.c_float)]


class VglDiscordanceOut(C.Structure):
    _fields_ = [("n_hom", C.c_int64), ("n_hom_discordant", C.c_int64), ("n_het", C.c_int64), ("n_het_discordant", C.c_int64),
                ("ms_kernel", C.c_float)]


class VglParseOut(C.Structure):
    _fields_ = [("n_records", C.c_int32), ("n_errors", C.c_int32), ("first_error_record", C.c_int32), ("n_kept", C.c_int32),
                ("bytes_consumed", C.c_int64), ("sites", C.c_void_p), ("ms_h2d", C.c_float), ("ms_kernels", C.c_float)]


class VglBcfDict(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("dp", "gl", "pl", "gp", "ad", "adf", "adr", "qs", "i16", "end", "min_dp")]


class VglBcfSiteIn(C.Structure):
    _fields_ = [("rid", C.c_int32), ("pos", C.c_int32), ("qual_bits", C.c_uint32), ("n_info", C.c_uint32),
                ("id_off", C.c_uint32), ("id_len", C.c_uint32), ("flt_info_off", C.c_uint32), ("flt_info_len", C.c_uint32),
                ("fmt_off", C.c_uint32), ("fmt_len", C.c_uint32), ("n_fmt", C.c_uint32), ("_pad", C.c_uint32)]


BCF_SITE_IN_DTYPE = np.dtype([("rid", "<i4"), ("pos", "<i4"), ("qual_bits", "<u4"), ("n_info", "<u4"), ("id_off", "<u4"),
                              ("id_len", "<u4"), ("flt_info_off", "<u4"), ("flt_info_len", "<u4"), ("fmt_off", "<u4"), ("fmt_len", "<u4"),
                              ("n_fmt", "<u4"), ("_pad", "<u4")])
assert BCF_SITE_IN_DTYPE.itemsize == C.sizeof(VglBcfSiteIn)


class VglParams(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_samples", C.c_int32), ("seed", C.c_int64),
                ("depth_mode", C.c_int32), ("depth_mean", C.c_double), ("depth_means", C.POINTER(C.c_double)),
                ("error_rate", C.c_double), ("error_qs", C.c_int32), ("beta_variance", C.c_double),
                ("gl_model", C.c_int32), ("gl1_theta", C.c_double), ("precise_gl", C.c_int32),
                ("adjust_qs", C.c_int32), ("adjust_by", C.c_double), ("n_qs_bins", C.c_int32),
                ("qs_bins", (C.c_uint8 * 3) * 255), ("do_unobserved", C.c_int32),
                ("rm_invar_sites", C.c_int32), ("rm_empty_sites", C.c_int32), ("do_gvcf", C.c_int32),
                ("tag_mask", C.c_uint32), ("i16_mapq", C.c_int32), ("device_id", C.c_int32),
                ("max_batch_sites", C.c_int32), ("n_slots", C.c_int32), ("sampler", C.c_int32),
                ("host_output", C.c_int32), ("bcf_dict", VglBcfDict), ("bcf_blob_bytes_per_site", C.c_int32)]


class VglReplay(C.Structure):
    _fields_ = [("depths", C.c_void_p), ("read_offsets", C.c_void_p), ("n_reads", C.c_int64),
                ("bases", C.c_void_p), ("strands", C.c_void_p), ("qs", C.c_void_p), ("adj_qs", C.c_void_p),
                ("error_probs", C.c_void_p), ("tail_dists", C.c_void_p), ("n_deep_cells", C.c_int64),
                ("deep_codes", C.c_void_p)]


class VglSiteOut(C.Structure):
    _fields_ = [("skip_code", C.c_int32), ("n_alleles", C.c_int32), ("n_alleles_observed", C.c_int32),
                ("n_genotypes", C.c_int32), ("alleles2acgt", C.c_int8 * 8), ("acgt2alleles", C.c_int8 * 8),
                ("info_dp", C.c_int32), ("info_ad", C.c_int32 * 5), ("info_adf", C.c_int32 * 5),
                ("info_adr", C.c_int32 * 5), ("qs", C.c_float * 5), ("i16", C.c_float * 16),
                ("_pad", C.c_int32), ("g_off", C.c_int64), ("r_off", C.c_int64)]


SITE_DTYPE = np.dtype([("skip_code", "<i4"), ("n_alleles", "<i4"), ("n_alleles_observed", "<i4"),
                       ("n_genotypes", "<i4"), ("alleles2acgt", "i1", 8), ("acgt2alleles", "i1", 8),
                       ("info_dp", "<i4"), ("info_ad", "<i4", 5), ("info_adf", "<i4", 5), ("info_adr", "<i4", 5),
                       ("qs", "<f4", 5), ("i16", "<f4", 16), ("_pad", "<i4"), ("g_off", "<i8"), ("r_off", "<i8")])
assert SITE_DTYPE.itemsize == C.sizeof(VglSiteOut), (SITE_DTYPE.itemsize, C.sizeof(VglSiteOut))


class VglBatchOut(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("n_samples", C.c_int32), ("sites", C.POINTER(VglSiteOut)),
                ("dp", C.c_void_p), ("gl", C.c_void_p), ("pl", C.c_void_p), ("gp", C.c_void_p),
                ("ad", C.c_void_p), ("adf", C.c_void_p), ("adr", C.c_void_p),
                ("g_elems", C.c_int64), ("r_elems", C.c_int64), ("status", C.c_int32),
                ("narrow_bits", C.c_int32), ("pl_u8", C.c_void_p), ("dp_n", C.c_void_p), ("ad_n", C.c_void_p),
                ("adf_n", C.c_void_p), ("adr_n", C.c_void_p),
                ("bcf", C.c_void_p), ("bcf_off", C.c_void_p), ("bcf_bytes", C.c_int64),
                ("bgzf", C.c_void_p), ("bgzf_bytes", C.c_int64), ("bgzf_blocks", C.c_int32), ("n_recs", C.c_int32)]


class VglDraws(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_reads", C.c_int64), ("depths", C.c_void_p),
                ("read_offsets", C.c_void_p), ("bases", C.c_void_p), ("strands", C.c_void_p),
                ("qs", C.c_void_p), ("adj_qs", C.c_void_p), ("tail_dists", C.c_void_p),
                ("error_probs", C.c_void_p)]


class VglError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libvgl: %s (status %d)" % (msg, code))
        self.code = code


_lib = None


def load():
    """Load libvgl.so; raises if it was not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: the CUDA extension must be built (make -C vcfgl_b200/csrc); "
                          "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.vgl_create.argtypes = [C.POINTER(VglParams), C.POINTER(C.c_void_p)]
    L.vgl_destroy.argtypes = [C.c_void_p]
    L.vgl_destroy.restype = None
    L.vgl_input_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.vgl_bcf_input_buffer.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.vgl_submit.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.POINTER(VglReplay), C.c_uint32]
    L.vgl_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(VglBatchOut)]
    L.vgl_set_stream.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.vgl_slot_timing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.vgl_copy_sites.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.vgl_native_draws.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.POINTER(VglDraws)]
    L.vgl_selftest.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]
    L.vgl_launch_count.argtypes = [C.c_void_p]
    L.vgl_launch_count.restype = C.c_int64
    L.vgl_native_kernels.argtypes = [C.c_void_p]
    L.vgl_native_kernels.restype = C.c_char_p
    L.vgl_algorithmic_bytes.argtypes = [C.POINTER(VglBatchOut), C.c_uint32]
    L.vgl_algorithmic_bytes.restype = C.c_int64
    L.vgl_strerror.argtypes = [C.c_int]
    L.vgl_strerror.restype = C.c_char_p
    L.vgl_last_error.argtypes = [C.c_void_p]
    L.vgl_last_error.restype = C.c_char_p
    L.vgl_abi_version.restype = C.c_int
    L.vgl_gvcf_merge.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(VglGvcfOut)]
    L.vgl_set_gvcf_dps.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.vgl_gvcf_flush.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.vgl_discordance.argtypes = [C.c_void_p, C.c_int, C.POINTER(VglDiscordanceOut)]
    L.vgl_parser_create.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]
    L.vgl_parser_destroy.argtypes = [C.c_void_p]
    L.vgl_parser_destroy.restype = None
    L.vgl_parser_text_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.vgl_parse_vcf.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.POINTER(VglParseOut)]
    L.vgl_parse_bcf.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.POINTER(VglParseOut)]
    L.vgl_parser_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.vgl_place_rows.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint8]
    _lib = L
    return L


def params_from_args(a: vargs.SimArgs, n_samples: int, max_batch_sites: int, n_slots: int = 2,
                     device_id: int = 0, host_output=True, sampler: int = 0,
                     fixed_depth: bool = False, bcf_dict: Optional[dict] = None, bcf_blob_bytes_per_site: int = 0) -> VglParams:
    """SimArgs (the reference CLI contract) -> vgl_params"""
    p = VglParams()
    p.abi_version = ABI_VERSION
    p.n_samples = n_samples
    p.seed = a.seed if a.seed != -1 else 0
    if a.depths is not None:
        p.depth_mode = vargs.DEPTH_POISSON_PER_SAMPLE
        arr = (C.c_double * n_samples)(*a.depths)
        p._keep = arr
        p.depth_means = C.cast(arr, C.POINTER(C.c_double))
        p.depth_mean = 0.0
    else:
        p.depth_mode = vargs.DEPTH_FIXED if fixed_depth else vargs.DEPTH_POISSON
        p.depth_mean = a.depth
        if a.depth == float("inf"):      # --depth inf: truth mode (vcfgl.cpp:1089-1262)
            p.depth_mode, p.depth_mean = vargs.DEPTH_INF, 0.0
    p.error_rate = a.error_rate
    p.error_qs = a.error_qs
    p.beta_variance = a.beta_variance
    p.gl_model = a.gl_model
    p.gl1_theta = a.gl1_theta
    p.precise_gl = a.precise_gl
    p.adjust_qs = a.adjust_qs
    p.adjust_by = a.adjust_by
    bins = a.qs_bins or []
    p.n_qs_bins = len(bins)
    for i, (s, e, q) in enumerate(bins):
        p.qs_bins[i][0], p.qs_bins[i][1], p.qs_bins[i][2] = s, e, q
    p.do_unobserved = a.do_unobserved
    p.rm_invar_sites = a.rm_invar_sites
    p.rm_empty_sites = a.rm_empty_sites
    p.do_gvcf = a.do_gvcf
    p.tag_mask = a.tag_mask
    p.i16_mapq = a.i16_mapq
    p.device_id = device_id
    p.max_batch_sites = max_batch_sites
    p.n_slots = n_slots
    p.sampler = sampler
    p.host_output = int(host_output)      # False / True / HOST_NARROW / HOST_BCF
    if bcf_dict:                          # {"DP": id, "GL": id, ...}: bcf_hdr_id2int() of the output header
        for k, v in bcf_dict.items():
            setattr(p.bcf_dict, k.lower(), int(v))
    p.bcf_blob_bytes_per_site = bcf_blob_bytes_per_site
    return p


class Batch:
    """A finished batch: numpy views over the pinned host result buffers (host_output=1 or 2).

    With HOST_NARROW the integer planes arrive narrowed (pl_u8, dp_n, ad_n, ...); `dp`, `pl`, `ad`, `adf`, `adr`
    are then widened copies (what a host would hand to bcf_update_format_int32), missing PL restored from DP == 0."""

    def __init__(self, out: VglBatchOut, tag_mask: int, host: bool):
        self.raw = out
        self.n_sites = out.n_sites
        self.S = out.n_samples
        self.status = out.status
        self.g_elems = out.g_elems
        self.r_elems = out.r_elems
        self.host = host
        self.tag_mask = tag_mask
        self.sites = None
        if host:
            self.sites = np.ctypeslib.as_array(C.cast(out.sites, C.POINTER(C.c_uint8)),
                                               shape=(out.n_sites * SITE_DTYPE.itemsize,)).view(SITE_DTYPE)

        def view(ptr, dtype, n):
            if not ptr or not host:
                return None
            ct = {np.float32: C.c_float, np.int32: C.c_int32}[dtype]
            if n == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
        self._i32 = dict(dp=view(out.dp, np.int32, out.n_sites * out.n_samples), pl=view(out.pl, np.int32, out.g_elems),
                         ad=view(out.ad, np.int32, out.r_elems), adf=view(out.adf, np.int32, out.r_elems),
                         adr=view(out.adr, np.int32, out.r_elems))
        self.gl = view(out.gl, np.float32, out.g_elems)
        self.gp = view(out.gp, np.float32, out.g_elems)
        self.narrow_bits = int(out.narrow_bits)
        # HOST_BCF: the serialised records and their byte offsets (views over pinned memory)
        self.bcf = self.bcf_off = None
        self.bcf_bytes = int(out.bcf_bytes)
        self.bgzf, self.bgzf_bytes, self.bgzf_blocks = None, int(out.bgzf_bytes), int(out.bgzf_blocks)
        self.n_recs = int(out.n_recs)
        if (out.bcf_off or out.bcf) and host:
            if out.bcf_off:
                self.bcf_off = np.ctypeslib.as_array(C.cast(out.bcf_off, C.POINTER(C.c_int64)), shape=(out.n_sites + 1,))
            if out.bcf:
                self.bcf = (np.ctypeslib.as_array(C.cast(out.bcf, C.POINTER(C.c_uint8)), shape=(self.bcf_bytes,))
                            if self.bcf_bytes > 0 else np.zeros(0, np.uint8))
            if out.bgzf:      # HOST_BGZF: the record stream as BGZF blocks (views over pinned memory)
                self.bgzf = (np.ctypeslib.as_array(C.cast(out.bgzf, C.POINTER(C.c_uint8)), shape=(self.bgzf_bytes,))
                             if self.bgzf_bytes > 0 else np.zeros(0, np.uint8))
        self.pl_u8 = self.dp_n = self.ad_n = self.adf_n = self.adr_n = None
        if self.narrow_bits and host:
            ct, dt = (C.c_uint8, np.uint8) if self.narrow_bits == 8 else (C.c_uint16, np.uint16)

            def nview(ptr, ct, n, dt):
                if not ptr:
                    return None
                if n == 0:
                    return np.zeros(0, dt)
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
            self.pl_u8 = nview(out.pl_u8, C.c_uint8, out.g_elems, np.uint8)
            self.dp_n = nview(out.dp_n, ct, out.n_sites * out.n_samples, dt)
            self.ad_n = nview(out.ad_n, ct, out.r_elems, dt)
            self.adf_n = nview(out.adf_n, ct, out.r_elems, dt)
            self.adr_n = nview(out.adr_n, ct, out.r_elems, dt)

    def _plane(self, k):
        """int32 plane `k`; with HOST_NARROW a widened copy made on first use (PL: see site())"""
        v = self._i32[k]
        if v is None and self.narrow_bits and k != "pl":
            n = getattr(self, k + "_n")
            if n is not None:
                v = self._i32[k] = n.astype(np.int32)
        return v

    dp = property(lambda self: self._plane("dp"))
    pl = property(lambda self: self._plane("pl"))
    ad = property(lambda self: self._plane("ad"))
    adf = property(lambda self: self._plane("adf"))
    adr = property(lambda self: self._plane("adr"))

    def site(self, i: int) -> dict:
        """Everything add_tags() (bcf_utils.cpp:426-507) would emit for site i, as arrays."""
        s = self.sites[i]
        S, G, A = self.S, int(s["n_genotypes"]), int(s["n_alleles"])
        d = dict(skip_code=int(s["skip_code"]), n_alleles=A, n_alleles_observed=int(s["n_alleles_observed"]),
                 n_genotypes=G, alleles2acgt=s["alleles2acgt"][:5].astype(np.int32),
                 acgt2alleles=s["acgt2alleles"][:5].astype(np.int32), info_dp=int(s["info_dp"]),
                 info_ad=s["info_ad"][:A].copy(), info_adf=s["info_adf"][:A].copy(), info_adr=s["info_adr"][:A].copy(),
                 qs=s["qs"][:A].copy(), i16=s["i16"].copy(),
                 fmt_dp=self.dp[i * S:(i + 1) * S] if self.dp is not None else None)
        if d["skip_code"] == 0:
            g0, r0 = int(s["g_off"]), int(s["r_off"])
            for k, plane in (("gl", self.gl), ("pl", self.pl), ("gp", self.gp)):
                d[k] = plane[g0:g0 + S * G] if plane is not None else None
            if self.pl_u8 is not None:      # narrow PL: a cell without reads has a missing PL (vgl.h)
                pl = self.pl_u8[g0:g0 + S * G].astype(np.int32).reshape(S, G)
                pl[d["fmt_dp"] == 0, :] = I32_MISSING
                d["pl"] = pl.reshape(-1)
            for k, plane in (("fmt_ad", self.ad), ("fmt_adf", self.adf), ("fmt_adr", self.adr)):
                d[k] = plane[r0:r0 + S * A] if plane is not None else None
        return d


class Context:
    """Owns one vgl_ctx (one GPU)."""

    def __init__(self, params: VglParams):
        self.L = load()
        self.params = params
        self.h = C.c_void_p()
        rc = self.L.vgl_create(C.byref(params), C.byref(self.h))
        if rc != VGL_OK:
            raise VglError(rc, self.L.vgl_strerror(rc).decode())
        self.S = params.n_samples
        self.cap = params.max_batch_sites
        self._keep = {}

    def close(self):
        if self.h:
            self.L.vgl_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != VGL_OK:
            raise VglError(rc, "%s: %s" % (self.L.vgl_strerror(rc).decode(), self.L.vgl_last_error(self.h).decode()))

    def input_buffer(self, slot: int) -> np.ndarray:
        ptr, cap = C.c_void_p(), C.c_int64()
        self._ck(self.L.vgl_input_buffer(self.h, slot, C.byref(ptr), C.byref(cap)))
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(cap.value, self.S))

    def bcf_input(self, slot: int):
        """HOST_BCF: (per-site pass-through records [cap] as a structured array, blob bytes) of a slot, pinned"""
        ps, pb, cap = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._ck(self.L.vgl_bcf_input_buffer(self.h, slot, C.byref(ps), C.byref(pb), C.byref(cap)))
        sites = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_uint8)), shape=(self.cap * BCF_SITE_IN_DTYPE.itemsize,)).view(BCF_SITE_IN_DTYPE)
        blob = np.ctypeslib.as_array(C.cast(pb, C.POINTER(C.c_uint8)), shape=(cap.value,))
        return sites, blob

    def set_stream(self, slot: int, stream_ptr: Optional[int]):
        self._ck(self.L.vgl_set_stream(self.h, slot, C.c_void_p(stream_ptr or 0)))

    def submit(self, slot: int, first_site_id: int, n_sites: int, replay: Optional[dict] = None, flags: int = 0):
        rp = None
        if replay is not None:
            keep = []

            def ptr(x, dt):
                if x is None:
                    return None
                x = np.ascontiguousarray(x, dtype=dt)
                keep.append(x)
                return x.ctypes.data
            r = VglReplay()
            r.depths = ptr(replay["depths"], np.int32)
            r.read_offsets = ptr(replay["read_offsets"], np.int64)
            r.n_reads = int(replay["n_reads"])
            r.bases = ptr(replay.get("bases"), np.uint8)
            r.strands = ptr(replay.get("strands"), np.uint8)
            r.qs = ptr(replay.get("qs"), np.uint8)
            r.adj_qs = ptr(replay.get("adj_qs"), np.uint8)
            r.error_probs = ptr(replay.get("error_probs"), np.float64)
            r.tail_dists = ptr(replay.get("tail_dists"), np.uint8)
            r.n_deep_cells = int(replay.get("n_deep_cells", 0))
            r.deep_codes = ptr(replay.get("deep_codes"), np.uint16)
            self._keep[slot] = keep
            rp = C.byref(r)
        self._ck(self.L.vgl_submit(self.h, slot, first_site_id, n_sites, rp, flags))

    def wait(self, slot: int) -> Batch:
        out = VglBatchOut()
        self._ck(self.L.vgl_wait(self.h, slot, C.byref(out)))
        return Batch(out, self.params.tag_mask, bool(self.params.host_output))

    def sync(self, slot: int) -> int:
        """vgl_wait without building the numpy views: blocks until the slot's batch is complete, returns its status"""
        out = VglBatchOut()
        self._ck(self.L.vgl_wait(self.h, slot, C.byref(out)))
        return int(out.status)

    def copy_sites(self, slot: int, batch: Batch) -> np.ndarray:
        """host_output=0: fetch the per-site records of a waited slot (also stored on `batch.sites`)"""
        arr = np.zeros(batch.n_sites, SITE_DTYPE)
        self._ck(self.L.vgl_copy_sites(self.h, slot, arr.ctypes.data))
        batch.sites = arr
        return arr

    def native_draws(self, slot: int, first_site_id: int, n_sites: int) -> dict:
        """The Philox simulator's own draws for the genotypes in the slot's input buffer, as a replay
        dict (copies) that can be passed back to submit(replay=...) or to the CPU oracle."""
        d = VglDraws()
        self._ck(self.L.vgl_native_draws(self.h, slot, first_site_id, n_sites, C.byref(d)))

        def arr(ptr, ct, n, dt):
            if not ptr or n == 0:
                return None if not ptr else np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()
        nc, nr = d.n_cells, d.n_reads
        return dict(depths=arr(d.depths, C.c_int32, nc, np.int32),
                    read_offsets=arr(d.read_offsets, C.c_int64, nc + 1, np.int64), n_reads=int(nr),
                    bases=arr(d.bases, C.c_uint8, nr, np.uint8), strands=arr(d.strands, C.c_uint8, nr, np.uint8),
                    qs=arr(d.qs, C.c_uint8, nr, np.uint8), adj_qs=arr(d.adj_qs, C.c_uint8, nr, np.uint8),
                    tail_dists=arr(d.tail_dists, C.c_uint8, nr, np.uint8),
                    error_probs=arr(d.error_probs, C.c_double, nr, np.float64))

    def timing(self, slot: int) -> np.ndarray:
        ms = (C.c_float * T_COUNT)()
        self._ck(self.L.vgl_slot_timing(self.h, slot, ms))
        return np.array(ms[:], np.float64)

    def launch_count(self) -> int:
        return int(self.L.vgl_launch_count(self.h))

    def native_kernels(self):
        return self.L.vgl_native_kernels(self.h).decode()

    def algorithmic_bytes(self, batch: Batch) -> int:
        """SURVEY.md 8(d) bytes of a batch; needs the site records on the host"""
        assert batch.sites is not None, "call copy_sites() first (host_output=0)"
        raw = VglBatchOut()
        C.memmove(C.byref(raw), C.byref(batch.raw), C.sizeof(raw))
        raw.sites = C.cast(batch.sites.ctypes.data, C.POINTER(VglSiteOut))
        return int(self.L.vgl_algorithmic_bytes(C.byref(raw), self.params.tag_mask))


    def gvcf_merge(self, slot: int, rid, pos, gvcf_dps) -> dict:
        """gVCF block merger over the slot's last (waited) batch: -> dict(recs structured array [n_recs], dp [n_blocks, S],
        pl [n_blocks, S, 3] or None, ms_kernels); arrays are copies"""
        sin = np.zeros(len(rid), GVCF_SITE_IN_DTYPE)
        sin["rid"], sin["pos"] = rid, pos
        dps = np.ascontiguousarray(gvcf_dps, np.int32)
        out = VglGvcfOut()
        self._ck(self.L.vgl_gvcf_merge(self.h, slot, sin.ctypes.data, dps.ctypes.data, len(dps), C.byref(out)))

        def arr(ptr, n, dt):
            if not ptr or n == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt).copy()
        S = self.S
        return dict(recs=arr(out.recs, out.n_recs, GVCF_REC_DTYPE), dp=arr(out.dp, out.n_blocks * S, np.int32).reshape(-1, S),
                    pl=arr(out.pl, out.n_blocks * S * 3, np.int32).reshape(-1, S, 3) if out.pl else None,
                    n_blocks=out.n_blocks, ms_kernels=out.ms_kernels)

    def set_gvcf_dps(self, gvcf_dps):
        """HOST_BCF with -doGVCF: the --gvcf-dps thresholds, before the first submit"""
        d = np.ascontiguousarray(gvcf_dps, np.int32)
        self._ck(self.L.vgl_set_gvcf_dps(self.h, d.ctypes.data, len(d)))

    def gvcf_flush(self) -> bytes:
        """HOST_BCF with -doGVCF: the block record still open after the last batch (b"" when there is none)"""
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self.L.vgl_gvcf_flush(self.h, C.byref(p), C.byref(n)))
        return C.string_at(p, n.value) if n.value else b""

    def discordance(self, slot: int) -> dict:
        """genotype-call discordance of the slot's last (waited) batch: {"hom": [cells, discordant], "het": [...], "ms_kernel"}"""
        out = VglDiscordanceOut()
        self._ck(self.L.vgl_discordance(self.h, slot, C.byref(out)))
        return dict(hom=[out.n_hom, out.n_hom_discordant], het=[out.n_het, out.n_het_discordant], ms_kernel=out.ms_kernel)

    def parser(self, max_text_bytes: int, max_records: int) -> "Parser":
        return Parser(self, max_text_bytes, max_records)

    def place_rows(self, slot: int, parser: "Parser", n_sites: int, row_map=None, first_record: int = 0, fill_gt: int = 0):
        """slot genotype matrix <- parsed rows (vgl_place_rows); follow with submit(..., flags=SUBMIT_GT_ON_DEVICE)"""
        mp = None
        if row_map is not None:
            row_map = np.ascontiguousarray(row_map, np.int32)
            assert len(row_map) == n_sites
            mp = row_map.ctypes.data
        self._ck(self.L.vgl_place_rows(self.h, slot, parser.h, mp, first_record, n_sites, fill_gt))


class ParseResult:
    def __init__(self, out: VglParseOut):
        self.n_records, self.n_errors, self.first_error_record = out.n_records, out.n_errors, out.first_error_record
        self.n_kept, self.bytes_consumed = out.n_kept, out.bytes_consumed
        self.ms_h2d, self.ms_kernels = out.ms_h2d, out.ms_kernels
        if out.n_records:
            self.sites = np.ctypeslib.as_array(C.cast(out.sites, C.POINTER(C.c_uint8)),
                                               shape=(out.n_records * IN_SITE_DTYPE.itemsize,)).view(IN_SITE_DTYPE)
        else:
            self.sites = np.zeros(0, IN_SITE_DTYPE)


class Parser:
    """Device-side VCF text -> packed-genotype parser of a context (include/vgl.h, "Input path")."""

    def __init__(self, ctx: Context, max_text_bytes: int, max_records: int):
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        ctx._ck(self.L.vgl_parser_create(ctx.h, max_text_bytes, max_records, C.byref(self.h)))
        ptr, cap = C.c_void_p(), C.c_int64()
        ctx._ck(self.L.vgl_parser_text_buffer(self.h, C.byref(ptr), C.byref(cap)))
        self.text = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(cap.value,))
        self.S = ctx.S

    def close(self):
        if self.h:
            self.L.vgl_parser_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def parse(self, data=None, gt_source: int = SOURCE_BINARY, flags: int = 0, n_bytes: Optional[int] = None) -> ParseResult:
        """`data` (bytes / uint8 array) is staged into the pinned text buffer first; pass None with n_bytes when the
        caller filled `self.text` itself (or with PARSE_TEXT_ON_DEVICE)."""
        if data is not None:
            a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
            n_bytes = len(a)
            self.text[:n_bytes] = a
        out = VglParseOut()
        rc = self.L.vgl_parse_vcf(self.h, int(n_bytes), gt_source, flags, C.byref(out))
        if rc != VGL_OK:
            raise VglError(rc, self.L.vgl_strerror(rc).decode())
        return ParseResult(out)

    def parse_bcf(self, data, rec_off, gt_key: int, gt_source: int = SOURCE_BINARY, flags: int = 0) -> ParseResult:
        """uncompressed BCF records (bytes after the header) + their offsets [n + 1] (vgl_parse_bcf)"""
        a = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else data
        self.text[:len(a)] = a
        off = np.ascontiguousarray(rec_off, np.uint32)
        out = VglParseOut()
        rc = self.L.vgl_parse_bcf(self.h, len(a), off.ctypes.data, len(off) - 1, gt_source, gt_key, flags, C.byref(out))
        if rc != VGL_OK:
            raise VglError(rc, self.L.vgl_strerror(rc).decode())
        return ParseResult(out)

    def rows(self, first_record: int, n_records: int) -> np.ndarray:
        out = np.empty((n_records, self.S), np.uint8)
        if n_records == 0:
            return out
        rc = self.L.vgl_parser_rows(self.h, first_record, n_records, out.ctypes.data)
        if rc != VGL_OK:
            raise VglError(rc, self.L.vgl_strerror(rc).decode())
        return out

"""Host-side mirror of the reference's command-line contract for the hot path.

Mirrors the option names, defaults and combination rules of the reference CLI
(io.cpp:428-526 defaults, io.cpp:538-752 option names, io.cpp:860-1000
validation) for exactly the fields that `simulate_record_values`
(vcfgl.cpp:327) and `calculate_gls` (gl_methods.cpp) read.  I/O options
(`-i`, `-o`, `-O`, `--threads`, print* switches) are accepted and kept but are
outside the accelerated path.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence

# tag bits of vgl_params.tag_mask (include/vgl.h)
TAG_GL, TAG_GP, TAG_PL, TAG_I16, TAG_QS = 1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4
TAG_FMT_DP, TAG_INFO_DP = 1 << 5, 1 << 6
TAG_FMT_AD, TAG_INFO_AD = 1 << 7, 1 << 8
TAG_FMT_ADF, TAG_INFO_ADF = 1 << 9, 1 << 10
TAG_FMT_ADR, TAG_INFO_ADR = 1 << 11, 1 << 12

DEPTH_POISSON, DEPTH_POISSON_PER_SAMPLE, DEPTH_FIXED, DEPTH_INF = 0, 1, 2, 3


class ArgError(ValueError):
    """Raised where the reference would ERROR()/exit(1) (shared.h:292-327)."""


@dataclasses.dataclass
class SimArgs:
    # io.cpp:428-526 defaults
    seed: int = -1
    source: int = 0
    depth: float = -1.0          # --depth; math.inf for "inf"
    depths_file: Optional[str] = None
    depths: Optional[List[float]] = None   # parsed --depths-file
    error_rate: float = -1.0
    error_qs: int = 0
    beta_variance: float = -1.0
    gl_model: int = 2
    gl1_theta: float = 0.83
    qs_bins_file: Optional[str] = None
    qs_bins: Optional[List[Sequence[int]]] = None  # parsed --qs-bins: (start, end, value)
    precise_gl: int = 0
    i16_mapq: int = 20
    gvcf_dps: Optional[str] = None
    adjust_qs: int = 0
    adjust_by: float = 0.499
    explode: int = 0
    rm_invar_sites: int = 0
    rm_empty_sites: int = 0
    do_unobserved: int = 1
    do_gvcf: int = 0
    add_gl: int = 1
    add_gp: int = 0
    add_pl: int = 0
    add_i16: int = 0
    add_qs: int = 0
    add_fmt_dp: int = 1
    add_info_dp: int = 0
    add_fmt_ad: int = 0
    add_info_ad: int = 0
    add_fmt_adf: int = 0
    add_info_adf: int = 0
    add_fmt_adr: int = 0
    add_info_adr: int = 0
    # accepted, not on the hot path
    input: Optional[str] = None
    output: Optional[str] = None
    output_mode: str = "b"
    threads: int = 1
    print_pileup: int = 0
    print_truth: int = 0
    other: dict = dataclasses.field(default_factory=dict)

    @property
    def tag_mask(self) -> int:
        m = 0
        for bit, on in ((TAG_GL, self.add_gl), (TAG_GP, self.add_gp), (TAG_PL, self.add_pl),
                        (TAG_I16, self.add_i16), (TAG_QS, self.add_qs),
                        (TAG_FMT_DP, self.add_fmt_dp), (TAG_INFO_DP, self.add_info_dp),
                        (TAG_FMT_AD, self.add_fmt_ad), (TAG_INFO_AD, self.add_info_ad),
                        (TAG_FMT_ADF, self.add_fmt_adf), (TAG_INFO_ADF, self.add_info_adf),
                        (TAG_FMT_ADR, self.add_fmt_adr), (TAG_INFO_ADR, self.add_info_adr)):
            if on:
                m |= bit
        return m


# option name (lower-cased where the reference uses strcasecmp) -> (field, type)
_OPTS = {
    "--seed": ("seed", int), "-s": ("seed", int),
    "--source": ("source", int),
    "--depths-file": ("depths_file", str), "-df": ("depths_file", str),
    "--error-rate": ("error_rate", float), "-e": ("error_rate", float),
    "--error-qs": ("error_qs", int), "-eq": ("error_qs", int),
    "--beta-variance": ("beta_variance", float), "-bv": ("beta_variance", float),
    "--gl-model": ("gl_model", int), "-gl": ("gl_model", int),
    "--gl1-theta": ("gl1_theta", float),
    "--qs-bins": ("qs_bins_file", str),
    "--precise-gl": ("precise_gl", int),
    "--i16-mapq": ("i16_mapq", int),
    "--gvcf-dps": ("gvcf_dps", str),
    "--adjust-qs": ("adjust_qs", int),
    "--adjust-by": ("adjust_by", float),
    "-explode": ("explode", int),
    "--rm-invar-sites": ("rm_invar_sites", int),
    "--rm-empty-sites": ("rm_empty_sites", int),
    "-dounobserved": ("do_unobserved", int),
    "-dogvcf": ("do_gvcf", int),
    "-addgl": ("add_gl", int), "-addformatgl": ("add_gl", int),
    "-addgp": ("add_gp", int), "-addformatgp": ("add_gp", int),
    "-addpl": ("add_pl", int), "-addformatpl": ("add_pl", int),
    "-addi16": ("add_i16", int), "-addformati16": ("add_i16", int),
    "-addqs": ("add_qs", int), "-addformatqs": ("add_qs", int),
    "-addformatdp": ("add_fmt_dp", int), "-addinfodp": ("add_info_dp", int),
    "-addformatad": ("add_fmt_ad", int), "-addinfoad": ("add_info_ad", int),
    "-addformatadf": ("add_fmt_adf", int), "-addinfoadf": ("add_info_adf", int),
    "-addformatadr": ("add_fmt_adr", int), "-addinfoadr": ("add_info_adr", int),
    "--input": ("input", str), "-i": ("input", str),
    "--output": ("output", str), "-o": ("output", str),
    "--output-mode": ("output_mode", str), "-O": ("output_mode", str),
    "--threads": ("threads", int), "-@": ("threads", int),
    "-printpileup": ("print_pileup", int), "-printtruth": ("print_truth", int),
}
_CASE_SENSITIVE = {"--seed", "-s", "--source", "--depths-file", "-df", "--error-rate", "-e",
                   "--input", "-i", "--output", "-o", "--output-mode", "-O", "--threads", "-@",
                   "--depth", "-d"}


def read_qs_bins_file(path: str) -> List[Sequence[int]]:
    """--qs-bins CSV 'start,end,value' (io.cpp:127-220), same checks."""
    bins = []
    with open(path) as fh:
        for ln, line in enumerate(fh, 1):
            line = line.strip()
            if not line:
                continue
            a, b, q = (int(x) for x in line.split(","))
            if a > b or not (0 <= a <= 255) or not (0 <= b <= 255) or not (0 <= q <= 255):
                raise ArgError("bad qs-bins range in line %d of %s" % (ln, path))
            if (not bins and a != 0) or (bins and a != bins[-1][1] + 1):
                raise ArgError("qs-bins ranges must start at 0 and be continuous (line %d)" % ln)
            bins.append((a, b, q))
    if not bins or len(bins) > 255:
        raise ArgError("could not read qs-bins ranges from %s" % path)
    return bins


def read_depths_file(path: str) -> List[float]:
    """--depths-file: one mean depth per sample per line (io.cpp:42-99)."""
    with open(path) as fh:
        return [float(x) for x in fh.read().split()]


def parse_args(argv: Sequence[str], qs_bins=None, depths=None) -> SimArgs:
    """Parse a vcfgl-style argv (without argv[0]); validate like io.cpp:860-1000."""
    a = SimArgs()
    if qs_bins:
        a.qs_bins = [tuple(x) for x in qs_bins]
    if depths:
        a.depths = list(depths)
    i = 0
    argv = list(argv)
    while i < len(argv):
        opt = argv[i]
        if i + 1 >= len(argv):
            raise ArgError("option %s needs a value" % opt)
        val = argv[i + 1]
        i += 2
        if opt in ("--depth", "-d"):
            a.depth = math.inf if val.lower() == "inf" else float(val)
            continue
        key = opt if opt in _CASE_SENSITIVE else opt.lower()
        if key not in _OPTS:
            a.other[opt] = val
            continue
        field, typ = _OPTS[key]
        setattr(a, field, typ(val))
    validate(a)
    if a.qs_bins_file and a.qs_bins is None:
        a.qs_bins = read_qs_bins_file(a.qs_bins_file)
    if a.depths_file and a.depths is None:
        a.depths = read_depths_file(a.depths_file)
    return a


def validate(a: SimArgs) -> None:
    def rng(v, lo, hi, name):
        if not (lo <= v <= hi):
            raise ArgError("%s is out of range [%s, %s]: %s" % (name, lo, hi, v))
    if a.depths_file is None and a.depths is None:
        if a.depth == -1.0:
            raise ArgError("--depth or --depths-file is required")
        if not math.isinf(a.depth):
            rng(a.depth, 0.0, 500.0, "--depth")      # shared.h:63
    if a.error_rate == -1.0:
        raise ArgError("--error-rate is required")
    if not (0.0 <= a.error_rate < 1.0):        # CHECK_ARG_INTERVAL_IE_DBL, io.cpp:868
        raise ArgError("--error-rate is out of range [0, 1): %s" % a.error_rate)
    rng(a.error_qs, 0, 2, "--error-qs")
    rng(a.gl_model, 1, 2, "--gl-model")
    rng(a.gl1_theta, 0.0, 1.0, "--gl1-theta")
    rng(a.precise_gl, 0, 1, "--precise-gl")
    rng(a.i16_mapq, 0, 60, "--i16-mapq")
    rng(a.adjust_qs, 0, 31, "--adjust-qs")
    rng(a.do_unobserved, 0, 5, "-doUnobserved")
    rng(a.rm_invar_sites, 0, 7, "--rm-invar-sites")
    if a.adjust_qs and a.adjust_by == 0.0:
        raise ArgError("--adjust-qs requires non-zero --adjust-by")
    if (a.adjust_qs & 1) and a.precise_gl:
        raise ArgError("--adjust-qs 1 requires --precise-gl 0")
    if (a.adjust_qs & 2) and not a.add_qs:
        raise ArgError("--adjust-qs 2 requires -addQS 1")
    other = {k.lower(): v for k, v in a.other.items()}       # the print flags are accepted, not on the hot path
    for bit, flag, on in ((4, "--printPileup", a.print_pileup), (8, "--printQScores", int(other.get("-printqscores", 0))),
                          (16, "--printGlError", int(other.get("-printglerror", 0)))):
        if (a.adjust_qs & bit) and not on:
            raise ArgError("--adjust-qs %d requires %s 1 (io.cpp:891-899)" % (bit, flag))
    if int(other.get("-printglerror", 0)) and a.gl_model == 1:
        raise ArgError("-printGlError 1 is not supported with --gl-model 1 (io.cpp:993-995)")
    if a.precise_gl and a.gl_model == 1:
        raise ArgError("--precise-gl 1 is not supported with --gl-model 1")
    if a.beta_variance >= 0 and a.error_qs == 0:
        raise ArgError("--beta-variance requires --error-qs 1 or 2")
    if a.error_qs in (1, 2):
        if a.error_rate <= 0.0:
            raise ArgError("--error-qs 1 or 2 requires --error-rate > 0")
        if not a.beta_variance > 0.0:
            raise ArgError("--error-qs 1 or 2 requires --beta-variance > 0")
    if math.isinf(a.depth):       # io.cpp:783-800, 1012-1019
        if a.rm_invar_sites:
            raise ArgError("--rm-invar-sites cannot be used with --depth inf")
        if a.add_fmt_dp:
            raise ArgError("(-addFormatDP 1) FORMAT/DP tag cannot be added when --depth inf is set")
        if a.do_gvcf:
            raise ArgError("[-doGVCF 1] Cannot output gVCF when --depth inf is set")
        if a.error_rate != 0:
            raise ArgError("Cannot simulate true values (--depth inf) with --error-rate %s: set it to 0 (io.cpp:847-853)" % a.error_rate)
        if a.tag_mask & ~(TAG_GL | TAG_GP | TAG_PL):
            raise ArgError("--depth inf: only GL, GP and PL can be added (io.cpp:796-846)")
    if a.do_gvcf:
        if not a.add_fmt_dp or a.rm_invar_sites or a.gvcf_dps is None or not a.add_pl \
                or a.do_unobserved not in (1, 2, 4, 5):
            raise ArgError("-doGVCF 1 requirements not met (io.cpp:958-985)")
    if a.gvcf_dps is not None and not a.do_gvcf:
        raise ArgError("--gvcf-dps requires -doGVCF 1 (io.cpp:986-989)")
    if not a.add_i16 and a.i16_mapq != 20:
        raise ArgError("--i16-mapq requires -addI16 1")


def beta_shape(mean: float, var: float):
    """Beta(alpha, beta) from mean/variance exactly as rng.h:368-371."""
    one_over_mean = 1.0 / mean
    alpha = (((1.0 - mean) / var) - one_over_mean) * pow(mean, 2)
    beta = alpha * (one_over_mean - 1)
    if alpha <= 0.0 or beta <= 0.0:
        raise ArgError("beta shape parameters must be positive; change --error-rate/--beta-variance")
    return alpha, beta

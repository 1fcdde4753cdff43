"""ctypes binding of ``libgenetest_b200.so`` (C ABI: ``include/genetest_b200.h``).

The library owns the CUDA kernels; this module only marshals numpy / torch
buffers into plain pointers.  PyTorch is used for device memory and streams,
nothing else.  There is no CPU fallback: if the shared library is missing or
no sm_100 device is present, creating an :class:`Engine` raises.
"""

import ctypes
import os

import numpy as np

__all__ = ["Engine", "EngineError", "library_path", "load_library",
           "MODELS", "STATUS_OK"]

_HERE = os.path.dirname(os.path.abspath(__file__))

MODELS = {"linear": 0, "logistic": 1, "coxph": 2}

(STATUS_OK, STATUS_ALL_MISSING, STATUS_MAF_BELOW, STATUS_COND_LARGE,
 STATUS_EIG_SMALL, STATUS_NAN_PVALUE, STATUS_PERFECT_SEPARATION,
 STATUS_SINGULAR, STATUS_NO_CONVERGENCE) = range(9)

F_MAF, F_AUX, F_STAT, F_LLF, F_PARAM0 = 0, 1, 2, 3, 4
I_STATUS, I_NOBS, I_FLIP, I_NITER, NUM_IFIELDS = 0, 1, 2, 3, 4

# every symbol include/genetest_b200.h declares
ABI_SYMBOLS = [
    "gt_abi_version", "gt_last_error", "gt_engine_create",
    "gt_engine_destroy", "gt_engine_set_stream", "gt_engine_synchronize",
    "gt_engine_set_option",
    "gt_engine_set_design", "gt_engine_set_sample_sex", "gt_engine_set_file_maf_output",
    "gt_engine_num_params", "gt_engine_num_fields",
    "gt_packed_pitch", "gt_run_packed_dev", "gt_run_packed_host",
    "gt_run_dosage_dev", "gt_run_dosage_host", "gt_synth_packed_dev",
    "gt_engine_enable_timing", "gt_engine_kernel_time",
    "gt_engine_dominant_kernel", "gt_engine_kernel_breakdown", "gt_measure_fp64_peak",
    "gt_measure_dmma_peak", "gt_debug_read_pattern", "gt_format_rows",
    "gt_host_t_sf2", "gt_host_t_ppf975", "gt_host_norm_sf2",
    "gt_host_lt_exp_neg", "gt_host_lt_log1p01",
]


class EngineError(RuntimeError):
    """A ``gt_*`` call returned a negative code."""


class _DesignDesc(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("n", ctypes.c_int32),
        ("k", ctypes.c_int32), ("n_inter", ctypes.c_int32),
        ("n_file", ctypes.c_int32), ("raw_mode", ctypes.c_int32),
        ("C", ctypes.c_void_p), ("y", ctypes.c_void_p),
        ("event", ctypes.c_void_p), ("W", ctypes.c_void_p),
        ("sample_pos", ctypes.c_void_p),
        ("condition_value_t", ctypes.c_double),
        ("eigenvals_t", ctypes.c_double), ("maf_t", ctypes.c_double),
        ("use_maf_t", ctypes.c_int32), ("reserved", ctypes.c_int32),
        ("strata", ctypes.c_void_p),
    ]


def library_path():
    return os.path.join(_HERE, "libgenetest_b200.so")


_LIB = None


def load_library():
    """dlopen the in-tree shared library and declare its signatures."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(
            "{} is missing: build it with `python -c 'import __graft_entry__ "
            "as g; g.build()'` (or `make -C genetest_b200/csrc`). There is no "
            "CPU fallback.".format(path))
    lib = ctypes.CDLL(path)
    c = ctypes
    vp, i32, i64, dbl = c.c_void_p, c.c_int32, c.c_int64, c.c_double
    lib.gt_abi_version.restype = c.c_int
    lib.gt_last_error.restype = c.c_char_p
    lib.gt_engine_create.argtypes = [c.c_int, c.POINTER(vp)]
    lib.gt_engine_destroy.argtypes = [vp]
    lib.gt_engine_set_stream.argtypes = [vp, vp]
    lib.gt_engine_synchronize.argtypes = [vp]
    lib.gt_engine_set_option.argtypes = [vp, c.c_char_p, c.c_int]
    lib.gt_engine_set_design.argtypes = [vp, c.POINTER(_DesignDesc)]
    lib.gt_engine_set_sample_sex.argtypes = [vp, vp]
    lib.gt_engine_set_file_maf_output.argtypes = [vp, vp]
    for name in ("gt_host_t_sf2", "gt_host_t_ppf975", "gt_host_norm_sf2",
                 "gt_host_lt_exp_neg", "gt_host_lt_log1p01"):
        getattr(lib, name).restype = c.c_double
    lib.gt_host_lt_exp_neg.argtypes = [c.c_double]
    lib.gt_host_lt_log1p01.argtypes = [c.c_double]
    lib.gt_host_t_sf2.argtypes = [c.c_double, c.c_double]
    lib.gt_host_t_ppf975.argtypes = [c.c_double]
    lib.gt_host_norm_sf2.argtypes = [c.c_double]
    lib.gt_engine_num_params.argtypes = [vp]
    lib.gt_engine_num_fields.argtypes = [vp]
    lib.gt_packed_pitch.argtypes = [i32]
    lib.gt_packed_pitch.restype = i64
    lib.gt_run_packed_dev.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.gt_run_packed_host.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.gt_run_dosage_dev.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.gt_run_dosage_host.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.gt_synth_packed_dev.argtypes = [vp, vp, i64, i32, i32, c.c_uint64,
                                        dbl, dbl, dbl]
    lib.gt_engine_enable_timing.argtypes = [vp, c.c_int]
    lib.gt_engine_kernel_time.argtypes = [vp, c.POINTER(dbl),
                                          c.POINTER(i64), c.POINTER(i64)]
    lib.gt_engine_dominant_kernel.argtypes = [vp]
    lib.gt_engine_dominant_kernel.restype = c.c_char_p
    lib.gt_engine_kernel_breakdown.argtypes = [vp, c.POINTER(dbl)]
    lib.gt_measure_fp64_peak.argtypes = [vp, c.POINTER(dbl)]
    lib.gt_measure_dmma_peak.argtypes = [vp, c.POINTER(dbl)]
    lib.gt_debug_read_pattern.argtypes = [vp, vp, i64, i32, c.c_int, c.c_int,
                                          c.c_int, c.POINTER(dbl)]
    lib.gt_format_rows.argtypes = [i32, i32, c.POINTER(i32), c.POINTER(vp), vp,
                                   c.c_char, vp, i64]
    lib.gt_format_rows.restype = i64
    for name in ABI_SYMBOLS:
        getattr(lib, name)
    _LIB = lib
    return lib


class Results(object):
    """Host copy of one batch's result blocks (field-major)."""
    def __init__(self, out, iout, n_params):
        self.out = out            # float64 [n_fields, n_snps]
        self.iout = iout          # int32   [4, n_snps]
        self.n_params = n_params

    @property
    def status(self):
        return self.iout[I_STATUS]

    @property
    def nobs(self):
        return self.iout[I_NOBS]

    @property
    def flip(self):
        return self.iout[I_FLIP]

    @property
    def niter(self):
        return self.iout[I_NITER]

    @property
    def maf(self):
        return self.out[F_MAF]

    @property
    def aux(self):
        return self.out[F_AUX]

    @property
    def stat(self):
        return self.out[F_STAT]

    @property
    def llf(self):
        return self.out[F_LLF]

    def param(self, j):
        """(coef, std_err, lower, upper, t, p) arrays of design column j."""
        b = F_PARAM0 + 6 * j
        return self.out[b:b + 6]


class Engine(object):
    """One engine per GPU (``gt_engine``)."""

    def __init__(self, device=0):
        # no torch here: a single-process analysis only needs the C ABI
        # (importing torch costs seconds); gt_engine_create itself refuses
        # to run without an sm_100 device
        self.lib = load_library()
        self.device = int(device)
        self._h = ctypes.c_void_p()
        if self.lib.gt_engine_create(self.device, ctypes.byref(self._h)) < 0:
            raise EngineError(
                "{} -- genetest_b200 has no CPU fallback (needs an sm_100a "
                "GPU)".format(self.lib.gt_last_error().decode()))
        self.n_params = None
        self.n_fields = None
        self.model = None
        self._keep = None

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.lib.gt_last_error().decode())
        return rc

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.gt_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def use_torch_stream(self):
        """Launch on torch's current stream (so torch.cuda.Event times it)."""
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s == 0:
            s = 1       # cudaStreamLegacy: torch's default stream
        self._check(self.lib.gt_engine_set_stream(self._h, ctypes.c_void_p(s)))

    def set_option(self, name, value):
        self._check(self.lib.gt_engine_set_option(self._h, name.encode(),
                                                  int(value)))

    def synchronize(self):
        self._check(self.lib.gt_engine_synchronize(self._h))

    # -- design -----------------------------------------------------------
    def set_design(self, model, y, C, W=None, event=None, sample_pos=None,
                   n_file=None, raw_mode=False, condition_value_t=1000.0,
                   eigenvals_t=1e-10, maf_t=None, strata=None):
        """``y``: (n,), ``C``: (n, k) covariates, ``W``: (n, I) interaction
        multipliers, ``sample_pos``: (n,) genotype-file positions, ``strata``:
        (n,) stratum labels of a Cox model (any values; equal = same stratum)."""
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        n = y.shape[0]
        C = np.asarray(C, dtype=np.float64).reshape(n, -1)
        k = C.shape[1]
        Cf = np.asfortranarray(C)
        I = 0
        Wf = None
        if W is not None and np.size(W):
            Wf = np.asfortranarray(np.asarray(W, dtype=np.float64)
                                   .reshape(n, -1))
            I = Wf.shape[1]
        ev = None
        if event is not None:
            ev = np.ascontiguousarray(event, dtype=np.float64).ravel()
        pos = None
        if sample_pos is not None:
            pos = np.ascontiguousarray(sample_pos, dtype=np.int32)
            if n_file is None:
                raise ValueError("n_file is required with sample_pos")
        d = _DesignDesc()
        d.model = MODELS[model]
        d.n, d.k, d.n_inter = n, k, I
        d.n_file = int(n_file) if n_file is not None else n
        d.raw_mode = int(bool(raw_mode))
        d.C = Cf.ctypes.data if k else None
        d.y = y.ctypes.data
        d.event = ev.ctypes.data if ev is not None else None
        d.W = Wf.ctypes.data if Wf is not None else None
        d.sample_pos = pos.ctypes.data if pos is not None else None
        st = None
        if strata is not None:
            # labels -> dense codes (the engine only compares them)
            st = np.unique(np.asarray(strata), return_inverse=True)[1].astype(np.float64)
            st = np.ascontiguousarray(st.ravel())
            assert st.shape[0] == n
        d.strata = st.ctypes.data if st is not None else None
        d.condition_value_t = float(condition_value_t)
        d.eigenvals_t = float(eigenvals_t)
        d.maf_t = float(maf_t) if maf_t is not None else 0.0
        d.use_maf_t = int(maf_t is not None)
        self._check(self.lib.gt_engine_set_design(self._h, ctypes.byref(d)))
        self.model = model
        self.n = n
        self.n_file = d.n_file
        self.n_params = self._check(self.lib.gt_engine_num_params(self._h))
        self.n_fields = self._check(self.lib.gt_engine_num_fields(self._h))

    def set_sample_sex(self, sex):
        """Sex of every design row (0 female, 1 male, NaN unknown) for the
        sex-aware MAF of ``execute(sexual_chromosome=True)``; None switches
        back to ``nanmean(g) / 2``.  Call after :meth:`set_design`."""
        if sex is None:
            self._check(self.lib.gt_engine_set_sample_sex(self._h, None))
            return
        sex = np.ascontiguousarray(sex, dtype=np.float64)
        if sex.shape != (self.n,):
            raise ValueError("sex: expected {} values, got {}".format(
                self.n, sex.shape))
        self._check(self.lib.gt_engine_set_sample_sex(self._h,
                                                      sex.ctypes.data))

    def packed_pitch(self, n_file=None):
        return int(self.lib.gt_packed_pitch(
            int(self.n_file if n_file is None else n_file)))

    # -- runs -------------------------------------------------------------
    def alloc_results(self, n_snps):
        import torch as t
        dev = "cuda:{}".format(self.device)
        out = t.empty((self.n_fields, n_snps), dtype=t.float64, device=dev)
        iout = t.empty((NUM_IFIELDS, n_snps), dtype=t.int32, device=dev)
        return out, iout

    def run_packed_dev(self, packed, n_snps, pitch=None, out=None, iout=None):
        """``packed``: uint8 CUDA tensor, rows ``pitch`` bytes apart.
        Asynchronous; returns the device result tensors."""
        if out is None:
            out, iout = self.alloc_results(n_snps)
        pitch = packed.stride(0) if pitch is None else pitch
        self._check(self.lib.gt_run_packed_dev(
            self._h, ctypes.c_void_p(packed.data_ptr()), int(pitch),
            int(n_snps), ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(iout.data_ptr())))
        return out, iout

    def run_dosage_dev(self, dosage, out=None, iout=None):
        n_snps = dosage.shape[0]
        if out is None:
            out, iout = self.alloc_results(n_snps)
        self._check(self.lib.gt_run_dosage_dev(
            self._h, ctypes.c_void_p(dosage.data_ptr()),
            int(dosage.stride(0)), int(n_snps),
            ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(iout.data_ptr())))
        return out, iout

    def run_packed_host(self, packed, out=None, iout=None, file_maf=None):
        """``packed``: uint8 ndarray [n_snps, row_bytes] (a .bed body; an
        ``np.memmap`` slice is read in place).  Copies in, runs, copies back,
        synchronises.  ``file_maf``: float64 [n_snps] that receives the MAF of
        every marker over ALL samples of the file (what ``MAFFilter`` tests)."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        n_snps, row_bytes = packed.shape
        if out is None:
            out = np.empty((self.n_fields, n_snps), dtype=np.float64)
            iout = np.empty((NUM_IFIELDS, n_snps), dtype=np.int32)
        if file_maf is not None:
            assert file_maf.dtype == np.float64 and file_maf.shape == (n_snps,)
            self._check(self.lib.gt_engine_set_file_maf_output(
                self._h, file_maf.ctypes.data))
        try:
            self._check(self.lib.gt_run_packed_host(
                self._h, packed.ctypes.data, int(row_bytes), int(n_snps),
                out.ctypes.data, iout.ctypes.data))
        finally:
            if file_maf is not None:
                self.lib.gt_engine_set_file_maf_output(self._h, None)
        return Results(out, iout, self.n_params)

    def run_dosage_host(self, dosage, out=None, iout=None):
        dosage = np.ascontiguousarray(dosage, dtype=np.float64)
        n_snps, ld = dosage.shape
        if out is None:
            out = np.empty((self.n_fields, n_snps), dtype=np.float64)
            iout = np.empty((NUM_IFIELDS, n_snps), dtype=np.int32)
        self._check(self.lib.gt_run_dosage_host(
            self._h, dosage.ctypes.data, int(ld), int(n_snps),
            out.ctypes.data, iout.ctypes.data))
        return Results(out, iout, self.n_params)

    def fetch(self, out, iout):
        """Device result tensors -> :class:`Results` (synchronises)."""
        self.synchronize()
        return Results(out.cpu().numpy(), iout.cpu().numpy(), self.n_params)

    def synth_packed(self, n_file, n_snps, seed=20261019, maf_lo=0.05,
                     maf_hi=0.5, missing_rate=0.0):
        """Synthetic .bed rows generated on the device (bench / tests)."""
        import torch as t
        pitch = self.packed_pitch(n_file)
        buf = t.empty((n_snps, pitch), dtype=t.uint8,
                      device="cuda:{}".format(self.device))
        self._check(self.lib.gt_synth_packed_dev(
            self._h, ctypes.c_void_p(buf.data_ptr()), pitch, int(n_file),
            int(n_snps), int(seed), float(maf_lo), float(maf_hi),
            float(missing_rate)))
        return buf

    # -- timing -----------------------------------------------------------
    def enable_timing(self, on=True):
        self._check(self.lib.gt_engine_enable_timing(self._h, int(on)))

    def kernel_time(self):
        ms = ctypes.c_double()
        n = ctypes.c_int64()
        a = ctypes.c_int64()
        self._check(self.lib.gt_engine_kernel_time(
            self._h, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(a)))
        return ms.value, n.value, a.value

    def fp64_peak_tflops(self):
        v = ctypes.c_double()
        self._check(self.lib.gt_measure_fp64_peak(self._h, ctypes.byref(v)))
        return v.value

    def read_pattern_gbs(self, packed, tpr, vec):
        v = ctypes.c_double()
        self._check(self.lib.gt_debug_read_pattern(
            self._h, ctypes.c_void_p(packed.data_ptr()), packed.stride(0),
            packed.shape[0], tpr, vec, 1, ctypes.byref(v)))
        return v.value

    def dmma_peak_tflops(self):
        v = ctypes.c_double()
        self._check(self.lib.gt_measure_dmma_peak(self._h, ctypes.byref(v)))
        return v.value

    def kernel_breakdown(self):
        ms = (ctypes.c_double * 4)()
        self._check(self.lib.gt_engine_kernel_breakdown(self._h, ms))
        return {"contract": ms[0], "corrections": ms[1], "finish": ms[2]}

    def dominant_kernel(self):
        return self.lib.gt_engine_dominant_kernel(self._h).decode()


_ENGINES = {}


def get_engine(device=0):
    """Process-wide engine of a device (created on first use)."""
    eng = _ENGINES.get(device)
    if eng is None or eng._h is None:
        eng = _ENGINES[device] = Engine(device)
    return eng


def format_rows(columns, keep=None, sep="\t"):
    """Format result rows with the library's native formatter (no GPU needed).

    ``columns``: list of ``list[str]`` (one string per row), float64 ndarray,
    integer ndarray, or a plain ``str`` (repeated on every row).  Floats are
    written like ``repr(float)``.  Returns ``bytes`` (one line per kept row).
    """
    lib = load_library()
    n_rows = None
    types, ptrs, hold = [], [], []
    for col in columns:
        if isinstance(col, str):
            buf = col.encode() + b"\0"
            types.append(3)
        elif isinstance(col, np.ndarray) and col.dtype.kind == "f":
            buf = np.ascontiguousarray(col, dtype=np.float64)
            types.append(1)
            n_rows = buf.shape[0]
        elif isinstance(col, np.ndarray):
            buf = np.ascontiguousarray(col, dtype=np.int64)
            types.append(2)
            n_rows = buf.shape[0]
        else:
            col = list(col)
            n_rows = len(col)
            buf = ("\0".join(col) + "\0").encode()
            types.append(0)
        hold.append(buf)
        if isinstance(buf, bytes):
            ptrs.append(ctypes.cast(ctypes.c_char_p(buf), ctypes.c_void_p).value)
        else:
            ptrs.append(buf.ctypes.data)
    if n_rows is None:
        raise ValueError("format_rows needs at least one per-row column")
    width = 2
    for t, b in zip(types, hold):
        if t in (1, 2):
            width += 26
        elif t == 3:
            width += len(b) + 1
    strbytes = sum(len(b) for t, b in zip(types, hold) if t == 0)
    cap = n_rows * width + strbytes + 64 * 17     # + slack per formatter thread
    out = ctypes.create_string_buffer(cap)
    kp = None
    if keep is not None:
        keep = np.ascontiguousarray(keep, dtype=np.uint8)
        kp = keep.ctypes.data
    n = lib.gt_format_rows(
        n_rows, len(columns), (ctypes.c_int32 * len(types))(*types),
        (ctypes.c_void_p * len(ptrs))(*ptrs), kp, sep.encode(), out, cap)
    if n < 0:
        raise EngineError(lib.gt_last_error().decode())
    return ctypes.string_at(out, n)

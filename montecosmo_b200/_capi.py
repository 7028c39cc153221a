"""ctypes prototypes for include/mcpm.h.  `bind(cdll)` attaches argtypes/restype; `check(lib, code)` raises on error.

Pointer arguments are passed as integers (device addresses from `tensor.data_ptr()`), except the small host arrays
(`scale[3]`, per-step coefficient arrays) which are ctypes float arrays.
"""
import ctypes as C

vp, f32, f64, i32, i64, sz = C.c_void_p, C.c_float, C.c_double, C.c_int, C.c_int64, C.c_size_t
hp = C.POINTER(C.c_float)  # host float array

OBS_SLOTS = 32  # MCPM_OBS_SLOTS


class McpmObs(C.Structure):
    """mcpm_obs of include/mcpm.h (the general observation transform applied inside the paint)."""
    _fields_ = [("curved", i32), ("lightcone", i32), ("ap", i32), ("rsd", i32),
                ("cell", f32 * 3), ("origin", f32 * 3), ("los", f32 * 3),
                ("gf", f32), ("a_par", f32), ("a_perp", f32), ("r0", f32), ("dr", f32), ("nt", i32),
                ("tab_gf", vp), ("tab_ap", vp), ("dvel", vp), ("rot", f32 * 9), ("par", vp)]


MESH = [i32, i32, i32]
XF = [hp, f32]  # scale[3] (host), shift

SIGNATURES = {
    "mcpm_version": ([], i32),
    "mcpm_last_error": ([], C.c_char_p),
    "mcpm_launch_count": ([i32], C.c_longlong),
    "mcpm_engine_create": (MESH + [C.POINTER(vp)], i32),
    "mcpm_engine_destroy": ([vp], i32),
    "mcpm_engine_scratch_bytes": ([vp], sz),
    "mcpm_engine_set_lattice": ([vp, i32, i32, i32], i32),
    "mcpm_engine_set_fused_fft": ([vp, i32], i32),
    "mcpm_engine_set_relative": ([vp, i32], i32),
    "mcpm_tune": ([C.c_char_p, i32], i32),
    "mcpm_engine_tune": ([vp, C.c_char_p, i32], i32),
    "mcpm_paint_brick": ([vp, i32, i32, i32, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_paint3_brick": ([vp, i32, i32, i32, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_paint_lattice": ([vp, vp, vp, vp, f32, i64, vp], i32),
    "mcpm_paint3_lattice": ([vp, vp, vp, vp, vp, f32, f32, i64, vp], i32),
    "mcpm_paint": ([vp, vp, vp, f32, i64] + MESH + [i32] + XF + [vp, i32], i32),
    "mcpm_read": ([vp, vp, vp, i32, i64] + MESH + [i32] + XF + [vp], i32),
    "mcpm_read_grad": ([vp, vp, vp, i32, vp, i64] + MESH + [i32] + XF + [vp, i32], i32),
    "mcpm_paint_vjp": ([vp, vp, vp, f32, vp, i64] + MESH + [i32] + XF + [vp, vp, i32], i32),
    "mcpm_paint_kb": ([vp, vp, vp, f32, i64] + MESH + [i32, f32] + XF + [vp, i32], i32),
    "mcpm_read_kb": ([vp, vp, vp, i32, i64] + MESH + [i32, f32] + XF + [vp], i32),
    "mcpm_read_grad_kb": ([vp, vp, vp, i32, vp, i64] + MESH + [i32, f32] + XF + [vp, i32], i32),
    "mcpm_paint_vjp_kb": ([vp, vp, vp, f32, vp, i64] + MESH + [i32, f32] + XF + [vp, vp, i32], i32),
    "mcpm_deconv_kb": ([vp, vp, vp] + MESH + [i32, f32], i32),
    "mcpm_paint3": ([vp, vp, vp, f32, i64] + MESH + [i32, vp, i32], i32),
    "mcpm_rfftn": ([vp, vp, vp, vp, i32], i32),
    "mcpm_irfftn": ([vp, vp, vp, vp, i32], i32),
    "mcpm_force_spectra": ([vp, vp, vp] + MESH + [i32, i32, f32, i32], i32),
    "mcpm_force_spectra_T": ([vp, vp, vp] + MESH + [i32, i32, f32, i32, i32, i32], i32),
    "mcpm_hessian_spectra": ([vp, vp, vp] + MESH + [i32, i32], i32),
    "mcpm_hessian_spectra_T": ([vp, vp, vp] + MESH + [i32, i32, i32, i32], i32),
    "mcpm_lpt2_source": ([vp, vp, vp, i64], i32),
    "mcpm_lpt2_source_vjp": ([vp, vp, vp, vp, i64], i32),
    "mcpm_deconv": ([vp, vp, vp] + MESH + [i32], i32),
    "mcpm_interlace_combine": ([vp, vp, vp, i32] + MESH + [f32, i32], i32),
    "mcpm_interlace_combine_T": ([vp, vp, vp, i32] + MESH + [f32, i32], i32),
    "mcpm_slabfft_create": (MESH + [i32, C.POINTER(vp)], i32),
    "mcpm_slabfft_destroy": ([vp], i32),
    "mcpm_slabfft_r2c_yz": ([vp, vp, vp, vp, i32], i32),
    "mcpm_slabfft_c2r_yz": ([vp, vp, vp, vp, i32], i32),
    "mcpm_slabfft_c2c_x": ([vp, vp, vp, i32, i32], i32),
    "mcpm_force_spectra_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, f32, i32, f32], i32),
    "mcpm_xfuse_supported": ([i32], i32),
    "mcpm_xfuse_force_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, f32, i32, f32], i32),
    "mcpm_xfuse_peer": ([vp, i32, vp, vp, vp, i32] + MESH + [i32, i32, i32, i32, i32, i32, f32], i32),
    "mcpm_xfuse_force_peer": ([vp, vp, vp, i32, i32] + MESH + [i32, i32, i32, i32, f32, i32, f32], i32),
    "mcpm_xfuse_force_T_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, f32, i32, f32], i32),
    "mcpm_force_spectra_T_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, f32, i32, i32, i32, f32], i32),
    "mcpm_hessian_spectra_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, f32], i32),
    "mcpm_hessian_spectra_T_slab": ([vp, vp, vp] + MESH + [i32, i32, i32, i32, i32, i32, f32], i32),
    "mcpm_interlace_combine_slab": ([vp, vp, vp, i32] + MESH + [i32, i32, f32, i32], i32),
    "mcpm_interlace_combine_T_slab": ([vp, vp, vp, i32] + MESH + [i32, i32, f32, i32, i32, f32], i32),
    "mcpm_chreshape_crop_xz_slab": ([vp, vp] + MESH + [i32, i32, vp, vp, vp, i32, i32, f32], i32),
    "mcpm_chreshape_crop_xz_slab_vjp": ([vp, vp, i32, i32, i32, i32, vp, vp] + MESH + [vp, f32], i32),
    "mcpm_hermitian_project": ([vp, vp] + MESH + [i32], i32),
    "mcpm_half_weight_axpy": ([vp, vp, vp, i64, i32, f32, i32, i32], i32),
    "mcpm_chreshape": ([vp, vp] + MESH + [vp] + MESH, i32),
    "mcpm_chreshape_vjp": ([vp, vp] + MESH + [vp] + MESH, i32),
    "mcpm_spectrum_bins": ([vp, vp, vp] + MESH + [f64, f64, f64, vp, i32, i32, i32, vp], i32),
    "mcpm_spectrum_bins_ell": ([vp, vp, vp] + MESH + [f64, f64, f64, vp, i32, i32, i32, i32, C.POINTER(C.c_double), vp], i32),
    "mcpm_rg2cgh": ([vp, vp, vp] + MESH + [f32, vp], i32),
    "mcpm_rg2cgh_vjp": ([vp, vp, vp] + MESH + [f32, vp], i32),
    "mcpm_cgh2rg": ([vp, vp, vp] + MESH + [f32], i32),
    "mcpm_hermitian_weights": ([vp, vp, vp] + MESH + [i32], i32),
    "mcpm_axpby": ([vp, vp, f32, vp, f32, f32, i64, vp], i32),
    "mcpm_dot": ([vp, vp, vp, i64, vp], i32),
    "mcpm_yz_gradients": ([vp, vp, i32, i32, i32, i32, i32], i32),
    "mcpm_absmax": ([vp, vp, i64, i32, vp], i32),
    "mcpm_rsd_shift": ([vp, vp, vp, hp, f32, i64, vp], i32),
    "mcpm_rsd_shift_vjp": ([vp, vp, hp, f32, i64, vp, i32], i32),
    "mcpm_bias_spectra": ([vp, vp] + MESH + [hp, vp, vp], i32),
    "mcpm_bias_spectra_vjp": ([vp, vp] + MESH + [hp, vp, vp, i32], i32),
    "mcpm_shear_invariants": ([vp, vp, i64, vp], i32),
    "mcpm_shear_invariants_vjp": ([vp, vp, vp, i64, vp], i32),
    "mcpm_bias_moments": ([vp, vp, i32, f32, vp, i64, vp], i32),
    "mcpm_bias_weights": ([vp, vp, i32, f32, vp, hp, vp, i64, vp, vp], i32),
    "mcpm_bias_weights_vjp": ([vp, vp, i32, f32, vp, hp, vp, vp, vp, i64, vp, vp, vp, vp], i32),
    "mcpm_scale_spectrum": ([vp, vp, vp, vp, i64], i32),
    "mcpm_lpt_combine": ([vp, vp, vp, vp, f32, f32, f32, i64, vp, vp, vp], i32),
    "mcpm_kick_drift": ([vp, vp, vp, vp, i64] + MESH + [i32, f32, f32, f32, vp], i32),
    "mcpm_interleave3": ([vp, vp, vp, i64], i32),
    "mcpm_deinterleave3": ([vp, vp, vp, i64], i32),
    "mcpm_kick_drift4": ([vp, vp, vp, vp, i64] + MESH + [f32, f32, f32], i32),
    "mcpm_paint3v4": ([vp, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_read_grad4v": ([vp, vp, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_paint_f": ([vp, vp, vp, vp, f32, i64] + MESH + [i32] + XF + [vp, i32], i32),
    "mcpm_read_f": ([vp, vp, vp, vp, i32, i64] + MESH + [i32] + XF + [vp], i32),
    "mcpm_read_grad_f": ([vp, vp, vp, vp, i32, vp, i64] + MESH + [i32] + XF + [vp, i32], i32),
    "mcpm_paint_vjp_f": ([vp, vp, vp, vp, f32, vp, i64] + MESH + [i32] + XF + [vp, vp, i32], i32),
    "mcpm_paint3_f": ([vp, vp, vp, vp, f32, i64] + MESH + [i32, vp, i32], i32),
    "mcpm_kick_drift_f": ([vp, vp, vp, vp, vp, i64] + MESH + [i32, f32, f32, f32, vp], i32),
    "mcpm_kick_drift4_f": ([vp, vp, vp, vp, vp, i64] + MESH + [f32, f32, f32], i32),
    "mcpm_paint3v4_f": ([vp, vp, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_read_grad4v_f": ([vp, vp, vp, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_read_grad4v_step_f": ([vp, vp, vp, vp, vp, vp, f32, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_paint_brick_f": ([vp, vp, i32, i32, i32, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_paint3_brick_f": ([vp, vp, i32, i32, i32, vp, vp, vp, f32, f32, i64] + MESH + [vp], i32),
    "mcpm_halo_reduce_peer": ([vp, vp, vp, vp, i32, i32, i64, i32, i32], i32),
    "mcpm_halo_gather_peer": ([vp, vp, vp, vp, i32, i32, i64, i32, i32], i32),
    "mcpm_halo_gather4_peer": ([vp, vp, vp, vp, vp, i32, i32, i64, i32], i32),
    "mcpm_drift": ([vp, vp, vp, f32, i64], i32),
    "mcpm_pm_forces": ([vp, vp, vp, i64, i32, i32, i32, i32, f32, vp, vp], i32),
    "mcpm_pm_forces_vjp": ([vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, vp, i32], i32),
    "mcpm_pm_forces_mesh": ([vp, vp, vp, vp, i64, i32, i32, i32, f32, vp], i32),
    "mcpm_pm_forces2": ([vp, vp, vp, vp, i64, i32, i32, i32, vp, vp], i32),
    "mcpm_lpt": ([vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp], i32),
    "mcpm_lpt_vjp": ([vp, vp, vp, i64, i32, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp, vp, vp, vp, i32], i32),
    "mcpm_nbody_steps": ([vp, vp, vp, vp, i64, i32, hp, hp, hp, hp, i32, i32, i32, i32, vp, vp, vp], i32),
    "mcpm_nbody_steps_vjp": ([vp, vp, vp, vp, i64, i32, hp, hp, hp, hp, i32, i32, i32, i32, vp, vp, vp, vp, vp], i32),
    "mcpm_nufft": ([vp, vp, vp, vp, f32, i64, hp, i32, i32, i32, vp], i32),
    "mcpm_nufft_rsd": ([vp, vp, vp, vp, hp, f32, vp, f32, i64, hp, i32, i32, i32, vp], i32),
    "mcpm_nufft_rsd_vjp": ([vp, vp, vp, vp, hp, f32, vp, f32, i64, hp, i32, i32, i32, vp, vp, vp, vp], i32),
    "mcpm_radial_tables": ([vp, vp, i64, vp, i32, vp, vp], i32),
    "mcpm_radial_tables_vjp": ([vp, vp, i64, vp, i32, vp, vp, vp, vp], i32),
    "mcpm_nufft_obs": ([vp, vp, vp, vp, vp, vp, f32, i64, hp, i32, f32, i32, i32, vp], i32),
    "mcpm_nufft_obs_vjp": ([vp, vp, vp, vp, vp, vp, f32, i64, hp, i32, f32, i32, i32, vp, vp, vp, vp, vp, vp], i32),
    "mcpm_nufft_vjp": ([vp, vp, vp, vp, f32, i64, hp, i32, i32, i32, vp, vp, vp], i32),
    "mcpm_nufft_kb": ([vp, vp, vp, vp, f32, i64, hp, i32, f32, i32, i32, vp], i32),
    "mcpm_nufft_vjp_kb": ([vp, vp, vp, vp, f32, i64, hp, i32, f32, i32, i32, vp, vp, vp], i32),
}

class Frame(C.Structure):
    """mcpm_frame (include/mcpm.h): absolute or lattice-relative positions for the stateless particle kernels."""
    _fields_ = [(n, C.c_int) for n in ("relative", "px", "py", "pz", "ox", "oy", "oz", "sx", "sy", "sz")]


def frame(ptcl_shape, span=None, origin=(0, 0, 0)):
    """Relative frame of a lattice `ptcl_shape` spanning `span` mesh cells (default: one cell per site) from `origin`."""
    span = tuple(ptcl_shape) if span is None else tuple(span)
    return Frame(1, *[int(v) for v in ptcl_shape], *[int(v) for v in origin], *[int(v) for v in span])


ERROR_NAMES = {1: "MCPM_EINVAL", 2: "MCPM_ECUDA", 3: "MCPM_ECUFFT", 4: "MCPM_ENOMEM", 5: "MCPM_EUNSUP"}


class McpmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {msg}")
        self.code = code


def bind(lib):
    """Attach prototypes; raises AttributeError if the library misses a symbol the header declares."""
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


def check(lib, code):
    if code != 0:
        raise McpmError(code, lib.mcpm_last_error().decode())


def host_floats(values):
    """Small host float array (scale[3], per-step coefficients)."""
    values = [float(v) for v in values]
    return (C.c_float * len(values))(*values)


def fd_code(fd):
    """np.inf / 2 / 4 -> MCPM_FD_* (invlaplace_hat / gradient_hat, nbody.py:109-163)."""
    if fd in (2, 4):
        return int(fd)
    if fd == float("inf"):
        return 0
    raise ValueError("Only orders 2, 4, and inf are supported.")

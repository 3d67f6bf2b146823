"""ctypes binding of the C-ABI declared in include/roms_b200.h (no torch types cross this boundary)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")


class Config(C.Structure):
    """struct roms_b200_config (include/roms_b200.h)."""
    _fields_ = [
        ("Lm", C.c_int), ("Mm", C.c_int), ("N", C.c_int), ("NT", C.c_int),
        ("NtileI", C.c_int), ("NtileJ", C.c_int), ("tile", C.c_int),
        ("ndtfast", C.c_int),
        ("dt", C.c_double),
        ("nonlin_eos", C.c_int), ("dj_gradps", C.c_int), ("curvgrid", C.c_int), ("mix_geo_ts", C.c_int), ("uv_qdrag", C.c_int),
        ("salinity", C.c_int), ("ana_vmix", C.c_int), ("wvelocity_every_step", C.c_int),
        ("hadv", C.c_int), ("vadv", C.c_int),
        ("rho0", C.c_double), ("g", C.c_double),
        ("R0", C.c_double), ("T0", C.c_double), ("S0", C.c_double), ("Tcoef", C.c_double), ("Scoef", C.c_double),
        ("Akt_bak", C.c_double * 2), ("Akv_bak", C.c_double),
        ("gamma2", C.c_double), ("lambda_", C.c_double), ("hc", C.c_double),
        ("itemp", C.c_int), ("isalt", C.c_int), ("device", C.c_int),
        ("bv_frequency", C.c_int), ("eos_tderivative", C.c_int), ("solar_source", C.c_int), ("lmd_nonlocal", C.c_int),
        ("bulk_fluxes", C.c_int), ("lmd_mixing", C.c_int), ("blk_ZQ", C.c_double), ("blk_ZT", C.c_double), ("blk_ZW", C.c_double),
        ("bvf_mixing", C.c_int), ("nospl_vvisc", C.c_int), ("nospl_vdiff", C.c_int), ("vtransform", C.c_int), ("bodyforce", C.c_int), ("levsfrc", C.c_int), ("levbfrc", C.c_int),
        ("atm_press", C.c_int), ("limit_bstress", C.c_int), ("uv_adv", C.c_int), ("qcorrection", C.c_int), ("limit_stflx_cooling", C.c_int), ("scorrection", C.c_int), ("ts_dif4", C.c_int),
        ("Tnudg_salt", C.c_double),
    ]


class TileArgs(C.Structure):
    """struct roms_b200_tile_t."""
    _fields_ = [("cfg", Config), ("iic", C.c_int), ("ntfirst", C.c_int), ("nstp", C.c_int), ("nnew", C.c_int), ("nrhs", C.c_int),
                ("iif", C.c_int), ("kstp", C.c_int), ("krhs", C.c_int), ("knew", C.c_int), ("predictor", C.c_int)]


PHASES = dict(set_massflux=1, rho_eos=2, set_vbc=3, ana_vmix=4, omega=5, wvelocity=6, set_zeta=7, pre_step3d=8, prsgrd=9,
              t3dmix=10, rhs3d=11, uv3dmix=12, step2d=13, set_depth=14, step3d_uv=15, omega2=16, step3d_t=17, diag=18,
              set_data=19, step2d_loop=20, set_avg=22, bulk_flux=23, lmd_vmix=24, bvf_mix=25, t3dmix4=26)
INDEX_NAMES = ["iic", "ntstart", "ntfirst", "nstp", "nnew", "nrhs", "iif", "indx1", "kstp", "krhs", "knew", "PREDICTOR", "exit_flag"]
DIAG_NAMES = ["avgke", "avgpe", "avgkp", "volume", "max_speed", "maxCu", "maxCv", "maxCw", "ubarmax", "vbarmax", "umax", "vmax"]

# every symbol include/roms_b200.h declares
EXPORTS = ["roms_b200_default_config", "roms_b200_bounds", "roms_b200_bounds_names", "roms_b200_create", "roms_b200_destroy",
           "roms_b200_array_bounds", "roms_b200_set_field", "roms_b200_get_field", "roms_b200_set_scoord", "roms_b200_set_weights",
           "roms_b200_set_indices", "roms_b200_get_indices", "roms_b200_run_phase", "roms_b200_main3d_step", "roms_b200_sync",
           "roms_b200_step_forced", "roms_b200_step_fields", "roms_b200_diag", "roms_b200_register_host", "roms_b200_unregister_host", "roms_b200_last_step_ms", "roms_b200_profile_enable", "roms_b200_profile_get",
           "roms_b200_launch_count", "roms_b200_attach_nccl", "roms_b200_nccl_unique_id", "roms_b200_nccl_init_rank",
           "roms_b200_peer_export", "roms_b200_peer_attach", "roms_b200_peer_enable", "roms_b200_peer_error", "roms_b200_peer_error_inject",
           "roms_b200_set_option", "roms_b200_set_avg",
           "roms_b200_rho_eos_tile", "roms_b200_prsgrd_tile", "roms_b200_set_massflux_tile", "roms_b200_omega_tile",
           "roms_b200_set_depth_tile", "roms_b200_routine_tile", "roms_b200_routine_args", "roms_b200_field_levels"]

_cache = {}
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


def lib_path(strict=False):
    # ROMS_B200_LIB: kernel-tuning aid (tools/variant.py builds alternative production libraries); never set by tests
    if not strict and os.environ.get("ROMS_B200_LIB"):
        return os.environ["ROMS_B200_LIB"]
    return os.path.join(LIBDIR, "libroms_b200_strict.so" if strict else "libroms_b200.so")


def load(strict=False):
    """Load the CUDA library.  Fails loudly when it has not been built: there is no CPU fallback."""
    key = bool(strict)
    if key in _cache:
        return _cache[key]
    path = lib_path(strict)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m roms_trunk_mgh_b200.build` (nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    H = C.c_void_p
    L.roms_b200_default_config.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Config)]
    L.roms_b200_bounds.argtypes = [C.c_int] * 6 + [IP]
    L.roms_b200_bounds_names.restype = C.c_char_p
    L.roms_b200_create.argtypes = [C.POINTER(Config), C.POINTER(H)]
    L.roms_b200_destroy.argtypes = [H]
    L.roms_b200_array_bounds.argtypes = [H, IP]
    L.roms_b200_set_field.argtypes = [H, C.c_char_p, DP, C.c_size_t]
    L.roms_b200_get_field.argtypes = [H, C.c_char_p, DP, C.c_size_t]
    L.roms_b200_set_scoord.argtypes = [H, C.c_int, DP, C.c_int]
    L.roms_b200_set_weights.argtypes = [H, C.c_int, DP, DP, C.c_int]
    L.roms_b200_set_indices.argtypes = [H, IP, DP]
    L.roms_b200_get_indices.argtypes = [H, IP, DP]
    L.roms_b200_run_phase.argtypes = [H, C.c_int]
    L.roms_b200_main3d_step.argtypes = [H, C.c_int]
    L.roms_b200_sync.argtypes = [H]
    L.roms_b200_step_forced.argtypes = [H, DP, DP, DP, C.c_size_t, DP]
    L.roms_b200_step_fields.argtypes = [H, C.c_int, C.POINTER(C.c_char_p), C.POINTER(DP), C.c_size_t, DP]
    L.roms_b200_diag.argtypes = [H, DP]
    L.roms_b200_register_host.argtypes = [H, C.c_void_p, C.c_size_t]
    L.roms_b200_unregister_host.argtypes = [H, C.c_void_p]
    L.roms_b200_last_step_ms.argtypes = [H, C.POINTER(C.c_float)]
    L.roms_b200_profile_enable.argtypes = [H, C.c_int]
    L.roms_b200_profile_get.argtypes = [H, DP, C.POINTER(C.c_longlong)]
    L.roms_b200_launch_count.argtypes = [H]
    L.roms_b200_launch_count.restype = C.c_longlong
    L.roms_b200_attach_nccl.argtypes = [H, C.c_void_p, C.c_int, C.c_int]
    L.roms_b200_nccl_unique_id.argtypes = [C.c_char_p]
    L.roms_b200_nccl_init_rank.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.roms_b200_peer_export.argtypes = [H, C.c_char_p]
    L.roms_b200_peer_attach.argtypes = [H, C.c_char_p, C.c_char_p]
    L.roms_b200_peer_enable.argtypes = [H, C.c_int]
    L.roms_b200_peer_error.argtypes = [H]
    L.roms_b200_peer_error_inject.argtypes = [H]
    L.roms_b200_set_option.argtypes = [H, C.c_char_p, C.c_double]
    L.roms_b200_set_avg.argtypes = [H, C.c_int, C.c_int]
    TP = C.POINTER(TileArgs)
    L.roms_b200_rho_eos_tile.argtypes = [TP] + [DP] * 9
    L.roms_b200_prsgrd_tile.argtypes = [TP] + [DP] * 8
    L.roms_b200_set_massflux_tile.argtypes = [TP] + [DP] * 7
    L.roms_b200_omega_tile.argtypes = [TP] + [DP] * 4
    L.roms_b200_set_depth_tile.argtypes = [TP] + [DP] * 9
    L.roms_b200_routine_tile.argtypes = [TP, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(DP), IP, DP, C.c_int, DP, DP, C.c_int]
    L.roms_b200_routine_args.argtypes = [C.c_int]
    L.roms_b200_routine_args.restype = C.c_char_p
    L.roms_b200_field_levels.argtypes = [H, C.c_char_p, IP, IP]
    _cache[key] = L
    return L


def default_config(app, Lm=0, Mm=0, N=0, strict=False):
    cfg = Config()
    rc = load(strict).roms_b200_default_config(app, Lm, Mm, N, C.byref(cfg))
    if rc:
        raise ValueError(f"roms_b200_default_config -> {rc}")
    return cfg


def bounds(Lm, Mm, NtileI, NtileJ, tile, distribute=False, strict=False):
    L = load(strict)
    out = (C.c_int * 57)()
    rc = L.roms_b200_bounds(Lm, Mm, NtileI, NtileJ, tile, int(distribute), out)
    if rc:
        raise ValueError(f"roms_b200_bounds -> {rc}")
    names = L.roms_b200_bounds_names().decode().split(",")
    return dict(zip(names, list(out)))

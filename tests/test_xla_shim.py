"""
The XLA-FFI shim (montecosmo_b200/xla/mcpm_xla.cc) and its JAX module (montecosmo_b200/jax_nbody.py) cannot be built or
imported in this image (no jax / jaxlib headers).  What can be checked without them, so that the uncompiled
source cannot drift from the ABI silently:
  * every `mcpm_*(...)` call in the shim names a symbol of include/mcpm.h with exactly as many arguments as its ctypes
    prototype in montecosmo_b200/_capi.py;
  * every handler the JAX module registers is defined in the shim (XLA_FFI_DEFINE_HANDLER_SYMBOL), and every attribute
    name the module passes to a handler is bound by that handler (`.Attr<...>("name")`, macros expanded);
  * every composite entry point INTEGRATION.md's symbol map lists is reachable from a handler.
"""
import ast
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "montecosmo_b200", "xla", "mcpm_xla.cc")
JAXMOD = os.path.join(ROOT, "montecosmo_b200", "jax_nbody.py")


def _strip_comments(src):
    src = re.sub(r"//[^\n]*", "", src)
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def _calls(src):
    """(name, n_args) for every mcpm_xxx( ... ) call: top-level commas of the balanced argument list."""
    out = []
    for m in re.finditer(r"\b(mcpm_[a-z0-9_]+)\s*\(", src):
        i, depth, commas, empty = m.end(), 1, 0, True
        while depth:
            c = src[i]
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            elif c == "," and depth == 1:
                commas += 1
            if depth and not c.isspace():
                empty = False
            i += 1
        out.append((m.group(1), 0 if empty else commas + 1))
    return out


def test_shim_calls_match_the_abi_prototypes():
    from montecosmo_b200._capi import SIGNATURES
    src = _strip_comments(open(SHIM).read())
    calls = [c for c in _calls(src) if c[0] != "mcpm_last_error"]
    assert len(calls) >= 30
    for name, nargs in calls:
        assert name in SIGNATURES, f"{name} is not declared in include/mcpm.h / _capi.py"
        assert nargs == len(SIGNATURES[name][0]), f"{name}: shim passes {nargs} arguments, the ABI takes {len(SIGNATURES[name][0])}"
    used = {c[0] for c in calls}
    for must in ("mcpm_paint", "mcpm_paint_vjp", "mcpm_read", "mcpm_read_grad", "mcpm_nufft", "mcpm_nufft_vjp",
                 "mcpm_pm_forces", "mcpm_pm_forces_vjp", "mcpm_pm_forces_mesh", "mcpm_pm_forces2", "mcpm_lpt",
                 "mcpm_lpt_vjp", "mcpm_nbody_steps", "mcpm_nbody_steps_vjp", "mcpm_chreshape", "mcpm_chreshape_vjp",
                 "mcpm_deconv", "mcpm_rfftn", "mcpm_irfftn", "mcpm_rg2cgh", "mcpm_rg2cgh_vjp", "mcpm_cgh2rg",
                 "mcpm_spectrum_bins_ell", "mcpm_paint_kb", "mcpm_nufft_kb", "mcpm_engine_set_relative"):
        assert must in used, f"the shim never calls {must}"


def _expand_macros(src):
    macros = dict(re.findall(r"#define\s+(\w+)\s+((?:[^\n\\]|\\\n)*)", src))
    macros = {k: v.replace("\\\n", " ") for k, v in macros.items() if k in ("XF", "LATTICE", "FD", "STEPS", "OBS")}
    for _ in range(3):
        for k, v in macros.items():
            src = re.sub(rf"\b{k}\b", v, src)
    return src


def test_jax_module_and_shim_agree_on_handlers_and_attributes():
    src = _expand_macros(_strip_comments(open(SHIM).read()))
    defs = {}
    for m in re.finditer(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),\s*(\w+),(.*?)\);", src, flags=re.S):
        defs[m.group(1)] = set(re.findall(r'\.Attr<[^>]+(?:>)?>\("(\w+)"\)', m.group(3)))
    tree = ast.parse(open(JAXMOD).read())
    handlers = None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "HANDLERS":
            handlers = ast.literal_eval(node.value)
    assert handlers and len(handlers) >= 25
    for target, sym in handlers.items():
        assert sym in defs, f"{sym} ({target}) is not defined in mcpm_xla.cc"
    # every ffi_call("target", ...)(..., attr=...) passes only attributes the handler binds
    steps_attrs = {"mesh", "alpha", "beta", "drift_pre", "drift_post", "order", "paint_deconv", "lap_fd", "grad_fd", "lattice",
                   "relative"}
    obs_attrs = {"flags", "geom", "rot", "wscalar", "scale", "paint_order", "kcut", "interlace_order", "paint_deconv",
                 "lattice", "relative"}
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "_obs_attrs")
    ret = next(n for n in ast.walk(fn) if isinstance(n, ast.Return))
    assert {k.arg for k in ret.value.keywords} == obs_attrs  # what **attrs of the observed nufft expands to
    checked = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Call):
            inner = node.func
            if getattr(inner.func, "attr", "") == "ffi_call" and inner.args and isinstance(inner.args[0], ast.Constant):
                target = inner.args[0].value
                bound = defs[handlers[target]]
                passed = {k.arg for k in node.keywords if k.arg is not None}
                if any(k.arg is None for k in node.keywords):  # **_steps_attrs(...) / **attrs of the observed nufft
                    passed |= obs_attrs if target.startswith("mcpm_nufft_obs") else steps_attrs
                assert passed == bound, f"{target}: module passes {sorted(passed)}, handler binds {sorted(bound)}"
                checked += 1
    assert checked >= 15

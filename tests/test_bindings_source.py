"""The jax.ffi shim (bindings/) is shipped as source because jax / jaxlib cannot be installed in this image.  What can
be checked without them: every C-ABI function the shim calls is declared in include/pdeopt_b200.h and exported by
the built library, and every FFI target the Python wrapper registers is defined by the shim."""
import os
import re
import subprocess

from pde_opt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    return open(os.path.join(ROOT, *p)).read()


def test_shim_calls_only_declared_and_exported_entry_points():
    shim, header = _read("bindings", "pdeopt_jax_ffi.cc"), _read("include", "pdeopt_b200.h")
    called = set(re.findall(r"\b(pdeopt_[a-z0-9_]+)\s*\(", shim))
    declared = set(re.findall(r"\b(pdeopt_[a-z0-9_]+)\s*\(", header))
    assert called and called <= declared, called - declared
    lib = _lib.load()
    for name in called:
        getattr(lib, name)
    assert called <= set(_lib.EXPORTS), called - set(_lib.EXPORTS)


def test_python_wrapper_targets_exist_in_shim():
    shim, py = _read("bindings", "pdeopt_jax_ffi.cc"), _read("bindings", "pdeopt_jax.py")
    defined = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", shim))
    registered = set(re.findall(r'"pdeopt_\w+":\s*"(\w+)"', py))
    assert registered and registered == defined, registered ^ defined
    used = set(re.findall(r'ffi_call\("(pdeopt_\w+)"', py))
    names = set(re.findall(r'"(pdeopt_\w+)":\s*"\w+"', py))
    assert used <= names, used - names


def test_argument_counts_match_the_header():
    """Each call in the shim passes as many arguments as the prototype in the header declares."""
    shim, header = _read("bindings", "pdeopt_jax_ffi.cc"), _read("include", "pdeopt_b200.h")

    def arg_count(text, start):
        depth, n, i = 0, 1, start
        while True:
            c = text[i]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
                if depth == 0:
                    return n
            elif c == "," and depth == 1:
                n += 1
            i += 1

    protos = {m.group(1): arg_count(header, m.end() - 1) for m in re.finditer(r"pdeopt_status\s+(pdeopt_[a-z0-9_]+)\s*\(", header)}
    for m in re.finditer(r"Status\((pdeopt_[a-z0-9_]+)\s*\(", shim):
        assert arg_count(shim, m.end() - 1) == protos[m.group(1)], m.group(1)


def test_shim_type_checks_against_the_header_with_a_mock_of_the_xla_ffi_api():
    """g++ -fsyntax-only of bindings/pdeopt_jax_ffi.cc against include/pdeopt_b200.h and a MOCK xla/ffi/api/ffi.h
    (tests/host/mock_xla: buffers, results, spans, errors, an unchecked binder): every call into the C ABI has the
    right argument count and types.  The real jaxlib header is what a maintainer builds against."""
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(ROOT, "tests", "host", "mock_xla"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "bindings", "pdeopt_jax_ffi.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""C-ABI contract (no GPU needed): libevcplm.so loads, exports every function include/evcplm.h declares, and the
ctypes binding (evcouplings_b200/_lib.py) covers exactly that set.  No compute entry point is called."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "evcplm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)          # drop comments
    return set(re.findall(r"\b(evc_[a-z0-9_]+)\s*\(", text))


def test_header_binding_and_library_agree():
    from evcouplings_b200 import _lib
    declared = _declared()
    assert len(declared) >= 25
    assert declared == set(_lib.PROTOTYPES), (declared ^ set(_lib.PROTOTYPES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libevcplm.so does not export " + name
    bound = _lib.load()
    assert bound.evc_abi_version() == _lib.ABI_VERSION == 1


def test_library_reports_errors_without_device():
    """argument validation and the error channel work without touching a GPU"""
    from evcouplings_b200 import _lib
    lib = _lib.load()
    assert lib.evc_plm_num_params(None) == -1
    assert lib.evc_hamming_num_tiles(300) == 6 and lib.evc_hamming_plane_words(300, 40) == 5 * 2 * 300
    rc = lib.evc_plm_set_backward(None, 1)
    assert rc != 0 and b"null handle" in lib.evc_last_error()
    rc = lib.evc_hamming_counts(None, 0, 0, 0, 0, None)
    assert rc != 0 and lib.evc_last_error()

"""pytest plugin (-p emu_plugin, PYTHONPATH=tests/emu): points the ctypes mirror at the emulated
build of the whole library (KML_EMU_LIB = libkml_emu.so from build_libkml_emu.py), so that the
`-m gpu` parity tests run unchanged on a host without a GPU: every kernel under the SIMT emulator,
the host code of the C ABI as it is.  Only ever loaded by tests/test_emulated_library.py, in a
subprocess of its own; the product loader (kml/_lib.py) has no such switch."""
import ctypes as C
import os


def pytest_configure(config):
    import kml._lib as L
    lib = C.CDLL(os.environ["KML_EMU_LIB"])
    lib.kml_last_error.restype = C.c_char_p
    lib.kml_last_error.argtypes = [C.c_void_p]
    L._lib = lib
    L.LIB_PATH = os.environ["KML_EMU_LIB"]

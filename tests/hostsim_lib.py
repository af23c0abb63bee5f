"""TEST INFRASTRUCTURE: runs the C ABI against tests/hostsim/_build/libljsim.so (the csrc/*.cu sources
compiled with g++ over a serial CUDA stand-in) so host orchestration and device math can be tested on
the GPU-less box.  The product package never loads this library."""
import contextlib
import ctypes as C
import os
import subprocess

from lajolla_public_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
SIM_SO = os.path.join(HERE, "hostsim", "_build", "libljsim.so")
_sim = None


def build():
    subprocess.run(["make", "-C", os.path.join(HERE, "hostsim"), "-j4"], check=True, stdout=subprocess.DEVNULL)


def sim_lib():
    global _sim
    if _sim is None:
        build()  # (make is incremental: a stale library would silently test old code)
        lib = C.CDLL(SIM_SO)
        for name, (res, args) in abi.PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _sim = lib
    return _sim


@contextlib.contextmanager
def simulated():
    """Within this context lajolla_public_b200 objects are created against the host simulation."""
    saved = abi._lib
    abi._lib = sim_lib()
    try:
        yield
    finally:
        abi._lib = saved

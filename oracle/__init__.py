"""CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header; parity unpinned).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import
this package. ``oracle_api()`` binds oracle/liboracle.so through the same ctypes class the product
uses for the CUDA library (the two export the same C ABI)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
SRC = os.path.join(_HERE, "oracle.cpp")


def build_oracle(force: bool = False) -> str:
    hdr = os.path.join(_HERE, "..", "include", "schic_mh.h")
    if (force or not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(SRC), os.path.getmtime(hdr))):
        # $CXX first, then the system g++ (an exotic $CXX may lack libgomp); last resort: no OpenMP (slower, same results)
        attempts, errors = [], []
        for cxx in dict.fromkeys([os.environ.get("CXX", "g++"), "g++"]):
            attempts += [[cxx, "-fopenmp"], ]
        attempts.append(["g++"])
        for head in attempts:
            cmd = [head[0], "-O3", "-std=c++17", *head[1:], "-shared", "-fPIC", "-o", LIB_PATH, SRC]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                break
            errors.append(" ".join(cmd) + "\n" + r.stderr)
        else:
            raise RuntimeError("oracle build failed:\n" + "\n".join(errors))
    return LIB_PATH


_api = None


def oracle_api():
    global _api
    if _api is None:
        from schicexplorer_b200._capi import CApi
        if not os.path.exists(LIB_PATH):
            build_oracle()
        _api = CApi(LIB_PATH)
        assert _api.backend == "oracle"
    return _api

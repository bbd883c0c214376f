"""The C-ABI library must load on a CPU-only box and export every symbol include/ipc_b200.h declares.
No compute entry point is called here (there is no CPU fallback: they fail with IPC_ERR_CUDA)."""
import ctypes
import os
import re

import numpy as np

from ipc_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ipc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = api.lib()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ipc_b200.h but not exported"
    assert sorted(api.SYMBOLS) == names


def test_struct_layouts_match_header():
    assert ctypes.sizeof(api.Config) == 32
    assert ctypes.sizeof(api.CheckInfo) == 40 == api.INFO_DTYPE.itemsize


def test_no_cpu_fallback():
    """Without a device, ipc_create refuses (IPC_ERR_CUDA) instead of computing on the host."""
    L = api.lib()
    if L.ipc_device_count() > 0:
        return
    om = np.zeros((3, 3)); oi = np.tile(np.eye(3), (3, 1, 1))
    try:
        api.IPC(2, om, oi, dict(s_factor=1.0, fast_reject_th=1.0, slow_reject_th=1.0, fast_reject_iter_base=1, slow_reject_iter_base=1))
    except api.IpcError as e:
        assert "-2" in str(e)
    else:
        raise AssertionError("ipc_create succeeded without a CUDA device")


def test_argument_validation_without_device():
    L = api.lib()
    h = ctypes.c_void_p()
    assert L.ipc_create(5, 10, None, None, None, 0, ctypes.byref(h)) == -1      # IPC_ERR_ARG
    assert b"dim" in L.ipc_last_error()

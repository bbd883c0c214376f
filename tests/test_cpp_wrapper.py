"""include/ipc_b200.hpp — the header-only C++ class with the shape of the reference's IPC<EDGE, VERTEX>
(/root/reference/include/ipc/consensus.hpp:5-33): both instantiations compile against the C ABI and link with libipc_b200.so,
argument errors are raised as ipc_b200::Error, and without a CUDA device the constructor fails loudly (no CPU path)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_wrapper_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = str(tmp_path / "wrapper_check")
    lib_dir = os.path.join(ROOT, "ipc_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "wrapper_check.cpp"), "-L" + lib_dir, "-lipc_b200", "-Wl,-rpath," + lib_dir, "-o", exe])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and "WRAPPER_OK" in p.stdout, p.stdout + p.stderr


def test_wrapper_has_the_reference_members():
    """Same member names as include/ipc/consensus.hpp:9-16 of the reference."""
    src = open(os.path.join(ROOT, "include", "ipc_b200.hpp")).read()
    for name in ("agreementCheck", "removeEdgeFromCnS", "addEdgeToCnS", "getMaxConsensusSet"):
        assert re.search(r"\b" + name + r"\s*\(", src), name
    # and every C entry point it names is declared by the C header
    hdr = open(os.path.join(ROOT, "include", "ipc_b200.h")).read()
    names = set(re.findall(r"\bipc_[a-z_]+\b", src)) - {"ipc_b200", "ipc_handle", "ipc_config", "ipc_check_info"}
    assert len(names) >= 15
    for fn in sorted(names):
        assert re.search(r"\b" + fn + r"\s*\(", hdr), fn

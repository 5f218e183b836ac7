"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol that
include/ddw.h declares (and the ctypes table binds exactly those), and rejects bad arguments before
touching the device.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from deepdow_b200 import _C, build
    if not os.path.exists(_C.LIB_PATH):
        build.build()
    return _C.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ddw.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ddw.h but not exported"


def test_ctypes_table_matches_header():
    from deepdow_b200 import _C
    assert sorted(_C.SIGNATURES) == declared_symbols()


def test_version_and_error_channel(lib):
    assert lib.ddw_version() == 102
    assert isinstance(lib.ddw_last_error(), bytes)


def test_argument_validation_without_gpu(lib):
    # bad dtype / sizes are rejected on the host, before any CUDA call
    rc = lib.ddw_markowitz_fwd(None, None, None, None, 1.0, 4, 5, 7, None, None, None, None, 0, None)
    assert rc == -1 and b"dtype" in lib.ddw_last_error()
    rc = lib.ddw_markowitz_fwd(None, None, None, None, 1.0, 4, 5000, 0, None, None, None, None, 0, None)
    assert rc == -2 and b"n_assets" in lib.ddw_last_error()
    rc = lib.ddw_markowitz_fwd(None, None, None, None, 1.0, 4, 5, 0, None, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.ddw_last_error()
    rc = lib.ddw_capped_simplex_fwd(None, None, 1.0, 1.0, 4, 5, 0, ctypes.c_void_p(8), None)
    assert rc == -1
    rc = lib.ddw_covariance_fwd(ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 0, 2, 1, 3, 0, ctypes.c_void_p(8), None, None, None, None, 0, None)
    assert rc == -1 and b"observations" in lib.ddw_last_error()
    rc = lib.ddw_covariance_fwd(ctypes.c_void_p(8), ctypes.c_void_p(8), 9, 0, 2, 4, 3, 0, ctypes.c_void_p(8), None, None, None, None, 0, None)
    assert rc == -1 and b"strategy" in lib.ddw_last_error()
    # empty batch is a no-op
    assert lib.ddw_markowitz_fwd(None, None, None, None, 1.0, 0, 5, 0, None, None, None, None, 0, None) == 0


def test_workspace_queries(lib):
    assert lib.ddw_markowitz_workspace_bytes(4096, 100, 0) > 4096 * 4
    assert lib.ddw_markowitz_workspace_bytes(0, 100, 0) == 0
    # N = 500 does not fit shared memory: the workspace holds Q per problem
    assert lib.ddw_markowitz_workspace_bytes(8, 500, 0) >= 8 * 500 * 500 * 4
    assert lib.ddw_covariance_workspace_bytes(8, 40, 50, 0, 1) == 0


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under deepdow_b200/ may reference it."""
    pkg = os.path.join(ROOT, "deepdow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ddw_oracle" not in text, f


def test_chunked_and_host_entry_points_validate_without_gpu(lib):
    # chunk outside the batch
    rc = lib.ddw_markowitz_fwd_chunk(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1.0,
                                     8, 5, 0, 10, 4, ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), None, 0, None)
    assert rc == -1 and b"outside the batch" in lib.ddw_last_error()
    rc = lib.ddw_markowitz_bwd_chunk(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16),
                                     ctypes.c_void_p(16), ctypes.c_void_p(16), 1.0, 8, 5, 0, 10, -1, ctypes.c_void_p(16),
                                     ctypes.c_void_p(16), None, 0, ctypes.c_void_p(16), None, 0, None)
    assert rc == -1 and b"outside the batch" in lib.ddw_last_error()
    # host step: argument errors are reported before any CUDA call
    z = ctypes.c_void_p(64)
    assert lib.ddw_markowitz_host_step(None, z, z, z, z, None, 1.0, 4, 5, 9, 2, z, z, None, None, None, None, z, 0, None) == -1
    assert b"dtype" in lib.ddw_last_error()
    assert lib.ddw_markowitz_host_step(None, z, z, z, z, None, 1.0, 4, 5000, 0, 2, z, z, None, None, None, None, z, 0, None) == -2
    assert lib.ddw_markowitz_host_step(None, z, z, z, z, None, 1.0, 4, 5, 0, 2, z, z, None, None, None, None, z, 0, None) == -1
    assert b"null" in lib.ddw_last_error()                      # no pipeline handle
    assert lib.ddw_markowitz_host_step(None, z, z, z, z, None, 1.0, 0, 5, 0, 2, z, z, None, None, None, None, z, 0, None) == 0   # empty batch
    assert lib.ddw_markowitz_host_scratch_bytes(4096, 100, 0, 128) > 3 * 128 * 100 * 100 * 4
    assert lib.ddw_markowitz_host_scratch_bytes(0, 100, 0, 128) == 0
    assert lib.ddw_host_pipeline_graph_launches(None) == 0
    lib.ddw_host_pipeline_destroy(None)                         # no-op


def test_hot_kernels_stage_their_matrix_with_bulk_tma(lib):
    """SASS evidence (B200_PROFILING.md): the sparse-support forward and backward kernels bring covmat_sqrt into shared
    memory with cp.async.bulk, which shows up as UBLKCP + SYNCS (mbarrier) in the sm_100a code of the built library."""
    import shutil
    import subprocess
    from deepdow_b200 import _C
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _C.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    seen = {}
    current = None
    for line in out.splitlines():
        if "Function :" in line:
            current = line.split("Function :")[1].strip()
        elif "UBLKCP" in line and current:
            seen[current] = seen.get(current, 0) + 1
    assert any("markowitz_fwd_sparse_kernelIf" in k for k in seen), sorted(seen)
    assert any("markowitz_bwd_sparse_kernelIf" in k for k in seen), sorted(seen)
    assert any("markowitz_gram_small_kernel" in k for k in seen), sorted(seen)      # A by one bulk copy per portfolio
    assert any("markowitz_qwarp_kernel" in k for k in seen), sorted(seen)           # rows of Q by one bulk copy per asset
    assert "arch = sm_100a" in out


def test_headline_forward_runs_its_gram_on_tcgen05(lib):
    """SASS evidence for the north star's "tensor cores for the X'X contraction": the float32 NumericalMarkowitz forward forms
    Q = A'A with tcgen05.mma (UTCHMMA) into TMEM accumulators read back by tcgen05.ld (LDTM), committed to mbarriers (UTCBAR)."""
    import shutil
    import subprocess
    from deepdow_b200 import _C
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", _C.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    body = out.split("markowitz_gram_small_kernel", 1)[1].split("Function :", 1)[0]
    for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP"):
        assert mnemonic in body, mnemonic

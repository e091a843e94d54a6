"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/pbx_gemm.h declares, reports the reference's exception texts, and fails loudly without a
GPU (no compute call is made here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(pbx_lib):
    header = (ROOT / "include" / "pbx_gemm.h").read_text()
    # strip comments, then collect function declarators
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pbx_[a-z0-9_]+)\s*\(", code))
    assert len(declared) >= 25
    from portblas_b200 import _lib
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(pbx_lib, name), f"libpbx_gemm.so does not export {name}"


def test_status_strings_match_reference_exceptions(pbx_lib):
    # reference src/interface/gemm_interface.hpp:144-165
    want = {1: "invalid _TransA", 2: "invalid _TransB", 3: "invalid _stridec", 4: "invalid _stridea",
            5: "invalid _strideb"}
    for code, text in want.items():
        assert pbx_lib.pbx_status_string(code).decode() == text


def test_no_gpu_fails_loudly(pbx_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    st = pbx_lib.pbx_create(ctypes.byref(h), 0, None)
    assert st == 8 and not h.value  # PBX_ERR_NO_DEVICE: there is no CPU fallback behind the ABI
    from portblas_b200 import SB_Handle, PbxError
    with pytest.raises(PbxError):
        SB_Handle(0)


def test_null_handle_is_rejected(pbx_lib):
    assert pbx_lib.pbx_destroy(None) == 6
    assert pbx_lib.pbx_synchronize(None) == 6
    assert pbx_lib.pbx_get_num_compute_units(None) == 0


def test_product_does_not_import_oracle():
    """The product path may never route through the oracle (or any CPU fallback)."""
    for p in list((ROOT / "portblas_b200").rglob("*.py")) + list((ROOT / "portblas_b200" / "csrc").glob("*.cu*")) + \
            list((ROOT / "include").rglob("*.h*")):
        txt = p.read_text()
        assert "oracle" not in txt.lower() or p.name == "__init__.py" and False, f"{p} mentions the oracle"


def test_sass_has_blackwell_instructions(pbx_lib):
    """tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM, fp64 tensor -> DMMA."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    from portblas_b200 import _lib
    out = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "DMMA"):
        assert mnemonic in out, f"{mnemonic} missing from SASS"

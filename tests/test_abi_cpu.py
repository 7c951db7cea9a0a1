"""CPU checks of the drop-in boundary: the C-ABI library builds/loads here (nvcc cross-compiles, no GPU needed),
exports every symbol include/dfine_sm100.h declares, and the header matches the DFINE_API definitions in csrc/.
No compute entry point is called."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def built_lib():
    from custom_d_fine_b200.build import build
    return build()


def test_library_exports_every_declared_symbol(built_lib):
    from custom_d_fine_b200 import cuda_ops
    protos = cuda_ops.abi_prototypes()
    assert len(protos) >= 35
    L = ctypes.CDLL(str(built_lib))
    missing = [n for n in protos if not hasattr(L, n)]
    assert not missing, missing
    L.dfine_abi_version.restype = ctypes.c_int
    assert L.dfine_abi_version() == 1
    L.dfine_last_error.restype = ctypes.c_char_p
    assert isinstance(L.dfine_last_error(), bytes)


def test_header_matches_sources():
    from custom_d_fine_b200 import cuda_ops
    declared = set(cuda_ops.abi_prototypes())
    defined = set()
    for src in (ROOT / "custom_d_fine_b200" / "csrc").glob("*.cu"):
        defined |= set(re.findall(r"DFINE_API\s+(?:const\s+char\*|int|long)\s+(dfine_\w+)\s*\(", src.read_text()))
    assert declared == defined, (sorted(declared - defined), sorted(defined - declared))


def test_product_path_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing in the package may import it."""
    for py in (ROOT / "custom_d_fine_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), py


def test_ops_fail_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from custom_d_fine_b200.cuda_ops import CudaOps
    with pytest.raises(RuntimeError):
        CudaOps()


def test_loss_descriptor_layout_matches_the_library(built_lib):
    """The ctypes mirror of include/dfine_loss_desc.h has the size the compiled library uses (no compute call)."""
    from custom_d_fine_b200 import loss_desc
    L = ctypes.CDLL(str(built_lib))
    L.dfine_loss_desc_size.restype = ctypes.c_int
    assert L.dfine_loss_desc_size() == ctypes.sizeof(loss_desc.LossDesc)
    L.dfine_loss_out_count.restype = ctypes.c_int
    assert L.dfine_loss_out_count(4) == loss_desc.out_count(4)


def test_every_kernel_waits_for_its_stream_predecessor(built_lib):
    """Programmatic dependent launch: every launch of the library carries the stream-serialization attribute (launch_k), so
    EVERY kernel must execute griddepcontrol.wait (SASS: ACQBULK) before it touches global memory — a kernel without it could
    start on its predecessor's unfinished output.  Checked on the shipped binary; no launch may bypass launch_k."""
    import shutil
    import subprocess
    csrc = ROOT / "custom_d_fine_b200" / "csrc"
    for src in csrc.glob("*.cu"):
        assert "<<<" not in src.read_text(), f"{src.name}: raw <<< >>> launch (use launch_k)"
    n_kernels = sum(len(re.findall(r"__global__", s.read_text())) for s in csrc.glob("*.cu"))
    n_entries = sum(len(re.findall(r"^\s*pdl_entry\(\);", s.read_text(), flags=re.M)) for s in csrc.glob("*.cu"))
    assert n_kernels == n_entries > 50, (n_kernels, n_entries)
    # ... as the FIRST statement of the kernel (the tcgen05 kernels run their shared-memory / tensor-memory prologue first)
    for src in csrc.glob("*.cu"):
        text = src.read_text()
        for m in re.finditer(r"__global__", text):
            head = text[m.start():m.start() + 1500]
            body = head[head.index("{", head.index(")" + " {") if ")" + " {" in head else 0) + 1:]
            name = re.search(r"(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)(, \d+\))?|__cluster_dims__\([^)]*\)", "", head)).group(1)
            if not name.startswith("tc_"):
                assert body.lstrip().startswith("pdl_entry();"), (src.name, name)
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(built_lib)], capture_output=True, text=True, check=True).stdout
    cur, waits = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            waits[cur] = 0
        elif cur and re.search(r"\bACQBULK\b", line):
            waits[cur] += 1
    assert len(waits) > 100
    missing = [k for k, v in waits.items() if v == 0]
    assert not missing, missing[:5]

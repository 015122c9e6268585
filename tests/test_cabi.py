"""The C-ABI library loads and exports every symbol include/speechclip_b200.h declares (no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from speechclip_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "speechclip_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    lib.ensure_built()
    assert os.path.exists(lib.LIB_PATH)
    so = ctypes.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30, names
    for n in names:
        assert hasattr(so, n), f"{n} declared in the header but not exported"
    assert set(lib.EXPORTS) == set(names), set(lib.EXPORTS) ^ set(names)


def test_abi_version_and_error_plumbing():
    so = lib.load()
    assert so.scb_abi_version() == 4
    assert so.scb_launch_count() >= 0
    # argument validation happens before any CUDA call: a NULL args struct is rejected with a message
    assert so.scb_gemm(None, None) == -1
    assert b"NULL" in so.scb_last_error()


def test_gemm_args_struct_matches_header_layout():
    # 2 x (ptr + 4 i64 | 6 i32) ... : spot-check the ctypes mirror against sizeof computed from the header's field list
    text = open(os.path.join(ROOT, "include", "speechclip_b200.h")).read()
    body = text[text.index("typedef struct scb_gemm_args {"):text.index("} scb_gemm_args;")]
    n_fields = len(re.findall(r"\b(?:a|a_inner|a_rows|a_row_stride|a_batch_stride|batch|m_per_batch|kb_per_tap|tap_row_shift|a_col0|"
                              r"a_group_cols|b|b_row_stride|b_group_stride|n|k|groups|out|out_dtype|out_group_cols|ldc|"
                              r"out_batch_stride|out2|out2_dtype|ab_format|bias|residual|residual_dtype|act|alpha|residual_ld|"
                              r"residual_batch_stride|workspace|workspace_bytes)\s*[;,]", body))
    assert n_fields == len(lib.GemmArgs._fields_) == 34

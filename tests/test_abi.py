"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/vqa_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqa_[a-zA-Z0-9_]+)\s*\(", text)))


def test_build_and_exports(pkg):
    import __graft_entry__ as entry
    entry.build()
    lib = pkg.lib.load()
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "libvqa_b200.so does not export %s" % name
    assert sorted(pkg.lib.EXPORTS) == declared          # the ctypes table covers exactly the header
    assert lib.vqa_version() >= 100
    assert isinstance(lib.vqa_last_error(), bytes)


def test_plan_objects_work_without_a_gpu(pkg):
    """Plans are host objects: create / record-free size / destroy must not need a device."""
    lib = pkg.lib.load()
    plan = lib.vqa_plan_create()
    assert plan
    assert lib.vqa_plan_size(plan) == 0
    assert lib.vqa_plan_destroy(plan) == 0


def test_argument_validation_is_reported_not_thrown(pkg):
    """Bad arguments return non-zero and set vqa_last_error() (reference callers see a RuntimeError)."""
    lib = pkg.lib.load()
    plan = lib.vqa_plan_create()
    a = pkg.lib.GemmArgs()
    a.M, a.N, a.K, a.bn, a.split_k = 128, 128, 128, 96, 1        # bn must be 64 / 128 / 256
    a.lda = a.ldb = a.ldo = 128
    rc = lib.vqa_gemm_bf16(plan, ctypes.byref(a), None)
    assert rc != 0 and b"bn" in lib.vqa_last_error()
    assert lib.vqa_rmsnorm_fwd(plan, None, None, None, None, None, 4, 100, 1e-6, 0.0, 0, None, None) != 0
    assert b"feature size" in lib.vqa_last_error()
    lib.vqa_plan_destroy(plan)
    try:
        pkg.lib.check(rc, "gemm")
        raise AssertionError("check() must raise")
    except RuntimeError:
        pass

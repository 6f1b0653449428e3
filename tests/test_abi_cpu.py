"""The C-ABI shared library: it loads without a GPU, exports every function
include/mjpdes_b200.h declares, agrees with the ctypes mirror on struct layout, and
fails loudly (status code, no crash, no CPU fallback) when no device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from mjpdes_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mjpdes_b200.h")


def _ensure_built():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mjp_[a-z_0-9]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    _ensure_built()
    lib = C.CDLL(abi.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(abi.EXPORTED_SYMBOLS) == names


def test_abi_version_and_aggregate_len():
    _ensure_built()
    lib = abi.load_library()
    assert lib.mjp_abi_version() == abi.ABI_VERSION
    assert lib.mjp_aggregate_len(5) == abi.aggregate_len(5) == 33
    assert lib.mjp_aggregate_len(50) == 303


def test_struct_layout_matches_header(tmp_path):
    """sizeof/offsetof from the C compiler vs the ctypes mirror."""
    fields = {
        "mjp_line_desc": (abi.LineDesc, {"M": "M", "K": "K", "lambda": "lam", "N": "N", "w": "w", "warm_up": "warm_up",
                                         "run_to_job": "run_to_job", "jm2dm": "jm2dm", "up_down": "up_down",
                                         "max_lines": "max_lines", "lambda_inst": "lam_inst", "w_inst": "w_inst"}),
        "mjp_rng_spec": (abi.RngSpec, {"mode": "mode", "seed": "seed", "first_instance_id": "first_instance_id"}),
        "mjp_stats_out": (abi.StatsOut, {"njobs": "njobs", "occ_time": "occ_time", "event_hash": "event_hash",
                                         "status": "status", "aggregate": "aggregate"}),
        "mjp_log_record": (abi.LogRecord, {"clk": "clk", "aux": "aux", "job": "job", "instance": "instance",
                                           "line": "line", "n": "n", "ord": "ord", "act": "act", "machine": "machine"}),
        "mjp_log_out": (abi.LogOut, {"records": "records", "capacity": "capacity", "count": "count", "dropped": "dropped",
                                     "first_state": "first_state"}),
        "mjp_config": (abi.Config, {"device": "device", "flags": "flags", "max_instances": "max_instances"}),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for sname, (_, fs) in fields.items():
        lines.append(f'printf("{sname} %zu\\n", sizeof({sname}));')
        for cf in fs:
            lines.append(f'printf("{sname}.{cf} %zu\\n", offsetof({sname}, {cf}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for sname, (ct, fs) in fields.items():
        assert int(out[sname]) == C.sizeof(ct), sname
        for cf, pf in fs.items():
            assert int(out[f"{sname}.{cf}"]) == getattr(ct, pf).offset, f"{sname}.{cf}"


def test_no_device_is_an_error_not_a_fallback():
    _ensure_built()
    lib = abi.load_library()
    if lib.mjp_device_count() > 0:
        pytest.skip("a GPU is present")
    from mjpdes_b200 import engine
    with pytest.raises(abi.MjpError):
        engine.Handle(device=0)
    h = C.c_void_p()
    cfg = abi.Config(0, 0, 0)
    assert lib.mjp_create(C.byref(cfg), C.byref(h)) == abi.ERR_CUDA
    assert not h.value
    assert b"mjp_create" in lib.mjp_last_error(None)


def test_missing_library_raises(tmp_path):
    with pytest.raises(ImportError):
        abi.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under mjpdes_b200/ may import, link or name it."""
    pkg = os.path.join(ROOT, "mjpdes_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dp, fn)).read()
                assert "mjp_oracle" not in text and "oracle." not in text and "import oracle" not in text and \
                    "hostsim" not in text, f"{fn} refers to the oracle / host shim"


def test_issue_models_are_tied_to_the_built_kernels():
    """bench.py's roofline reads warp-instructions per event from profiles/issue_model*.json; each file names the
    kernel it was captured from and carries a hash of that kernel's instruction stream in the built library, so a
    capture of other code shows as roofline.stale instead of being reused silently."""
    import glob
    import json
    import shutil
    import bench
    assert bench.mangled_line_kernel("void mjp_line_kernel<unsigned long, 0, 0, 0, 1, 0, 1>(KParams)") == \
        "_ZN3mjp15mjp_line_kernelImLi0ELb0ELb0ELb1ELb0ELb1EEEvNS_7KParamsE"
    assert bench.mangled_line_kernel("something else") is None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "issue_model*.json")))
    assert len(files) >= 8
    for f in files:
        d = json.load(open(f))
        assert bench.mangled_line_kernel(d["kernel"]) and d["kernel_source_sha256"] and d["kernel_sass_sha256"]
    if shutil.which("cuobjdump") and os.path.exists(abi.LIB_PATH):
        m = bench.issue_model("")
        assert m["stale_basis"] == "kernel machine code" and m["stale"] in (False, True)
        h = bench.kernel_sass_hash(m["kernel"])
        assert h and h == bench.kernel_sass_hash(m["kernel"])                 # deterministic

"""The C-ABI library loads on a CPU-only box and exports every symbol include/bridge_b200.h declares."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as G
    path = os.path.join(ROOT, "bridge.jl_b200", "lib", "libbridge_b200.so")
    if not os.path.exists(path):
        G.build()
    return ctypes.CDLL(path)


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bridge_b200.h")).read()
    names = sorted(set(re.findall(r"\b(bb_[a-zA-Z0-9_]+)\(", hdr)))
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    import bridge_jl_b200 as B
    assert sorted(B.SYMBOLS) == names  # the Python binding covers the whole ABI


def test_status_strings_are_the_reference_messages(lib):
    lib.bb_strerror.restype = ctypes.c_char_p
    assert lib.bb_strerror(-1) == b"Y and W differ in length."                              # src/euler.jl:137
    assert lib.bb_strerror(-2) == b"Time axis mismatch between bridge P and driving W."     # src/euler.jl:248
    assert lib.bb_strerror(-3) == b"Starting point has wrong length."                       # src/sde!.jl:30
    assert lib.bb_abi_version() == 1


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to compute (BB_ERR_NODEVICE); it never routes to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = ctypes.c_void_p()
    assert lib.bb_ctx_create(0, ctypes.byref(ctx)) == -10
    import bridge_jl_b200 as B
    with pytest.raises(B.BridgeError) as ei:
        B.Context(0)
    assert ei.value.status == -10
    # nothing under the product tree loads, links, includes or imports the oracle (comments may NAME the oracle
    # builds a kernel is bit-identical to; code may not touch them)
    import re
    import subprocess
    for dp, _, fs in os.walk(os.path.join(ROOT, "bridge.jl_b200")):
        if os.sep + "build" in dp:
            continue
        for f in fs:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".cpp", "Makefile")):
                continue
            text = open(os.path.join(dp, f), errors="ignore").read()
            assert "from oracle" not in text and "import oracle" not in text, f
            assert not re.search(r"#\s*include\s*[\"<][^\">]*oracle", text), f
            for m in re.finditer(r"(dlopen|CDLL|LoadLibrary|cdll\.|-l\s*oracle|subprocess)[^\n]*", text):
                assert "oracle" not in m.group(0), (f, m.group(0))
    needed = subprocess.run(["readelf", "-d", os.path.join(ROOT, "bridge.jl_b200", "lib", "libbridge_b200.so")],
                            capture_output=True, text=True).stdout
    assert "oracle" not in needed and "NEEDED" in needed


def test_julia_shim_agrees_with_the_header():
    """julia/BridgeB200.jl cannot be executed here (no Julia): every ccall in it is checked mechanically against
    include/bridge_b200.h (name, arity, argument classes and widths, return type) and the hand-mirrored structs against
    gcc's sizeof / offsetof (tools/check_shim.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_shim
    problems, ncalls, nfun, ndecl = check_shim.check()
    assert not problems, problems
    assert ncalls >= 50 and nfun >= 38
    # block structure (brackets, block openers vs `end`) of the Julia sources nobody can run here; the checker itself is
    # exercised on a mutated copy so that a silent pass means something
    assert check_shim.jl_structure() == []
    assert check_shim.jl_structure(os.path.join(ROOT, "baseline", "julia_cpu.jl")) == []
    shim = open(os.path.join(ROOT, "julia", "BridgeB200.jl")).read()
    import tempfile
    for mutated in (shim.replace("lower, upper\nend", "lower, upper\nend\nend", 1), shim.replace("similar(lower)", "similar(lower", 1)):
        assert mutated != shim
        with tempfile.NamedTemporaryFile("w", suffix=".jl", delete=False) as f:
            f.write(mutated)
        try:
            assert check_shim.jl_structure(f.name)
        finally:
            os.unlink(f.name)
    # the reference's own signatures are there (SURVEY 8a "exact reference signatures")
    src = open(os.path.join(ROOT, "julia", "BridgeB200.jl")).read()
    for sig in ("function solve!(s::EulerMaruyama, Y::SamplePath, u, W::SamplePath, Po::Proposal)",
                "function llikelihood(r::LeftRule, X::SamplePath, Po::Proposal; skip = 0)",
                "bridge!(Y::SamplePath, W::SamplePath, Po::Proposal)", "bridge!(X::SamplePath, x0, W::SamplePath, Po::Proposal)",
                "function sample!(W::SamplePath{T}, P::Wiener{T}, y1 = W.yy[1])",
                "function Guide(Po::GuidedBridge;", "function Guide(Po::PartialBridge;", "function Guide(Po::PartialBridgeνH;"):
        assert sig in src, sig


def test_committed_traffic_records_belong_to_the_committed_kernels():
    """bench.py reports `roofline.traffic` only while the hash of the kernel sources equals the one recorded with the ncu
    capture (profiles/ncu_traffic.json); the committed records must be those of the committed sources."""
    import json
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for config in (2, 3, 4, 5):
        assert rec[str(config)]["src_hash"] == bench.kernel_source_hash(config), config
        t, note = bench.measured_traffic(config, 1.0)
        assert t is not None and abs(t - rec[str(config)]["dram_bytes_per_unit"]) < 1e-9, note

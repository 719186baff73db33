"""CPU tests of the drop-in boundary: librdm_sm100.so loads without a GPU and exports every symbol that
include/rdm_sm100.h declares; the ctypes table of rdmnet_b200/_lib.py covers exactly those symbols; the host shim
fails loudly (RuntimeError, the reference extension's error convention) instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rdm_sm100.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rdm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "rdm_grid_subsample" in syms and "rdm_radius_search" in syms and "rdm_kpconv_gather" in syms
    assert len(syms) >= 25


def test_library_exports_every_header_symbol():
    from rdmnet_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "librdm_sm100.so not built (python __graft_entry__.py)"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    from rdmnet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    # argument counts agree with the C prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_host_calls_without_gpu_are_safe():
    from rdmnet_b200 import _lib
    lib = _lib.lib()
    assert lib.rdm_version() >= 100
    assert lib.rdm_grid_subsample_workspace(1000, 2) > 0
    assert lib.rdm_radius_search_workspace(1000, 2) > 0
    assert lib.rdm_selfcheck_bucket_table(50000) == 0  # embedded libstdc++ bucket table == this process' unordered_map
    # argument validation happens before any CUDA call
    rc = lib.rdm_maxpool(None, None, 3, 1, 1, 1, 4, None, None)
    assert rc == 1 and b"index_bytes" in lib.rdm_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from rdmnet_b200 import ops
    pts = torch.rand(100, 3)
    lens = torch.tensor([60, 40])
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.grid_subsample(pts, lens, 0.1)
    with pytest.raises(RuntimeError):
        ops.radius_search(pts, pts, lens, lens, 0.2, 8)
    with pytest.raises(RuntimeError, match="float32"):  # dtype preconditions of torch_helper.h:6-35
        ops.grid_subsample(pts.double(), lens, 0.1)


def test_ctypes_struct_mirrors_match_compiled_layout():
    """sizeof / offsetof of every C struct of the header == the ctypes.Structure mirrors of rdmnet_b200/_lib.py."""
    from rdmnet_b200 import _lib as L
    buf = (ctypes.c_int64 * 64)()
    n = L.lib().rdm_abi_layout(ctypes.cast(buf, ctypes.c_void_p), 64)
    got = [int(buf[i]) for i in range(n)]
    mirrors = [L.ProfRecord, L.TfProjJob, L.TfAttnJob, L.UnaryDesc, L.BlockDesc, L.PyramidDesc, L.PyramidCfg,
               L.ThdroformerDesc, L.BackboneDesc, L.BackboneOut, L.MatchDesc, L.MatchIO, L.MatchResult]
    want = [ctypes.sizeof(m) for m in mirrors] + [L.BlockDesc.sigma.offset, L.PyramidDesc.order.offset,
                                                  L.MatchDesc.nms_limit.offset, L.MatchIO.transform.offset,
                                                  L.MatchResult.transform.offset]
    assert got == want, list(zip(got, want))


def test_bench_reference_arm_emits_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract keys. One bounded pair on the host cores (a few seconds)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--pairs", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    line = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
@pytest.mark.parametrize("mode", [[], ["--sweep"], ["--train"], ["--precision", "tf32"]])
def test_bench_gpu_arm_fails_loudly_without_gpu(mode):
    """The product arm of bench.py - the default run, the config-5 sweep, the config-4 training steps and the config-3 precision
    mode - has no CPU path: without a CUDA device it must exit non-zero with a clear message and print no JSON line (a silent
    fallback would void the measurement)."""
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"] + mode, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode != 0
    assert "CUDA device" in out.stderr and "no CPU path" in out.stderr
    assert out.stdout.strip() == ""

"""The C-ABI library loads and exports every symbol include/*.h declares; compute calls
fail loudly without a GPU (no CPU fallback).  CPU only."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from ldpc_b200 import _native as N
    L = N.lib()
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"
    assert L.ldpc_abi_version() == N.ABI_VERSION


def test_every_header_entry_cites_the_reference():
    src = open(os.path.join(ROOT, "include", "ldpc_b200.h")).read()
    for needle in ("bp/bp.py:43-51", "bp/masking.py:12-147", "ofdm/ofdm_functions.py:131-163",
                   "evaluate_quantized_snr.py:169-188"):
        assert needle in src


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ldpc_b200 import _native as N
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.codes import peg_64_32
    assert N.lib().ldpc_device_count() == 0
    with pytest.raises(N.LdpcError):
        LdpcCode(peg_64_32()[0])
    from ofdm.ofdm_functions import decode_bits
    with pytest.raises(N.LdpcError):
        decode_bits(np.zeros((4, 64)), peg_64_32()[0], 3, 2, 20)


def test_argument_validation_without_gpu():
    from ldpc_b200 import _native as N
    L = N.lib()
    h = ctypes.c_void_p()
    bad = np.array([1, 2], dtype=np.int32)
    assert L.ldpc_code_create(bad.ctypes.data, bad.ctypes.data, 1, 4, 0, None, ctypes.byref(h)) == N.EINVAL
    assert b"row_ptr" in L.ldpc_last_error()
    assert L.ldpc_decode(None, None, 0, 1, 1, 0, 1.0, 1.0, *([None] * 8)) == N.EINVAL


def test_product_never_imports_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ldpc-sims_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
            txt = open(path, errors="ignore").read()
            assert not re.search(r"^\s*(import|from)\s+(bp_oracle|c_oracle|linksim_oracle|oracle)\b", txt, flags=re.M), path
            assert "libldpc_oracle" not in txt, path
            for line in txt.splitlines():
                if "sys.path" in line:
                    assert "oracle" not in line, (path, line)


def test_runtime_specialisation_builds_and_registers(tmp_path):
    """ldpc_b200.jit: a quasi-cyclic prototype outside the built-in family is compiled with the system nvcc into a plug-in
    and registered (no device needed for either step); unsupported shapes and non-plug-ins are refused with a message."""
    import numpy as np
    from ldpc_b200 import jit, _native as N
    if jit.find_nvcc() is None:
        import pytest
        pytest.skip("no nvcc")
    rng = np.random.RandomState(3)
    proto = rng.randint(-1, 31, size=(5, 10)).astype(np.int16)
    proto[:, :2] = rng.randint(0, 31, size=(5, 2))
    src = jit.source_for(proto, 31)
    assert "QcCodeImplLite" in src and "static constexpr int Z = 31, MB = 5, NB = 10;" in src
    so = jit.build(proto, 31, cache_dir=str(tmp_path))
    assert so == jit.build(proto, 31, cache_dir=str(tmp_path))                       # cached by content
    rid = N.lib().ldpc_qc_register_plugin(so.encode())
    assert rid >= 12 and N.lib().ldpc_qc_register_plugin(so.encode()) == rid          # after the twelve built-in codes; idempotent
    assert N.lib().ldpc_qc_register_plugin(b"/no/such/file.so") < 0 and b"ldpc_qc_register_plugin" in N.lib().ldpc_last_error()
    bogus = tmp_path / "bogus.so"
    import shutil, ctypes.util
    shutil.copy(ctypes.util.find_library("m") and "/lib/x86_64-linux-gnu/libm.so.6" or so, bogus)
    assert N.lib().ldpc_qc_register_plugin(str(bogus).encode()) < 0                  # a shared object, but not a plug-in
    ok, why = jit.supported(np.zeros((40, 68), np.int16), 384)
    assert not ok and "block columns" in why


def test_header_is_plain_c_and_reference_arm_prints_the_contract_line(tmp_path):
    """include/ldpc_b200.h must be consumable from C (the drop-in boundary is a C ABI), and `bench.py --impl reference`
    (the CPU arm: oracle port on host cores, no GPU) must print one JSON line with the contract's keys."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include "ldpc_b200.h"\nint main(void) { ldpc_sim_params_t p; ldpc_decode_params_t d; (void)p; (void)d; return ldpc_abi_version() < 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), "-c", str(src), "-o", str(tmp_path / "use.o")])
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.3"],
                         capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Gbit/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "workload" in line["config"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 of a torchrun launch exit without work
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""

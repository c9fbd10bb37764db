"""The C-ABI library loads on a CPU-only host and exports every symbol include/kvq.h declares.
No compute entry point is exercised here (there is no GPU); argument validation and the loud failure are."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from kindergarten_vq_vae_b200 import _lib
    return _lib.load()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "kvq.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kvq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from kindergarten_vq_vae_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/kvq.h but not exported by libkvq.so"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == declared


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "kvq.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)       # declarations only (comments cite the reference's torch calls)
    assert "torch" not in code.lower() and "at::" not in code and "Tensor" not in code and "#include <cuda" not in code


def test_version_and_workspace(lib):
    assert lib.kvq_version() == 100
    small = lib.kvq_workspace_bytes(4096, 768, 512)
    big = lib.kvq_workspace_bytes(1 << 20, 256, 65536)
    assert 0 < small < big
    assert big >= (1 << 20) * 8        # room for one packed key / one bucket slot per latent
    assert lib.kvq_workspace_bytes(-1, 256, 512) == 0


def test_pack_key_orders_by_score_then_index(lib):
    f = lambda s, i: lib.kvq_pack_key(ctypes.c_float(s), ctypes.c_uint32(i))
    scores = [-float("inf"), -3.5e10, -1.0, -1e-38, 0.0, 1e-38, 0.5, 7e20, float("inf")]
    keys = [f(s, 9) for s in scores]
    assert keys == sorted(keys) and len(set(keys)) == len(keys)
    assert f(1.25, 3) < f(1.25, 4) < f(1.25, 0xFFFFFFFF)      # same score: lower index wins
    assert f(-0.0, 7) == f(0.0, 7)                             # -0 and +0 compare equal as floats
    assert f(2.0, 0) > f(1.0, 0xFFFFFFFF)                      # score dominates index
    assert f(1.0, 123456) & 0xFFFFFFFF == 123456
    # torch.argmin treats NaN as the minimum (VectorQuantizer.py:65): a NaN score packs below -inf, lowest index first
    nan = float("nan")
    assert f(nan, 5) < f(-float("inf"), 0) and f(nan, 5) < f(nan, 6) and f(-nan, 5) == f(nan, 5)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_compute_entry_points_fail_loudly_without_gpu(lib):
    assert lib.kvq_device_info(None, None, None) != 0
    assert b"no CUDA device" in lib.kvq_last_error()
    rc = lib.kvq_forward(None, None, 16, 32, 8, 0.25, 0, None, None, None, None, None, None, 0, None)
    assert rc != 0 and len(lib.kvq_last_error()) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_module_refuses_cpu_tensors():
    from kindergarten_vq_vae_b200 import VectorQuantizer, seq_acc, replace_pct_rand_values
    vq = VectorQuantizer(n_e=16, e_dim=32, beta=0.25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq.forward(torch.randn(2, 4, 32), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        seq_acc(torch.zeros(2, 3, dtype=torch.long), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError):
        replace_pct_rand_values(torch.zeros(2, 3, dtype=torch.long), 0.5, 0, 10)


def test_header_is_plain_c_and_a_c_host_links(lib, tmp_path):
    """include/kvq.h compiles as C99 and a C program links against libkvq.so with nothing but the header."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "kvq.h"
int main(void) {
  size_t ws = kvq_workspace_bytes(1 << 20, 256, 65536);
  long long key_lo = kvq_pack_key(-1.0f, 7u), key_hi = kvq_pack_key(2.0f, 0u);
  int rc = kvq_forward(NULL, NULL, 16, 32, 8, 0.25f, KVQ_SEARCH_AUTO, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL);
  printf("%d %zu %d %d\n", kvq_version(), ws, key_lo < key_hi, rc != KVQ_OK);
  return 0;
}
''')
    pkg = os.path.join(ROOT, "kindergarten-vq-vae_b200")
    exe = tmp_path / "host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", pkg, "-l:libkvq.so", f"-Wl,-rpath,{pkg}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert out[0] == "100" and int(out[1]) > 0 and out[2] == "1" and out[3] == "1"


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU, oracle port) emits one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "latents/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("VQ latents/sec (fwd+bwd) at K=65536,D=256")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["value"] > 0


def test_pack_key_order_property(lib):
    """Signed order of packed keys == lexicographic (score, index) order, for arbitrary finite / infinite floats."""
    import struct
    from hypothesis import given, settings, strategies as st

    f32 = st.floats(width=32, allow_nan=False, allow_infinity=True)
    u32 = st.integers(min_value=0, max_value=0xFFFFFFFF)

    @settings(max_examples=400, deadline=None)
    @given(f32, u32, f32, u32)
    def check(s1, i1, s2, i2):
        k1 = lib.kvq_pack_key(ctypes.c_float(s1), ctypes.c_uint32(i1))
        k2 = lib.kvq_pack_key(ctypes.c_float(s2), ctypes.c_uint32(i2))
        a = (struct.unpack("f", struct.pack("f", s1))[0] + 0.0, i1)      # -0.0 folds into +0.0
        b = (struct.unpack("f", struct.pack("f", s2))[0] + 0.0, i2)
        assert (k1 < k2) == (a < b) and (k1 == k2) == (a == b)
        assert k1 & 0xFFFFFFFF == i1

    check()


def _plan(lib, N, K, kind, sms=148, D=256):
    out = (ctypes.c_int64 * 10)()
    rc = lib.kvq_search_plan(N, D, K, kind, sms, out)
    assert rc == 0, lib.kvq_last_error()
    names = ("row_groups", "code_tiles", "ksplit", "tiles_per_split", "main_items", "tail_group0", "tail_split",
             "tail_tiles", "tail_rows", "n_items")
    return dict(zip(names, [int(v) for v in out]))


def test_search_plan_known_cases(lib):
    """Host arithmetic of the tensor-core search's item decomposition (no device needed)."""
    # headline shape, default mode: 4096 row groups = 55 full rounds of 74 CTA pairs + 26 groups cut into 2 code ranges
    p = _plan(lib, 1 << 20, 65536, kind=1)
    assert (p["row_groups"], p["code_tiles"], p["ksplit"]) == (4096, 256, 1)
    assert (p["main_items"], p["tail_group0"], p["tail_split"], p["tail_tiles"]) == (4070, 4070, 2, 128)
    assert p["tail_rows"] == 26 * 256 and p["n_items"] == 4070 + 52
    # the host pipeline's 4096-row lead-in chunk: 16 groups, each cut into 4 ranges of 64 code tiles
    p = _plan(lib, 4096, 65536, kind=1)
    assert (p["main_items"], p["tail_group0"], p["tail_split"], p["tail_tiles"], p["n_items"]) == (0, 0, 4, 64, 64)
    # one full wave: nothing to split
    p = _plan(lib, 74 * 256, 65536, kind=1)
    assert p["tail_split"] == 1 and p["n_items"] == 74
    # the reference's own shape: 2 code tiles cannot be cut (8 tiles per range at least), top-2 search stays unsplit
    p = _plan(lib, 4096, 512, kind=1, D=768)
    assert (p["ksplit"], p["tail_split"], p["n_items"]) == (1, 1, 16)
    # plain argmin: few rows facing a large codebook split EVERY group's code range (merged through packed keys) ...
    p = _plan(lib, 200, 16384, kind=0)
    assert p["row_groups"] == 1 and p["ksplit"] == 64 and p["tail_split"] == 1 and p["n_items"] == 64
    # ... and many rows split the tail round like the top-2 search (sharded codebook, 131072 codes per rank)
    p = _plan(lib, 1 << 20, 131072, kind=0)
    assert (p["ksplit"], p["tail_group0"], p["tail_split"], p["tail_tiles"]) == (1, 4070, 2, 256)
    assert lib.kvq_search_plan(100, 36, 512, 1, 148, (ctypes.c_int64 * 10)()) != 0        # D % 32 != 0: not this kernel


@pytest.mark.parametrize("kind", [0, 1])
def test_search_plan_invariants(lib, kind):
    """Every code tile of every row group is covered exactly once, no item is empty, the split tail fits one round of the
    grid and its records fit their workspace block."""
    import random
    rng = random.Random(1234 + kind)
    for _ in range(400):
        sms = rng.choice([148, 148, 132, 64, 2])
        groups = sms // 2
        N = rng.choice([1, 200, 4096, 18944, 20000, 1 << 20, rng.randrange(1, 3_000_000)])
        K = rng.choice([1, 300, 512, 8192, 65536, 140000, rng.randrange(1, 1_200_000)])
        p = _plan(lib, N, K, kind, sms=sms)
        assert p["row_groups"] == -(-N // 256) and p["code_tiles"] == -(-K // 256)
        # main items: ksplit ranges of tiles_per_split tiles cover the tiles, the last range is not empty
        assert p["ksplit"] * p["tiles_per_split"] >= p["code_tiles"] > (p["ksplit"] - 1) * p["tiles_per_split"]
        assert p["main_items"] == p["tail_group0"] * p["ksplit"]
        tail_groups = p["row_groups"] - p["tail_group0"]
        assert p["n_items"] == p["main_items"] + tail_groups * p["tail_split"]
        assert p["tail_rows"] == (tail_groups * 256 if p["tail_split"] > 1 else 0)
        if kind == 1:
            assert p["ksplit"] == 1                          # the top-2 epilogue never splits a whole sweep
        if p["tail_split"] > 1:
            assert p["ksplit"] == 1 and 0 < tail_groups < groups and p["tail_group0"] % groups == 0
            assert p["tail_split"] * p["tail_tiles"] >= p["code_tiles"] > (p["tail_split"] - 1) * p["tail_tiles"]
            assert p["tail_tiles"] >= 8 and tail_groups * p["tail_split"] <= groups
            assert p["tail_split"] * p["tail_rows"] * 16 <= 128 * 256 * 16       # TOP2_TAIL_REC_BYTES
        else:
            assert tail_groups == 0

"""CPU: the C-ABI library loads and exports every symbol include/b2l.h declares; the
ctypes structure mirrors have the layout a C compiler gives the header's structs."""
import ctypes as C
import os
import re
import subprocess

import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2l.h")


@pytest.fixture(scope="module")
def L():
    entry.build()
    from lit_llama_b200 import _lib

    return _lib


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2l_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    names = declared_functions()
    assert len(names) >= 20
    handle = C.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in b2l.h but not exported"
    assert sorted(L.EXPORTS) == names, "ctypes binding and header disagree"
    assert L.lib().b2l_version() == 100


def test_struct_layout_matches_c_compiler(L, tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "b2l.h"\n'
        "int main(void){\n"
        'printf("%zu %zu %zu %zu\\n", sizeof(b2l_q4_linear_args), sizeof(b2l_q4_weight), sizeof(b2l_layer), sizeof(b2l_decode_args));\n'
        'printf("%zu %zu %zu\\n", offsetof(b2l_q4_linear_args, flags), offsetof(b2l_layer, k_cache), offsetof(b2l_decode_args, logits));\n'
        "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [C.sizeof(L.Q4LinearArgs), C.sizeof(L.Q4Weight), C.sizeof(L.Layer), C.sizeof(L.DecodeArgs),
           L.Q4LinearArgs.flags.offset, L.Layer.k_cache.offset, L.DecodeArgs.logits.offset]
    assert [int(v) for v in out] == got


def test_errors_are_reported_not_swallowed(L):
    lib = L.lib()
    rc = lib.b2l_q_linear(None, 0, None, None, None, 0, None, None, 0, 1, 1, 2, 4, 2, None)
    assert rc == -1 and b"null pointer" in lib.b2l_last_error()
    assert lib.b2l_q4_tiled_bytes(128, 64) == 128 * 64 // 2
    assert lib.b2l_q4_tiled_bytes(130, 64) == 256 * 64 // 2  # rows padded to a multiple of 128
    assert lib.b2l_q4_tiled_bytes(128, 48) == 0              # K must be a multiple of 32
    assert lib.b2l_decode_step(None, None) == -1
    # tensor-parallel exchange: buffer size = 2 epoch parities x world senders x n/2 words of 8 bytes; bad arguments say so
    assert lib.b2l_tp_buffer_bytes(2, 4096) == 2 * 2 * 2048 * 8
    assert lib.b2l_tp_buffer_bytes(9, 4096) == 0 and lib.b2l_tp_buffer_bytes(2, 4095) == 0
    assert lib.b2l_tp_allreduce(None, None, None, 0, 0, None) == -1 and b"null pointer" in lib.b2l_last_error()
    comm = L.TPComm()
    comm.rank, comm.world, comm.max_elems = 3, 2, 4096
    assert lib.b2l_tp_allreduce(C.byref(comm), C.c_void_p(16), C.c_void_p(16), 4096, 0, None) == -1
    assert b"bad rank" in lib.b2l_last_error()

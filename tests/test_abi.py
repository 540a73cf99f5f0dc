"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/ecgvit_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'ecgvit_b200.h')


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ecgvit_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for s in ('ecgvit_gemm', 'ecgvit_patchify', 'ecgvit_layernorm_fwd', 'ecgvit_layernorm_bwd', 'ecgvit_attention_fwd',
              'ecgvit_attention_bwd', 'ecgvit_head_fwd', 'ecgvit_head_bwd', 'ecgvit_grad_sumsq', 'ecgvit_adamw_step'):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    import ecg_b200
    lib = ecg_b200._lib.load()
    raw = ctypes.CDLL(ecg_b200._lib.LIB_PATH)
    for s in _declared_symbols():
        assert hasattr(raw, s), f'{s} declared in the header but not exported'
    assert set(_declared_symbols()) == set(ecg_b200._lib.SIGNATURES), 'ctypes table and header disagree'
    assert lib.ecgvit_abi_version() == 7


def test_gemm_args_struct_layout_matches_header():
    import ecg_b200
    g = ecg_b200._lib.GemmArgs
    # int M,N,K | ptr A | i64 lda | int a_kmajor | ptr B | i64 ldb | int b_kmajor, epilogue | ptr out | i64 ldo | 3 ptr | 3 int
    assert g.A.offset == 16 and g.lda.offset == 24 and g.B.offset == 40 and g.out.offset == 64
    assert g.bias.offset == 96 and g.dtype.offset == 104 and g.dropout_p.offset == 116
    assert g.dropout_seed.offset == 120 and ctypes.sizeof(g) == 128


def test_argument_validation_returns_error_without_touching_the_gpu():
    import ecg_b200
    lib = ecg_b200._lib.load()
    assert lib.ecgvit_gemm(None, None) != 0
    assert 'null' in ecg_b200._lib.last_error()
    with pytest.raises(RuntimeError):
        ecg_b200._lib.check(lib.ecgvit_patchify(None, None, 0, 0, 0, 0, 0, 0, None), 'patchify')


def _declared_prototypes():
    """{name: [C parameter types]} parsed from the header (comments stripped)"""
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char \*)\s*\*?\s*(ecgvit_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', src):
        params = [p.strip() for p in m.group(2).replace('\n', ' ').split(',')]
        out[m.group(1)] = [] if params == ['void'] else params
    return out


def test_ctypes_signatures_match_the_header_prototypes():
    """argument count and C type class (pointer / int / int64 / float) of every entry point: the binding in _lib.py is
    what every caller goes through, and a drifted signature corrupts the stack silently"""
    import ecg_b200
    kinds = {ctypes.c_void_p: 'ptr', ctypes.c_char_p: 'ptr', ctypes.c_int: 'int', ctypes.c_int64: 'i64',
             ctypes.c_float: 'f32'}

    def kind_of_c(decl):
        if '*' in decl:
            return 'ptr'
        t = decl.rsplit(' ', 1)[0].replace('const ', '').strip()
        return {'int': 'int', 'int64_t': 'i64', 'float': 'f32'}[t]

    protos = _declared_prototypes()
    assert set(protos) == set(ecg_b200._lib.SIGNATURES)
    for name, params in protos.items():
        sig = ecg_b200._lib.SIGNATURES[name]
        got = ['ptr' if (isinstance(t, type) and issubclass(t, ctypes._Pointer)) else kinds[t] for t in sig]
        assert got == [kind_of_c(p) for p in params], (name, got, params)

"""CPU: the C-ABI shared library loads and exports every symbol that include/vqgan_b200.h declares, the ctypes
signature table mirrors the header, and the product path fails loudly without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'vqgan_b200.h')


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = re.findall(r'\b(?:int|void|size_t|const char\*)\s+(vqb_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S)
    out = {}
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[name] = n
    return out


@pytest.fixture(scope='module')
def built():
    import __graft_entry__ as g
    g.build()
    from vqvae_vqgan_pytorch_lightning_b200 import lib
    return lib


def test_library_exports_every_declared_symbol(built):
    funcs = header_functions()
    assert len(funcs) >= 25
    dll = ctypes.CDLL(built.LIB_PATH)
    for name in funcs:
        assert hasattr(dll, name), f'{name} declared in the header but not exported'


def test_ctypes_table_matches_header(built):
    funcs = header_functions()
    assert set(funcs) == set(built.SIGNATURES), set(funcs) ^ set(built.SIGNATURES)
    for name, nargs in funcs.items():
        assert len(built.SIGNATURES[name][1]) == nargs, name


def test_non_compute_entry_points(built):
    dll = built.load()
    assert dll.vqb_version().decode().startswith('vqgan_b200')
    assert dll.vqb_vq_workspace_bytes(16384, 1024, 256) == 4096
    assert isinstance(dll.vqb_last_error(), bytes)


def test_argument_validation_without_gpu(built):
    """Bad arguments are rejected before any CUDA call, with a readable message."""
    dll = built.load()
    assert dll.vqb_vq_assign(None, None, 0, None, None, None, None, None, 0, 0, 0, None, 0, None) == -1
    assert b'vq_assign' in dll.vqb_last_error()
    assert dll.vqb_conv2d_fwd(7, None, 0, None, None, None, None, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 0, 0.0, 1.0, None) == -1


def test_product_path_has_no_cpu_fallback(built):
    from vqvae_vqgan_pytorch_lightning_b200 import ops
    with pytest.raises(built.VQBError):
        ops.images_to_nhwc(torch.rand(1, 3, 8, 8), torch.float32)          # CPU tensor -> loud failure


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'vqvae_vqgan_pytorch_lightning_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('CPU oracle', '').replace('fp32 oracle', '').replace("the oracle", '') \
                    or 'import oracle' not in src and 'from oracle' not in src, f
                assert 'from oracle' not in src and 'import oracle' not in src, f

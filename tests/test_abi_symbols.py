"""The C-ABI library loads on a CPU-only box and exports exactly what include/sktensor_b200.h
declares and what the ctypes binding expects (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'sktensor_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'SKT_API[^;(]*?\b(skt_\w+)\s*\(', src)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ('skt_mttkrp_dense', 'skt_mttkrp_coo', 'skt_coo_build_host', 'skt_hadamard_pinv',
                 'skt_solve_stats', 'skt_normalise_gram', 'skt_cp_fit', 'skt_ttm_dense', 'skt_gram',
                 'skt_mttkrp_dense_host', 'skt_last_error'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from sktensor import _lib
    assert os.path.exists(_lib.LIB_PATH), 'build the library first: make -C scikit-tensor_b200'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (skt_\w+)', out))
    assert exported == set(declared_symbols())


def test_ctypes_prototypes_match_header():
    from sktensor import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.skt_version() >= 100
    # argument counts agree with the header
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.PROTOTYPES.items():
        m = re.search(r'\b%s\s*\(([^;]*?)\)\s*;' % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(args), (name, n, len(args))


def test_argument_errors_without_a_gpu():
    """Pure argument validation returns SKT_E* codes and a message; nothing touches the device."""
    from sktensor import _lib
    lib = _lib.load()
    shp = _lib.shape_array((4, 5, 6))
    assert lib.skt_mttkrp_dense_workspace(shp, 3, 8, 7) == 0          # mode out of range
    assert b'mode' in lib.skt_last_error()
    assert lib.skt_mttkrp_dense_workspace(shp, 3, 8, 1) >= 256
    rc = lib.skt_mttkrp_dense(None, shp, 3, None, 8, 0, None, None, 0, None)
    assert rc == -1
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.skt_gram(None, 10, 4, None, None, 0, None) == -1
    assert lib.skt_hadamard_pinv(None, 3, 200, 0, None, None, None) == -1


def test_sliced_contraction_eligibility_is_host_logic():
    """What the opt-in INT8 path accepts is decided on the host (skt_planes_supported / skt_planes_pass_supported): the
    engine asks before it builds byte planes, and silently keeps the FP64 passes otherwise."""
    from sktensor import _lib
    lib = _lib.load()

    def ttm_ok(shape, mode, rank):
        return bool(lib.skt_planes_supported(_lib.shape_array(shape), len(shape), mode, rank))

    def pass_ok(shape, split, rank):
        return bool(lib.skt_planes_pass_supported(_lib.shape_array(shape), len(shape), split, rank))

    assert ttm_ok((512, 512, 512), 0, 32) and ttm_ok((512, 512, 512), 2, 32)      # BASELINE configs[1]
    assert not ttm_ok((512, 512, 512), 1, 32)                                     # only the outer modes contract
    assert not ttm_ok((512, 512, 512), 2, 33)                                     # one 32-column block per slice
    assert not ttm_ok((512, 512, 528), 2, 32)                                     # factor planes no longer resident
    assert not ttm_ok((512, 512, 500), 2, 32)                                     # plane rows must be 16-byte aligned
    assert not ttm_ok((30, 5, 7), 0, 4)
    assert pass_ok((160, 160, 160, 160), 2, 64)                                   # BASELINE configs[4]
    assert pass_ok((160, 20, 160, 160), 2, 64)                                    # ... one rank's slab of it on 8 GPUs
    assert pass_ok((512, 512, 528), 2, 32) and pass_ok((64, 32, 32), 1, 50)
    assert not pass_ok((160, 160, 160, 160), 2, 65)
    assert not pass_ok((10, 11, 8), 2, 3)                                         # trailing modes not a multiple of 16 bytes
    assert not pass_ok((16, 16), 0, 4) and not pass_ok((16, 16), 2, 4)            # the split lies strictly inside
    # without a handle the sizing queries answer 0 instead of touching the device
    assert lib.skt_planes_ttm_workspace(None, 0, 8) == 0
    assert lib.skt_planes_pass_workspace(None, 1, 0, 8) == 0
    assert lib.skt_planes_fused_slabs(None, 0, 8, None) == 0
    assert lib.skt_planes_device_bytes(None) == 0 and lib.skt_planes_nslices(None) == 0
    assert lib.skt_planes_destroy(None) == 0


def test_shipped_cubins_are_sm_100a_and_use_the_blackwell_units():
    """The library is built for sm_100a only and its SASS contains what DESIGN.md says the kernels use: TMA tensor loads,
    FP64 tensor-core MMA, and -- in the opt-in sliced contraction -- tcgen05 MMA / tensor-memory loads."""
    import shutil
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not on PATH')
    from sktensor import _lib
    path = _lib.library_path() if hasattr(_lib, 'library_path') else os.path.join(ROOT, 'scikit-tensor_b200', 'lib', 'libsktensor_b200.so')
    archs = subprocess.run(['cuobjdump', '-lelf', path], capture_output=True, text=True, timeout=300).stdout
    assert 'sm_100a' in archs and 'sm_90' not in archs and 'sm_80' not in archs
    sass = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True, timeout=600).stdout
    for mnemonic in ('UTMALDG', 'DMMA.8x8x4', 'UTCIMMA', 'LDTM', 'UTCBAR', 'UCGABAR_ARV', 'LDGSTS'):
        assert mnemonic in sass, mnemonic

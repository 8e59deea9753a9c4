"""CPU: the C-ABI library builds, loads, and exports every symbol include/sga_b200.h declares
(no compute calls -- there is no GPU in the build container)."""
import os
import re

import pytest

from tests.util import ROOT


def _declared():
    src = open(os.path.join(ROOT, 'include', 'sga_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(sga_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    import __graft_entry__ as entry
    entry.build()
    from sgaligner_b200 import _lib
    lib = _lib.get_lib()
    names = _declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names, (sorted(set(names) ^ set(_lib.EXPORTS)))


def test_version_and_error_string():
    from sgaligner_b200 import _lib
    lib = _lib.get_lib()
    assert lib.sga_version() >= 100
    assert isinstance(lib.sga_last_error(), bytes)


def test_library_is_sm100a_only():
    """The shipped cubin targets sm_100a and contains the Blackwell-native instructions."""
    import shutil
    import subprocess
    from sgaligner_b200 import build
    exe = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(exe):
        return
    out = subprocess.run([exe, '-lelf', build.LIB], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    sass = subprocess.run([exe, '-sass', build.LIB], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UBLKCP'):      # tcgen05.mma / tcgen05.ld / cp.async.bulk
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback_in_product_path():
    """The product package must never import the oracle."""
    pkg = os.path.join(ROOT, 'sgaligner_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_header_is_plain_c():
    """The boundary is a C ABI: include/sga_b200.h must compile as C99 (no C++ or torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('gcc not available')
    hdr = os.path.join(ROOT, 'include', 'sga_b200.h')
    r = subprocess.run([gcc, '-x', 'c', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', hdr], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""The node binding as source: bindings/napi/genstark_b200_addon.cc is generated from
include/genstark_b200.h and must export every entry point exactly once; the TypeScript shim may only call exports that exist.
No node in this image: the addon is syntax-checked against a declaration-only napi.h."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'scripts'))
import gen_bindings  # noqa: E402


def _header_names():
    header = re.sub(r'/\*.*?\*/', '', open(os.path.join(ROOT, 'include', 'genstark_b200.h')).read(), flags=re.S)
    return re.findall(r'\b(gs_[a-z0-9_]+)\s*\(', header)


def test_addon_is_current_and_binds_every_entry_point_once():
    text, names = gen_bindings.generate()
    assert open(gen_bindings.OUT).read() == text, 'run python scripts/gen_bindings.py'
    assert sorted(names) == sorted(_header_names())
    for n in names:
        if n not in ('gs_last_error', 'gs_stark_last_error'):            # these two also supply the message of every thrown error
            assert len(re.findall(r'\b%s\(' % n, text)) == 1, f'{n} must be called from exactly one export'
        assert text.count(f'exports.Set("{gen_bindings.camel(n)}"') == 1


def test_typescript_shim_only_calls_existing_exports():
    _, names = gen_bindings.generate()
    exports = {gen_bindings.camel(n) for n in names}
    used = set()
    for f in os.listdir(os.path.join(ROOT, 'bindings', 'ts')):
        if f.endswith('.ts'):
            used |= set(re.findall(r'\bnative\.(\w+)\(', open(os.path.join(ROOT, 'bindings', 'ts', f)).read()))
    assert used and used <= exports, sorted(used - exports)
    # the FiniteField / Hash / MerkleTree methods genSTARK's lib/ calls are all present
    field = open(os.path.join(ROOT, 'bindings', 'ts', 'B200Field.ts')).read()
    for m in ['add', 'sub', 'mul', 'div', 'exp', 'neg', 'prng', 'newVectorFrom', 'newMatrixFrom', 'newMatrixFromVectors', 'addVectorElements',
              'subVectorElements', 'mulVectorElements', 'divVectorElements', 'combineVectors', 'combineManyVectors', 'pluckVector',
              'getPowerSeries', 'matrixRowsToVectors', 'subMatrixElementsFromVectors', 'divMatrixElements', 'transposeVector', 'transposeMatrix',
              'joinMatrixRows', 'interpolateRoots', 'evalPolyAtRoots', 'evalPolysAtRoots', 'interpolate', 'evalPolyAt', 'mulPolys',
              'interpolateQuarticBatch', 'evalQuarticBatch', 'getRootOfUnity']:
        assert re.search(r'^\s+%s\(' % m, field, flags=re.M), m
    hash_ts = open(os.path.join(ROOT, 'bindings', 'ts', 'B200Hash.ts')).read()
    for m in ['digest', 'digestValues', 'mergeVectorRows', 'proveBatch', 'verifyBatch', 'create']:
        assert re.search(r'\b%s\(' % m, hash_ts), m


@pytest.mark.skipif(shutil.which('g++') is None, reason='no host compiler')
def test_addon_is_valid_cpp_against_the_napi_surface_it_uses():
    r = subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-I' + os.path.join(ROOT, 'bindings', 'napi', 'test_stub'),
                        '-I' + os.path.join(ROOT, 'include'), gen_bindings.OUT], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]

"""The oracle is test infrastructure: nothing shipped may import, call or fall back to it."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def imports_of(path):
    tree = ast.parse(open(path).read())
    mods = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            mods |= {a.name.split('.')[0] for a in node.names}
        elif isinstance(node, ast.ImportFrom) and node.level == 0 and node.module:
            mods.add(node.module.split('.')[0])
    return mods


def test_product_package_never_imports_the_oracle_or_scipy():
    pkg = os.path.join(ROOT, 'dskd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                mods = imports_of(os.path.join(dirpath, f))
                assert 'oracle' not in mods, f
                assert 'scipy' not in mods, f            # the LSAP is the C++ one, not SciPy
    for dirpath, _, files in os.walk(os.path.join(pkg, 'csrc')):
        for f in files:
            if f.endswith(('.cu', '.cuh', '.cpp', '.h')):
                assert 'oracle/' not in open(os.path.join(dirpath, f)).read(), f


def test_oracle_is_only_used_as_checker_in_bench_and_entry():
    bench = open(os.path.join(ROOT, 'bench.py')).read()
    tree = ast.parse(bench)
    # every oracle import in bench.py sits inside a function whose name marks it as the CPU baseline leg
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef):
            uses = any(isinstance(n, (ast.Import, ast.ImportFrom)) and
                       ('oracle' in (getattr(n, 'module', None) or '') or
                        any(a.name.split('.')[0] == 'oracle' for a in getattr(n, 'names', [])))
                       for n in ast.walk(node))
            if uses:
                assert 'cpu' in node.name or 'reference' in node.name, node.name
    top_level = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    for n in top_level:
        assert 'oracle' not in (getattr(n, 'module', None) or '')
        assert all(a.name.split('.')[0] != 'oracle' for a in n.names)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    import pytest
    from dskd_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'libdskd_b200.so'))
    with pytest.raises(_lib.DskdError, match='no CPU or eager-PyTorch fallback'):
        _lib.load()

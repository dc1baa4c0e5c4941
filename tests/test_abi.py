"""CPU: the C-ABI library builds, loads, and exports every symbol include/*.h declares
(no compute calls — there is no GPU here); host-side mirror keeps the reference's contract."""
import ctypes
import os
import re
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("tg_build", os.path.join(ROOT, "pytorch-tecogan_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def _declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            syms |= set(re.findall(r"\b(tg_[a-zA-Z0-9_]+)\s*\(", text))
    return syms


def test_build_is_up_to_date_by_content_not_by_file_time(built_lib, tmp_path):
    """The in-tree library carries a hash of the sources it was built from: file times do not survive the copy to a GPU
    box (a stale .so once ran two measurements), contents do."""
    import importlib.util
    import shutil
    spec = importlib.util.spec_from_file_location("tg_build2", os.path.join(ROOT, "pytorch-tecogan_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.up_to_date()                                   # right after build()
    os.utime(os.path.join(ROOT, "include", "tecogan_b200.h"))   # a newer file time alone does not invalidate it
    assert mod.up_to_date()
    # another set of flags (a measurement variant) or a missing / foreign stamp is never "up to date"
    assert not mod.up_to_date(built_lib, defines=("TG_SOMETHING=1",))
    fake = str(tmp_path / "libtecogan_b200.so")
    shutil.copy(built_lib, fake)
    assert not mod.up_to_date(fake)
    with open(fake + ".srchash", "w") as f:
        f.write("0" * 64 + "\n")
    assert not mod.up_to_date(fake)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for s in sorted(declared):
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_binding_table_matches_header(built_lib):
    from tecogan_b200 import _native
    assert set(_native.SIGNATURES) == _declared_symbols()
    lib = _native.load()
    assert lib.tg_version() >= 100


def test_size_queries_need_no_gpu(built_lib):
    from tecogan_b200 import _native
    lib = _native.load()
    assert lib.tg_gen_param_count(16) == 1765251          # SURVEY.md section 2 #1 [probed]
    assert lib.tg_gen_packed_bytes(16) % 256 == 0
    assert lib.tg_gen_workspace_bytes(1, 180, 320) > 0
    assert lib.tg_gen_workspace_bytes(0, 180, 320) == 0
    assert lib.tg_packed_conv_bytes(0, 64, 64) == 9 * 64 * 64 * 2 + 256


def test_bad_arguments_return_error_codes(built_lib):
    from tecogan_b200 import _native
    lib = _native.load()
    rc = lib.tg_space_to_depth(None, None, 1, 3, 4, 4, 4, None)
    assert rc == -1 and b"null" in lib.tg_last_error_string()
    rc = lib.tg_pack_weights(7, None, None, 64, 64, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _native.check(rc)


def test_generator_mirror_matches_reference_contract():
    from oracle import tecogan_oracle as O
    from tecogan_b200 import models
    args = types.SimpleNamespace(num_resblock=16, discrim_resblocks=4, discrim_channels=128)
    with pytest.raises(ValueError, match="No args is provided for generator"):      # code/models.py:65-66
        models.generator(3, None)
    with pytest.raises(ValueError, match="No args is provided for discriminator"):  # code/models.py:100-101
        models.discriminator(None)
    G = models.generator(3, args)
    ref = O.OracleGenerator(3, 16)
    a, b = G.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    D = models.discriminator(args)
    refd = O.OracleDiscriminator(4, 128, 48)
    a, b = D.state_dict(), refd.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)
    # reference checkpoints load unmodified
    G.load_state_dict(ref.state_dict())


def test_no_cpu_fallback():
    from tecogan_b200 import models, ops
    args = types.SimpleNamespace(num_resblock=1)
    G = models.generator(3, args)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            G(torch.zeros(1, 51, 8, 8))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            ops.space_to_depth(torch.zeros(1, 3, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pytorch-tecogan_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f

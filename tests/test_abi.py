"""The C-ABI library loads (no GPU needed) and exports every symbol include/emote_b200.h declares."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "emote_b200.h"


def _declared():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(emote_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert "emote_gemm_bf16" in names and "emote_attention_bf16" in names and len(names) >= 20


def test_library_loads_and_exports_every_declared_symbol():
    from emote_hack_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    lib = _lib.load()
    assert lib.emote_abi_version() == 2
    assert lib.emote_operand_dtype() == (2 if _lib.OPERAND == "fp16" else 1)   # EMOTE_OP_F16 / EMOTE_OP_BF16
    assert isinstance(lib.emote_launch_count(), int)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in emote_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | set(_lib.INTROSPECTION)
    assert bound == set(_declared()), f"ctypes binding out of sync with the header: {bound ^ set(_declared())}"


def test_both_operand_builds_export_the_same_abi():
    """libemote_b200.so (fp16 operands) and libemote_b200_bf16.so are the same sources built twice"""
    import ctypes
    from emote_hack_b200 import _lib
    libdir = _lib.LIB_PATH.parent
    for name, want in (("libemote_b200.so", 2), ("libemote_b200_bf16.so", 1)):
        if not (libdir / name).exists():
            _lib.build()
        lib = ctypes.CDLL(str(libdir / name))
        assert lib.emote_operand_dtype() == want and lib.emote_abi_version() == 2
        for sym in _declared():
            assert hasattr(lib, sym), f"{sym} missing from {name}"


def test_exported_symbols_are_plain_c():
    from emote_hack_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in _declared():
        assert name in exported


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest
    from emote_hack_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.EmoteKernelError):
        _lib.load()


def test_sass_uses_blackwell_tensor_and_tma_paths():
    """UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld (B200_PROFILING.md evidence table)."""
    import shutil
    import pytest
    from emote_hack_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, f"{mnemonic} missing from SASS"


def test_entry_points_reject_bad_arguments_without_touching_the_gpu():
    """Argument validation happens before any CUDA call: invalid shapes / null pointers return EMOTE_ERR_INVALID (1)
    with a message in emote_last_error() — checked here on a box without a GPU."""
    import ctypes as C
    from emote_hack_b200 import _lib
    lib = _lib.load()
    args = _lib.EmoteGemmArgs()
    assert lib.emote_gemm_bf16(None, None, None, C.byref(args), None) != 0
    assert b"null pointer" in lib.emote_last_error()
    buf = (C.c_char * 4096)()
    p = C.cast(buf, C.c_void_p)
    args.M, args.N, args.K, args.conv_taps = 128, 128, 60, 1      # K not a multiple of 8
    assert lib.emote_gemm_bf16(p, p, p, C.byref(args), None) != 0
    assert b"multiple of 8" in lib.emote_last_error()
    args.K, args.conv_taps = 64, 5
    assert lib.emote_gemm_bf16(p, p, p, C.byref(args), None) != 0
    assert b"conv_taps" in lib.emote_last_error()
    args.conv_taps, args.epilogue, args.out_dtype, args.N = 1, 1, 0, 128   # GEGLU needs bf16 output
    assert lib.emote_gemm_bf16(p, p, p, C.byref(args), None) != 0
    assert b"GEGLU" in lib.emote_last_error()
    attn = _lib.EmoteAttnArgs()
    assert lib.emote_attention_tc_bf16(C.byref(attn), None) != 0
    assert lib.emote_attention_tc_supported(40) == 1 and lib.emote_attention_tc_supported(64) == 1 and lib.emote_attention_tc_supported(48) == 0
    assert lib.emote_im2col3x3_s2_pad01(p, 1, 7, 8, 8, p, None) != 0      # odd height
    assert b"even H" in lib.emote_last_error()
    assert lib.emote_upsample2x(p, 1, 4, 4, 12, p, None) != 0             # C % 8 != 0
    assert lib.emote_layernorm(p, 4, 100, p, p, 1e-5, None, 0, 0, p, None) != 0
    assert b"multiple of 64" in lib.emote_last_error()
    with __import__("pytest").raises(_lib.EmoteKernelError):
        _lib.check(1, "emote_layernorm")


def test_tuning_knob_accepts_known_keys_only():
    from emote_hack_b200 import _lib
    lib = _lib.load()
    assert lib.emote_set_tuning(b"gn_reduce", 0) == 0 and lib.emote_set_tuning(b"gn_reduce", 1) == 0
    assert lib.emote_set_tuning(b"no_such_knob", 1) != 0
    assert b"unknown key" in lib.emote_last_error()
    assert lib.emote_set_tuning(None, 1) != 0

"""C-ABI: libmtus_b200.so loads and exports every symbol include/mtus_b200.h declares (no compute, CPU only)."""
import ctypes as C
import os
import re

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mtus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtus_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = _header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtus_b200.h but not exported"


def test_python_binding_covers_the_header():
    import mtus_b200
    declared = set(_header_functions())
    bound = set(mtus_b200._lib.SIGNATURES)
    assert declared == bound, f"header/binding mismatch: {sorted(declared ^ bound)}"


def test_version_and_status_strings(lib):
    assert lib.mtus_version() >= 100
    assert lib.mtus_status_string(0) == b"ok"
    assert b"argument" in lib.mtus_status_string(-1)
    assert b"unsupported" in lib.mtus_status_string(-2)


def test_null_arguments_are_rejected_without_touching_the_gpu(lib):
    # argument validation happens before any CUDA call: usable (and required to hold) on a CPU-only box
    assert lib.mtus_layernorm_fwd(None, None, None, None, None, None, 4, 32, 1e-5, 0, None) == -1
    assert lib.mtus_window_attn_fwd(None, None, None, None, None, 1, 7, 7, 32, 1, 7, 7, 0, 0, 0, None) == -1
    assert lib.mtus_swin_forward(None, None, 1, None, None, None, None, None, 0, 0, None) == -1
    assert lib.mtus_swin_input_grad(None, None, None, None, None, None) == -1
    assert lib.mtus_patch_embed_col2im(None, 48, None, 1, 8, 8, 0, None) == -1
    import ctypes as C
    buf = (C.c_float * 4)()
    assert lib.mtus_patch_embed_col2im(buf, 40, buf, 1, 8, 8, 0, None) == -1        # fewer than 48 columns
    assert lib.mtus_patch_embed_col2im(buf, 48, buf, 1, 6, 8, 0, None) == -1        # height not a multiple of the patch
    assert lib.mtus_patch_embed_col2im(buf, 48, buf, 0, 8, 8, 0, None) == 0         # empty batch: nothing launched


def test_swin_param_layout_matches_the_published_trunk_sizes(lib):
    """Trunk parameter counts (torchvision ``.features`` = timm features_only without the
    final norm / head): swin_t 27 517 818, swin_b 86 741 176."""
    import mtus_b200 as m
    from mtus_b200._native import enumerate_params
    for name, trunk in (("swin_tiny_patch4_window7_224", 27_517_818), ("swin_base_patch4_window7_224", 86_741_176)):
        ed, depths, heads, win = m.SWIN_ARCHS[name]
        cfg = m._lib.SwinConfig()
        cfg.batch, cfg.img_size, cfg.embed_dim, cfg.window = 1, 224, ed, win
        for i in range(4):
            cfg.depths[i], cfg.heads[i] = depths[i], heads[i]
        cfg.dtype, cfg.backend, cfg.training, cfg.ln_eps = 0, 0, 1, 1e-5
        infos = enumerate_params(lib.mtus_swin_param_info, cfg)
        n = 0
        for nm, off, shape in infos:
            k = 1
            for s in shape:
                k *= s
            n += k
            assert off % 8 == 0
        assert n == trunk, (name, n)
        assert lib.mtus_swin_param_count(C.byref(cfg)) >= n
        numel = C.c_int64()
        off = lib.mtus_swin_param_offset(C.byref(cfg), b"layers_2.blocks.3.attn.qkv.bias", C.byref(numel))
        assert off > 0 and numel.value == 3 * 4 * ed
        assert lib.mtus_swin_param_offset(C.byref(cfg), b"nope", C.byref(numel)) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import mtus_b200
    monkeypatch.setattr(mtus_b200._lib, "_lib", None)
    monkeypatch.setattr(mtus_b200._lib, "_SO", str(tmp_path / "absent.so"))
    try:
        mtus_b200._lib.lib()
    except RuntimeError as e:
        assert "no fallback" in str(e).lower()
    else:
        raise AssertionError("a missing libmtus_b200.so must raise")

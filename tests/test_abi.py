"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2r.h declares (no compute calls)."""
import ctypes
import re

from _util import ROOT


def declared_symbols():
    text = (ROOT / "include" / "b2r.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from b200restore import _lib
    syms = declared_symbols()
    assert syms, "no declarations found in include/b2r.h"
    assert sorted(_lib.SYMBOLS) == syms, "ctypes binding list and header disagree"
    lib = ctypes.CDLL(str(_lib.lib_path())) if _lib.lib_path().exists() else _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"{s} is declared in b2r.h but not exported by libb2r.so"


def test_version_and_error_string():
    from b200restore import _lib
    lib = _lib.load()
    assert lib.b2r_version() == _lib.EXPECTED_ABI == 200
    assert lib.b2r_abi_sizeof(0) == ctypes.sizeof(_lib.ConvGemmDesc) and lib.b2r_abi_sizeof(99) == -1
    assert isinstance(lib.b2r_last_error(), bytes)


def test_no_libcuda_link_dependency():
    """libb2r.so must dlopen without a driver: tensor-map encoding is resolved through cudaGetDriverEntryPoint."""
    import subprocess
    from b200restore import _lib
    _lib.load()
    out = subprocess.run(["ldd", str(_lib.lib_path())], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


def test_struct_layout_matches_header():
    """sizeof(b2r_conv_gemm_desc) as ctypes sees it == what a C compiler computes from the header."""
    import subprocess
    import tempfile
    from b200restore import _lib
    src = '#include <stdio.h>\n#include "b2r.h"\nint main(){printf("%zu %zu %zu", sizeof(b2r_conv_gemm_desc), ' \
          '__builtin_offsetof(b2r_conv_gemm_desc, weights), __builtin_offsetof(b2r_conv_gemm_desc, out_C));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        c = f"{td}/s.c"
        open(c, "w").write(src)
        subprocess.run(["gcc", "-I", str(ROOT / "include"), c, "-o", f"{td}/s"], check=True)
        size, off_w, off_oc = map(int, subprocess.run([f"{td}/s"], capture_output=True, text=True).stdout.split())
    assert ctypes.sizeof(_lib.ConvGemmDesc) == size
    assert _lib.ConvGemmDesc.weights.offset == off_w
    assert _lib.ConvGemmDesc.out_C.offset == off_oc


def test_net_weight_bytes_bounds_the_packed_checkpoint():
    """b2r_net_weight_bytes is host-only arithmetic: an upper bound of what b2r_net_create writes (bf16 weights, tap-folded
    copies for C_out = 64 layers, f32 biases, identity shortcut matrices), from the state_dict's shapes alone."""
    import ctypes as C
    from b200restore import _lib, netplan, synth
    lib = _lib.load()
    for arch, params in (("simple_unet", 1_862_979), ("resunet", 12_625_869)):
        sd = synth.synthetic_state_dict(arch, 0)
        arr, keep = netplan.state_to_ctypes(sd)
        n = C.c_size_t()
        assert lib.b2r_net_weight_bytes(arr, len(sd), C.byref(n)) == 0
        assert 2 * params < n.value < 4 * params + (4 << 20) and n.value % 256 == 0
    assert lib.b2r_net_weight_bytes(None, 0, C.byref(n)) == -22 and b"null" in lib.b2r_last_error()

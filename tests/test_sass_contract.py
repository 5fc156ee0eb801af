"""The built library must take the Blackwell paths it claims: `cuobjdump -sass` of the in-tree .so (no GPU needed) shows
tcgen05 MMAs on CTA pairs, tensor-memory loads and TMA loads in the three tensor-core kernels of the sampling path, PDL in
the hidden-layer kernel, and no legacy mma.sync anywhere."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "egohmr_b200", "lib", "libegohmr_b200.so")


@pytest.fixture(scope="module")
def sass_counts():
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, kern = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)          # mangled name: the kernel's identifier is a substring
            counts[kern] = {}
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern][op] = counts[kern].get(op, 0) + 1
    return counts


def _total(c, prefix):
    return sum(v for k, v in c.items() if k == prefix or k.startswith(prefix + "."))


def _kernels(counts, ident):
    ks = [c for name, c in counts.items() if ident in name]
    assert ks, f"no kernel named *{ident}* in the library"
    return ks


@pytest.mark.parametrize("ident", ["gcn_hidden_umma_t_kernel", "conv_gemm_kernel", "linear_umma_kernel"])
def test_tensor_core_kernels_use_tcgen05_pairs_tmem_and_tma(sass_counts, ident):
    for c in _kernels(sass_counts, ident):
        assert _total(c, "UTCHMMA") >= 12            # tcgen05.mma kind::f16: 3 products x 4 k-slices per k-block
        assert _total(c, "UTCHMMA.2CTA") == _total(c, "UTCHMMA")   # every MMA is a cta_group::2 MMA
        assert _total(c, "LDTM") >= 1                # tcgen05.ld
        assert _total(c, "UTMALDG") >= 2             # cp.async.bulk.tensor
        assert _total(c, "UTCBAR") >= 1              # tcgen05.commit -> mbarrier


def test_hidden_layer_kernel_uses_programmatic_dependent_launch(sass_counts):
    (c,) = _kernels(sass_counts, "gcn_hidden_umma_t_kernel")
    assert _total(c, "ACQBULK") >= 1 and _total(c, "PREEXIT") >= 1


def test_no_legacy_tensor_core_path_anywhere(sass_counts):
    for name, c in sass_counts.items():
        assert _total(c, "HMMA") == 0, name

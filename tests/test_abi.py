"""The C-ABI library loads without a GPU, exports every symbol include/ldpcb200.h declares, and every compute
entry point refuses to run without a CUDA device (there is no CPU fallback in the product)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_binding

HEADER = os.path.join(ROOT, "include", "ldpcb200.h")


def declared_symbols():
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(ldpcb200_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = load_binding().lib()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n


def test_header_compiles_as_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "ldpcb200.h"\nint main(void){ ldpcb200_params p; ldpcb200_sim_params s; ldpcb200_counters c; (void)p;(void)s;(void)c; return LDPCB200_VERSION > 0 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_struct_layouts_match_binding():
    """ctypes mirrors of the ABI structs have the sizes the C compiler gives them."""
    L = load_binding()
    prog = r'''
#include <stdio.h>
#include "ldpcb200.h"
int main(void){ printf("%zu %zu %zu\n", sizeof(ldpcb200_params), sizeof(ldpcb200_sim_params), sizeof(ldpcb200_counters)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    assert sizes == [C.sizeof(L.Params), C.sizeof(L.SimParams), C.sizeof(L.Counters)]


def test_sigma_is_host_arithmetic():
    L = load_binding()
    assert L.sigma(16, 32, 0, 2.0) == pytest.approx(np.sqrt(10 ** -0.2 / 2 / 0.5))
    assert L.sigma(46, 68, 2, 2.0, L.MOD_QAM64) == pytest.approx(np.sqrt(10 ** -0.2 / (2 * (22 / 66) * 3 * 2) * 42.0))
    assert L.lib().ldpcb200_version() >= 100


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    L = load_binding()
    hd = np.zeros((2, 4), np.int16)
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(hd, 8, L.LMS_DEC)
    assert e.value.code == L.ENODEV
    with pytest.raises(L.LdpcError) as e:
        L.demodulate(16, 4, 0.5, np.zeros(8))
    assert e.value.code == L.ENODEV


def test_argument_validation_comes_before_the_device():
    L = load_binding()
    hd = np.zeros((2, 4), np.int16)
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(hd, 8, 6)                        # FHT_DEC: GF(q), out of scope
    assert e.value.code == L.EINVAL
    with pytest.raises(L.LdpcError) as e:
        L.Decoder(hd, 8, L.TASP_DEC, precision=32)
    assert e.value.code in (L.EUNSUPPORTED, L.ENODEV)


def test_product_does_not_reference_the_oracle():
    """Nothing under ldpc-lib_b200/ may include, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "ldpc-lib_b200")
    for base, _, files in os.walk(pkg):
        if os.sep + "build" in base or os.sep + "bin" in base:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "ldpc_oracle" not in txt and "pyoracle" not in txt and "libldpcref" not in txt and "oracle/" not in txt.replace("the oracle", ""), f


def test_code_specialised_kernel_generates_and_compiles_without_a_gpu():
    """The run-time generator + NVRTC produce an sm_100a cubin for an arbitrary matrix (no device needed)."""
    import shutil
    if not (os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12") or shutil.which("nvcc")):
        pytest.skip("NVRTC not installed")
    from codes import load_code
    L = load_binding()
    hd, _ = load_code("c4_wifi_12x24")
    assert L.jit_check(hd, 81) > 10000
    big = np.full((60, 70), -1, np.int16)                # 60 block rows x Z = 1000: neither layout variant fits an SM
    for j in range(60):
        big[j, j] = 0; big[j, (j + 7) % 70] = 3; big[j, 65 + j % 5] = 11
    with pytest.raises(L.LdpcError) as e:
        L.jit_check(big, 1000)
    assert e.value.code == L.EUNSUPPORTED


def test_runtime_generator_emits_the_same_tensor_memory_tables_as_the_build_time_one(tmp_path, monkeypatch):
    """spec_jit.cpp (run time, any matrix) and tools/gen_lms_spec.py (build time, benchmark matrices) both derive the
    rotation tables of lms_tmem.cuh / ms_tmem.cuh; they must agree, and the tables must be self-consistent:
    following DELTA through one iteration returns every column to its ROT."""
    import re
    import shutil
    if not (os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12") or shutil.which("nvcc")):
        pytest.skip("NVRTC not installed")
    import importlib.util
    from codes import load_code
    spec = importlib.util.spec_from_file_location("gen_lms_spec", os.path.join(ROOT, "tools", "gen_lms_spec.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    L = load_binding()
    for code, Z in [("ref32x16_b", 256), ("c4_wifi_12x24", 81), ("ref32x16_a", 126)]:
        hd, _ = load_code(code)
        dump = tmp_path / ("%s_%d.cu" % (code, Z))
        monkeypatch.setenv("LDPCB200_JIT_DUMP", str(dump))
        monkeypatch.delenv("LDPCB200_NO_TMEM", raising=False)
        assert L.jit_check(hd, Z) > 10000
        txt = dump.read_text()
        assert "LmsTmem<ldpcb200::gen_jit::Code>" in txt                 # these codes take the tensor-memory variant

        def arr(name):
            m = re.search(r"static constexpr \w+ %s\[\d+\] = \{([^}]*)\}" % name, txt)
            return [int(x) for x in m.group(1).replace(",", " ").split()]
        b, c = hd.shape
        rp, col, sh = [0], [], []
        for j in range(b):
            for i in range(c):
                if hd[j, i] != -1:
                    col.append(i); sh.append(int(hd[j, i]) % Z)
            rp.append(len(col))
        zp = (Z + 31) // 32 * 32
        rot, delta, ri, synsh, tcols, lastw = gen.tmem_tables(b, c, Z, rp, col, sh, zp)
        assert arr("DELTA") == delta and arr("ROT") == rot and arr("RI") == ri and arr("SYNSH") == synsh and arr("LAST") == lastw
        assert int(re.search(r"TCOLS = (\d+)", txt).group(1)) == tcols
        assert arr("EARLY") == gen.early_table(b, rp, col)
        # one pass over the edges in schedule order: a column read at n + DELTA and rewritten lane-aligned ends rotated by ROT
        cur = list(rot)
        for e in range(len(col)):
            assert (cur[col[e]] + delta[e]) % Z == sh[e]
            cur[col[e]] = sh[e]
        assert cur == rot


def test_bench_kernel_sass_uses_tensor_memory_and_stays_within_its_instruction_budget():
    """The SASS of the ahead-of-time bench kernel (lmst_spec_c2t): tensor-memory loads / stores (LDTM / STTM =
    tcgen05.ld / st), three-input minima, packed adds, no local-memory spills, and the per-edge instruction count that
    DESIGN.md §4.0 and bench.py's `issue` block quote (tools/sass_mix.py)."""
    import re
    import shutil
    import subprocess
    obj = os.path.join(ROOT, "ldpc-lib_b200", "build", "lms_spec_aot.o")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(obj) or not os.path.exists(cuobjdump):
        pytest.skip("no build directory / cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "lmst_spec_c2t", obj], capture_output=True, text=True, check=True).stdout
    ops = re.findall(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", sass, re.M)
    assert len(ops) > 1000
    for needed in ("LDTM", "STTM", "FMNMX3", "FADD2", "FFMA"):
        assert needed in ops, needed
    import json
    mix = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_mix.py"), "128", "16", "8", "--json"], input=sass,
                                    capture_output=True, text=True, check=True).stdout)
    pe = mix["per_edge"]
    assert "LDL" not in mix["opcodes"] and "STL" not in mix["opcodes"]      # nothing spilled in the block rows
    assert pe["total"] <= 11.2 and pe["alu"] <= 4.5 and pe["lsu"] <= 2.5, pe
    # bench.py's roofline block quotes the committed copy of this output: it must describe the object that was built
    committed = json.load(open(os.path.join(ROOT, "profiles", "sass_mix_lmst_spec_c2t.json")))
    for k, v in pe.items():
        assert abs(committed["per_edge"][k] - v) < 1e-9, (k, v, committed["per_edge"][k], "run `make -C ldpc-lib_b200 sassmix`")


def test_every_kernel_family_compiles_at_run_time_without_a_gpu():
    """The run-time generator + NVRTC for the flooding min-sum families too (spec_jit.cpp kinds 1 / 2: ms_spec, ms_tmem and the
    fp16-pair IMS_DEC kernel with one and two groups per CTA) -- NVRTC is stricter than nvcc (no host functions, no headers)."""
    import shutil
    import subprocess
    if not (os.path.exists("/usr/local/cuda/lib64/libnvrtc.so.12") or shutil.which("nvcc")):
        pytest.skip("NVRTC not installed")
    for check in ("1,0,2", "2,0,2", "1,2,2", "2,2,2", "2,4,4", "2,5,2"):
        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); from conftest import load_binding; from codes import load_code; "
                "L = load_binding(); hd, _ = load_code('c4_wifi_12x24'); assert L.jit_check(hd, 81) > 10000" % (os.path.join(ROOT, "tests"), ROOT))
        r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, LDPCB200_JIT_CHECK=check), capture_output=True, text=True)
        assert r.returncode == 0, (check, r.stderr[-2000:])


def test_jit_defines_reach_the_run_time_compilation():
    """LDPCB200_JIT_DEFINES (development switch): macro definitions handed to NVRTC change the kernel
    that is built -- here lms_tmem's doubled-column layout with the split mbarrier instead of the padded two-buffer (PP)
    layout; both compile for sm_100a without a device."""
    code = ("import sys, ctypes as C, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); "
            "from codes import load_code; import pyldpcb200 as L; hd = np.ascontiguousarray(load_code('ref32x16_b')[0], np.int16); "
            "n = C.c_int(0); rc = L.lib().ldpcb200_jit_check(hd.ctypes.data_as(C.POINTER(C.c_int16)), 16, 32, 256, 10, 0, C.byref(n)); "
            "print(rc, n.value)") % (os.path.join(ROOT, "tests"), os.path.join(ROOT, "ldpc-lib_b200"))
    sizes = []
    for defs in ("", "-DLMS_TMEM_PP=0"):
        env = dict(os.environ, LDPCB200_JIT_DEFINES=defs)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        rc, n = (int(x) for x in r.stdout.split())
        if rc != 0:
            pytest.skip("NVRTC not available")
        sizes.append(n)
    assert sizes[0] > 0 and sizes[1] > 0 and sizes[0] != sizes[1]

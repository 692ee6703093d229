"""Runs the side-by-side C++ drop-in driver (tests/cpp/dropin_driver.cpp, built by `make -C oracle/ref_build dropin`
into oracle/_ref/): the reference's classes and the pu:: classes of include/pu/pu_dropin.hpp in ONE binary, same
inputs, results compared bit for bit -- the reference's own test drivers for this surface, re-run against both."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "dropin_driver")


@pytest.mark.gpu
@pytest.mark.ref
def test_cpp_dropin_driver_all_pass():
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/dropin_driver not built (needs the reference headers)")
    r = subprocess.run([DRIVER], capture_output=True, text=True, timeout=600)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    assert r.returncode == 0 and "ALL PASS" in r.stdout, tail


@pytest.mark.gpu
@pytest.mark.ref
def test_reference_multiblock_ldpc_test_runs_unmodified_on_the_dropin_classes():
    """/root/reference/tests/test_multiblock_ldpc.cpp compiled UNMODIFIED against tests/cpp/shim/ultra/{fec,ofdm}.hpp (aliases of the
    pu:: classes; no reference object code linked): the reference's own assertions -- 5 rates x {1, 2, 5} blocks, boundary bytes,
    protocol frame sizes, and the full modem pipeline generatePreamble + modulate -> process -> getSoftBits -> decodeSoft for
    60...279-byte frames -- must all pass on libpu_b200.so."""
    exe = os.path.join(ROOT, "oracle", "_ref", "test_multiblock_ldpc_pu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/test_multiblock_ldpc_pu not built (needs the reference sources)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    tail = "\n".join(r.stdout.splitlines()[-30:])
    assert r.returncode == 0 and "ALL TESTS PASSED" in r.stdout, tail + r.stderr[-2000:]


def test_dropin_header_compiles_standalone(tmp_path):
    """Without the reference headers the drop-in header must still compile (stand-alone type mirrors) and link."""
    src = tmp_path / "t.cpp"
    src.write_text('#include "pu/pu_dropin.hpp"\nint main() { pu::ChannelInterleaver ci(60, 648); pu::Interleaver il(6, 108);\n'
                   'pu::LDPCEncoder e(pu::CodeRate::R1_2); auto c = e.encode(pu::Bytes(40, 0x5A));\n'
                   'pu::WaveformPtr (*mk)(pu::protocol::WaveformMode) = &pu::WaveformFactory::create;   // needs a GPU to call\n'
                   'return (c.size() == 81 && ci.getStep() == 181 && il.getPermutation(1) == 6 && pu::WaveformFactory::isSupported(pu::protocol::WaveformMode::AUTO)'
                   ' && sizeof(pu::DPSKDemodulator) > 0 && sizeof(pu::MultiCarrierDPSKDemodulator) > 0 && sizeof(pu::OFDMNvisWaveform) > 0 && mk != nullptr) ? 0 : 1; }\n')
    exe = tmp_path / "t"
    lib = os.path.join(ROOT, "projectultra_b200")
    from projectultra_b200 import build
    build.build()
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", lib, "-lpu_b200", "-Wl,-rpath," + lib])
    assert subprocess.run([str(exe)]).returncode == 0

"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/wae.h declares,
validates arguments with the reference's error text, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    import __graft_entry__ as ge
    ge.build()
    header = open(os.path.join(ROOT, "include", "wae.h")).read()
    declared = set(re.findall(r"WAE_API\s+[\w\s\*]+?\b(wae_\w+)\s*\(", header))
    assert len(declared) >= 40
    lib = ctypes.CDLL(pkg.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert missing == []
    assert set(pkg._binding.WAE_SYMBOLS) <= declared


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.WaeError) as e:
        pkg.Engine(0)
    assert e.value.status == 7  # WAE_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, link or name it
    pkg_dir = os.path.join(ROOT, "web-audio-api-rs_b200")
    for dp, _, files in os.walk(pkg_dir):
        if os.path.basename(dp) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text and "wao_core" not in text, f


def test_header_is_plain_c99_and_struct_layouts_match_the_binding(pkg, tmp_path):
    # the boundary is a C ABI: include/wae.h must compile as C (no C++-isms), and the ctypes structures of the binding must have
    # the sizes the C compiler gives the header's structs (a field added on one side only would silently shift arguments)
    import subprocess
    pairs = [("wae_channel_config", "ChannelConfig"), ("wae_audio_buffer", "AudioBufferDesc"), ("wae_param_event", "ParamEvent"),
             ("wae_oscillator_options", "OscillatorOptions"), ("wae_biquad_options", "BiquadOptions"), ("wae_iir_options", "IirOptions"),
             ("wae_gain_options", "GainOptions"), ("wae_delay_options", "DelayOptions"), ("wae_stereo_panner_options", "StereoPannerOptions"),
             ("wae_panner_options", "PannerOptions"), ("wae_analyser_options", "AnalyserOptions"),
             ("wae_dynamics_compressor_options", "DynamicsCompressorOptions"), ("wae_channel_merger_options", "ChannelMergerOptions"),
             ("wae_channel_splitter_options", "ChannelSplitterOptions"), ("wae_batch_stats", "BatchStats")]
    B = pkg._binding
    pairs = [(c, p) for c, p in pairs if hasattr(B, p)]
    assert len(pairs) >= 10
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "wae.h"\nint main(void) {\n' +
                   "".join('  printf("%s %%zu\\n", sizeof(%s));\n' % (c, c) for c, _ in pairs) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for c, p in pairs:
        assert int(sizes[c]) == ctypes.sizeof(getattr(B, p)), (c, p)


def build_c_client(tmp_path):
    import subprocess
    exe = tmp_path / "c_client"
    pkg_dir = os.path.join(ROOT, "web-audio-api-rs_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_client.c"), "-L", pkg_dir, "-lwae_b200", "-Wl,-rpath," + pkg_dir, "-lm", "-o", str(exe)])
    return str(exe)


def test_plain_c_client_links_and_fails_loudly_without_a_gpu(pkg, tmp_path):
    # examples/c_client.c: the boundary used from C, no Python / torch in the process.  Without a device it must stop at
    # wae_engine_create with WAE_NO_DEVICE (exit code 2) — never render on the CPU
    import subprocess
    import torch
    import __graft_entry__ as ge
    ge.build()
    exe = build_c_client(tmp_path)
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=120)
    # planned on the host, before any device is touched: oscillator -> biquad fused, the gain is automated (ramp): param track + gain + mix
    assert "plan: 4 stage(s) per chunk [k_mix x 1, k_gain x 1, k_chain x 1, k_param x 1], 1 chunk(s) of 48000 frames" in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr

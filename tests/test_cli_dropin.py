"""Drop-in boundary: the reference's UNMODIFIED host and command line (mmc.c, mmc_host.c, mmc_mesh.c, mmc_utils.c, ...)
linked against integration/mmc_cu_host_b200.cpp + libmmc_b200.so instead of src/mmc_cu_host.cu
(oracle/Makefile.ref target `b200cli` -> oracle/_ref/mmc_b200cli).  `-c cuda` then runs the B200 engine behind
mmc_run_cu(); `-c sse` in the same binary is the reference's own CPU path."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "oracle", "_ref", "mmc_b200cli")
needs_cli = pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/mmc_b200cli not built (needs /root/reference at build time)")


def _run(args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    return subprocess.run([CLI] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd="/tmp")


def _absorbed(out):
    m = re.findall(r"absorbed:\s*(?:\x1b\[[0-9;]*m)*\s*([0-9.]+)%", out)
    assert m, out[-2000:]
    return float(m[-1]) / 100.0


@needs_cli
def test_cli_fails_with_the_reference_error_convention_without_a_gpu():
    import mmc_b200
    if mmc_b200.gpuinfo():
        pytest.skip("a GPU is present")
    r = _run(["--bench", "dmmc-cube60", "-c", "cuda", "-n", "1000", "-D", "T", "-S", "0"])
    assert r.returncode != 0
    assert "MMC ERROR(-1):No GPU device found" in (r.stdout + r.stderr)


@needs_cli
@pytest.mark.gpu
@pytest.mark.parametrize("bench", ["dmmc-cube60", "dmmc-cube60b"])
def test_reference_cli_drives_the_b200_engine(bench):
    """Built-in benchmarks of the reference binary (src/mmc_bench.c:41-110; the CI smoke test of the reference,
    .github/workflows/build_all.yml:151-160): same command line, `-c cuda` (our engine) vs `-c sse` (reference CPU)."""
    gpu = _run(["--bench", bench, "-c", "cuda", "-n", "1e6", "-D", "T", "-S", "0"])
    assert gpu.returncode == 0, gpu.stdout[-2000:] + gpu.stderr[-2000:]
    assert "MMC-B200" in gpu.stdout and "MCX simulation speed" in gpu.stdout
    cpu = _run(["--bench", bench, "-c", "sse", "-n", "1e6", "-D", "T", "-S", "0"])
    assert cpu.returncode == 0, cpu.stdout[-2000:] + cpu.stderr[-2000:]
    fg, fc = _absorbed(gpu.stdout), _absorbed(cpu.stdout)
    assert abs(fg - fc) < 2.5e-3, (fg, fc)          # 1e6 photons: sigma ~ 4e-4

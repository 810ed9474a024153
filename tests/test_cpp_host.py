"""The compiled host (host/mcbrat_host.hpp + examples/i3rc_driver.cpp, C++17 straight above the C ABI) against the
Python host mirror: same decks, same seeds.  CPU: it builds and fails loudly without a GPU.  GPU: its batch
statistics agree with the Python path within the combined Monte Carlo error."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "i3rc_driver")


def _build():
    csrc = os.path.join(ROOT, "mcbrat3d_b200", "csrc")
    if not os.path.exists(os.path.join(csrc, "libmcbrat_cuda.so")):
        subprocess.check_call(["make", "-C", csrc], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", EXE, os.path.join(ROOT, "examples", "i3rc_driver.cpp"),
                           "-L" + csrc, "-lmcbrat_cuda", "-Wl,-rpath,$ORIGIN/../mcbrat3d_b200/csrc"])


def test_compiled_host_builds_and_has_no_cpu_fallback():
    import torch
    _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([EXE, "homog", "2", "1000"], capture_output=True, text=True)
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("deck,views", [("homog", 0), ("stepcloud", 0), ("stepcloud", 1)])
def test_compiled_host_matches_python_host(deck, views):
    from mcbrat3d_b200 import domains
    from mcbrat3d_b200.batchStatistics import computeRadiativeTransferBatches, reportStatistics, resetDeviceStatistics
    from mcbrat3d_b200.monteCarloIllumination import new_PhotonStream
    from mcbrat3d_b200.monteCarloRadiativeTransfer import finalize_Integrator, new_Integrator, specifyParameters
    from mcbrat3d_b200.RandomNumbersForMC import new_RandomNumberSequence
    _build()
    nb, n = 24, 40000
    out = subprocess.run([EXE, deck, str(nb), str(n), "10", str(views)], capture_output=True, text=True, check=True).stdout
    m = re.search(r"first batch: (\d+) photons\s+meanFluxUp (\S+) meanFluxDown (\S+) meanFluxAbsorbed (\S+)", out)
    assert int(m.group(1)) == n
    cpp = {k: tuple(float(v) for v in re.search(k + r"\s+(\S+) \+- (\S+)", out).groups())
           for k in ("Flux Up", "Flux Down", "Flux Absorbed")}
    assert "batches %d photons %d" % (nb, nb * n) in out
    dom, case = domains.homogeneous_slab(ssa=0.99) if deck == "homog" else domains.step_cloud(ssa=0.99, solarMu=0.5)
    g = new_Integrator(dom)
    try:
        if views:
            specifyParameters(g, intensityMus=case["intensityMus"], intensityPhis=case["intensityPhis"], computeIntensity=True,
                              useRussianRouletteForIntensity=True, zetaMin=0.3)
        specifyParameters(g, minInverseTableSize=10001)
        rs = new_RandomNumberSequence([10, 1, 0])           # the driver's batch loop starts at photon id 0 as well
        ps = new_PhotonStream(0.5, 0.0, nb * n, rs)
        resetDeviceStatistics(g, dom)
        computeRadiativeTransferBatches(g, dom, rs, ps, n, nb)
        mean, err, tot, done = reportStatistics(g)
    finally:
        finalize_Integrator(g)
    for k, q in (("Flux Up", "meanFluxUp"), ("Flux Down", "meanFluxDown"), ("Flux Absorbed", "meanFluxAbsorbed")):
        sig = np.hypot(cpp[k][1], float(err[q]))
        assert abs(cpp[k][0] - float(mean[q])) <= 4.0 * sig + 1e-6, (k, cpp[k], float(mean[q]), float(err[q]))
    albedo = 0.2 if deck == "homog" else 0.0
    assert abs(cpp["Flux Up"][0] + (1 - albedo) * cpp["Flux Down"][0] + cpp["Flux Absorbed"][0] - 1.0) < 4e-3
    if views:
        rad = [float(v) for v in re.findall(r"Radiance view \d+\s+(\S+)", out)]
        want = mean["intensity"].reshape(5, -1).mean(axis=1)
        assert len(rad) == 5 and np.allclose(rad, want, rtol=0.05)


@pytest.mark.gpu
def test_compiled_host_writes_the_ascii_tables(tmp_path):
    """--out PREFIX: the driver's four ASCII tables (writeResults_ASCII, DRV:1324-1495) from device-side statistics."""
    _build()
    prefix = str(tmp_path / "step")
    subprocess.run([EXE, "stepcloud", "8", "20000", "10", "1", "--out", prefix], capture_output=True, text=True, check=True)
    flux = open(prefix + "_flux.out").read().split("\n")
    assert flux[0] == "!   I3RC Monte Carlo 3D Solar Radiative Transfer: Flux"
    assert flux[2] == "!  Num_Photons=    160000" and flux[11].startswith("!  Average:   ")
    rows = [l for l in flux if l and not l.startswith("!")]
    assert len(rows) == 32 and all(len(l) == 14 + 3 * 21 for l in rows)
    up = np.array([float(l[14:25]) for l in rows])
    avg = float(flux[11][14:25])
    assert abs(up.mean() - avg) < 2e-4                      # the average line is the mean of the pixel column
    rad = open(prefix + "_rad.out").read().split("\n")
    assert sum(1 for l in rad if l.endswith("<- (mu,phi)")) == 5
    assert "NXO=  32   NYO=   1   NDIR=   5" in rad[10] and rad[11].startswith("!   X      Y")
    prof = [l for l in open(prefix + "_absprof.out").read().split("\n") if l and not l.startswith("!")]
    assert len(prof) == 32
    vol = [l for l in open(prefix + "_absvol.out").read().split("\n") if l and not l.startswith("!")]
    assert len(vol) == 32 * 32


@pytest.mark.gpu
def test_compiled_host_ranks_reduce_through_the_c_abi():
    """--ranks 2: one process per GPU, mcb_comm_init + ONE mcb_reduce_statistics (ncclReduce) -- the same batches as
    the single-process run (same global photon ids), so the means agree to f64 summation order."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _build()
    args = ["stepcloud", "16", "50000", "10", "0"]
    one = subprocess.run([EXE] + args, capture_output=True, text=True, check=True).stdout
    two = subprocess.run([EXE] + args + ["--ranks", "2"], capture_output=True, text=True, check=True, timeout=300).stdout
    assert "ranks 2 batches 16 photons 800000" in two and "ranks 1 batches 16 photons 800000" in one

    def means(text):
        m = re.search(r"means (\S+) (\S+) (\S+)  errors (\S+) (\S+) (\S+)", text)
        return np.array([float(v) for v in m.groups()])
    np.testing.assert_allclose(means(two)[:3], means(one)[:3], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(means(two)[3:], means(one)[3:], rtol=1e-5, atol=1e-12)     # sqrt(m2 - m1^2): cancellation

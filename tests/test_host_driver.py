"""GPU tests of the reference-side binding: block2's own two-site DMRG driver with
b2g_host::install(mpo) (block2-preview_b200/host/b2g_adapter.hpp), run as the prebuilt
binary block2-preview_b200/host/_build/b2g_dmrg_* (built by __graft_entry__.build() where the
reference tree exists; the binary and its FCIDUMP inputs travel to the GPU box).

Bars (BASELINE.json north_star): H.C <= 1e-11 relative against the reference's CPU executor
on the live H_eff of every site; converged energy within 1e-8 Ha of the reference's stored values
(unit_test/test_dmrg_n2_sto3g.cpp:191, unit_test/test_rotation_h10_sto6g.cpp:43)."""
import json
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
BUILD = os.path.join(ROOT, "block2-preview_b200", "host", "_build")
E_N2_1AG = -107.654122447525      # reference golden, SU2 singlet Ag
E_H10 = -5.424385375684663        # reference golden, H10 STO-6G R=1.8


def run_driver(exe, *args):
    path = os.path.join(BUILD, exe)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (needs the reference tree at build time)")
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    out = subprocess.run([path, *args, "--scratch", "/tmp/b2g_test_scratch"], env=env, capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.parametrize("davidson", ["device", "host"])
def test_n2_sto3g_su2_energy_and_matvec_parity(davidson):
    r = run_driver("b2g_dmrg_su2", "--fcidump", os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "250",
                   "--nsweeps", "8", "--threads", "4", "--noise", "1e-6", "--davidson", davidson, "--verify")
    assert abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r
    assert r["matvec_sites_verified"] > 0 or davidson == "host"
    assert r["max_matvec_rel_err"] < 1e-11, r
    assert r["launches"] > 0


def test_h10_sto6g_sz_energy_matches_reference_run():
    r = run_driver("b2g_dmrg_sz", "--fcidump", os.path.join(BUILD, "data", "H10.STO6G.R1.8.FCIDUMP"), "--bond",
                   "500", "--nsweeps", "8", "--threads", "8", "--noise", "1e-6", "--compare", "--verify")
    assert abs(r["e_gpu"] - r["e_ref"]) < 1e-8, r          # same run, CPU path first
    assert abs(r["e_gpu"] - E_H10) < 1e-7, r               # the reference's own tolerance for this value
    assert r["max_matvec_rel_err"] < 1e-11, r


def test_renormalisation_on_device_matches_reference_executor():
    """--gpu-rotate: left_rotate / right_rotate lists (OperatorFunctions::tensor_rotate ->
    BatchGEMMSeq::rotate) run through b2g_pairs_execute; --verify replays every list with the
    reference's own auto_perform() and compares all rotated operator blocks."""
    r = run_driver("b2g_dmrg_su2", "--fcidump", os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "250",
                   "--nsweeps", "8", "--threads", "4", "--noise", "1e-6", "--gpu-rotate", "--verify")
    assert r["rotations"] > 0 and r["max_rotate_rel_err"] < 1e-11, r
    assert abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r
    r = run_driver("b2g_dmrg_sz", "--fcidump", os.path.join(BUILD, "data", "H10.STO6G.R1.8.FCIDUMP"), "--bond",
                   "300", "--nsweeps", "6", "--threads", "8", "--noise", "1e-6", "--gpu-rotate")
    assert r["rotations"] > 0 and abs(r["e_gpu"] - E_H10) < 1e-6, r


def test_blocking_on_device_matches_reference_executor():
    """--gpu-contract: left_contract / right_contract recorded by the reference's own walker / OperatorFunctions
    and executed by b2g_tensor_product_execute (resident blocks feeding the rotation list and the H.C plan);
    --verify re-records every call with the reference's recorder, runs its auto_perform() and compares all
    blocked operators, and does the same for every rotation list and H.C list."""
    r = run_driver("b2g_dmrg_su2", "--fcidump", os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "250",
                   "--nsweeps", "8", "--threads", "4", "--noise", "1e-6", "--gpu-contract", "--gpu-rotate", "--verify")
    assert r["contractions"] > 0 and r["max_contract_rel_err"] < 1e-11, r
    assert r["rotations"] > 0 and r["max_rotate_rel_err"] < 1e-11, r
    assert r["max_matvec_rel_err"] < 1e-11 and r["resident_hit_gbytes"] > 0, r
    assert abs(r["e_gpu"] - E_N2_1AG) < 1e-8, r
    r = run_driver("b2g_dmrg_sz", "--fcidump", os.path.join(BUILD, "data", "H10.STO6G.R1.8.FCIDUMP"), "--bond",
                   "300", "--nsweeps", "6", "--threads", "8", "--noise", "1e-6", "--gpu-contract", "--gpu-rotate",
                   "--verify")
    assert r["contractions"] > 0 and r["max_contract_rel_err"] < 1e-11 and r["max_rotate_rel_err"] < 1e-11, r
    assert abs(r["e_gpu"] - E_H10) < 1e-6, r


def test_blocking_on_device_zero_fill_route(monkeypatch):
    """B2G_ZERO_OUTPUTS=1: zero-initialised outputs + add (B2G_DST_ZERO without B2G_DST_COVERED)."""
    monkeypatch.setenv("B2G_ZERO_OUTPUTS", "1")
    r = run_driver("b2g_dmrg_su2", "--fcidump", os.path.join(BUILD, "data", "N2.STO3G.FCIDUMP"), "--bond", "120",
                   "--nsweeps", "4", "--threads", "4", "--noise", "1e-6", "--gpu-contract", "--gpu-rotate", "--verify")
    assert r["contractions"] > 0 and r["max_contract_rel_err"] < 1e-11 and r["max_rotate_rel_err"] < 1e-11, r


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

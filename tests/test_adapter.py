"""The C++ adapter (dftpav_b200/host/poly_traj_optimizer.hpp) keeps the reference's optimizer API
(setParam / OptimizeTrajectory / getMinJerkOptPtr) on top of the C-ABI.  It is compiled here in standalone mode
(Eigen is not installed: tests/mini_matrix.hpp stands in for MatrixXd/VectorXd) and, on the GPU box, driven like
TrajPlanner::RunMINCOParking drives the reference (traj_manager.cpp:509-625)."""
import os
import re
import subprocess

import numpy as np
import pytest

from dftpav_b200 import capi, synth
from dftpav_b200.batch import samples_per_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "dftpav_b200", "csrc")


def _build(tmp_path):
    exe = str(tmp_path / "adapter_main")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-o", exe,
           os.path.join(ROOT, "tests", "adapter_main.cpp"), "-L" + CSRC, "-ldftpav_b200", "-Wl,-rpath," + CSRC]
    subprocess.check_call(cmd, cwd=os.path.join(ROOT, "tests"))
    return exe


def _dump(batch, prm, p, path):
    K, Kd = prm.traj_resolution, prm.des_traj_resolution
    g0, g1 = int(batch.seg_ptr[p]), int(batch.seg_ptr[p + 1])
    Ns = batch.piece_num
    inner_pref = np.concatenate([[0], np.cumsum(2 * (Ns - 1))])
    samp = np.array([samples_per_segment(int(N), K, Kd) for N in Ns])
    samp_pref = np.concatenate([[0], np.cumsum(samp)])
    H = batch.uniform_H
    with open(path, "w") as f:
        f.write(f"{g1 - g0} {K} {Kd} {prm.max_forward_vel!r} {prm.max_forward_acc!r} {prm.max_forward_cur!r}\n")
        for g in range(g0, g1):
            N, M = int(Ns[g]), int(samp[g])
            f.write(f"{N} {int(batch.singul[g])} {float(batch.init_T[g])!r} {M}\n")
            f.write(" ".join(repr(float(v)) for v in batch.ini_state[g]) + "\n")
            f.write(" ".join(repr(float(v)) for v in batch.fin_state[g]) + "\n")
            f.write(" ".join(repr(float(v)) for v in batch.inner_pts[inner_pref[g]:inner_pref[g + 1]]) + "\n")
            for q in range(int(samp_pref[g]), int(samp_pref[g + 1])):
                pl = batch.planes[q * H:(q + 1) * H].reshape(-1)
                f.write(f"{H} " + " ".join(repr(float(v)) for v in pl) + "\n")


def test_adapter_compiles_and_links(tmp_path):
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_adapter_matches_batch_path(tmp_path, oracle, gpu_opt_factory):
    exe = _build(tmp_path)
    for cfg in (3, 4):
        batch, prm = synth.make_config(cfg, B=2, seed=8)
        path = str(tmp_path / f"problem{cfg}.txt")
        _dump(batch, prm, 1, path)
        res = subprocess.run([exe, path], capture_output=True, text=True)
        m = re.search(r"success (\d) status (-?\d+) cost (\S+)", res.stdout)
        assert m, res.stdout + res.stderr
        opt = gpu_opt_factory(prm)
        sol = opt.solve(batch)
        # same library, same arithmetic: identical to the batch path
        assert int(m.group(1)) == int(sol.success[1]) and int(m.group(2)) == int(sol.status[1])
        assert float(m.group(3)) == float(sol.cost[1])
        segs = re.findall(r"segment (\d+) pieces (\d+) dT (\S+)", res.stdout)
        g0 = int(batch.seg_ptr[1])
        for s, N, dT in segs:
            assert int(N) == int(batch.piece_num[g0 + int(s)])
            assert float(dT) == float(sol.piece_dT[g0 + int(s)])


REF_INC = "/root/reference/src/Plan/traj_planner/include"


@pytest.mark.parametrize("mode", ["reference", "dropin"])
def test_plan_manager_call_site_compiles_against_reference_header_and_dropin(mode, tmp_path):
    """tests/dropin_callsite.cpp is the optimizer call sequence of the reference's plan manager
    (traj_manager.cpp:29, 604-623).  The SAME text must compile against the reference's own
    plan_manage/traj_optimizer.h and against this repository's drop-in header in reference-types mode (Eigen / ROS /
    protobuf come from the test-only layers under oracle/)."""
    if not os.path.isdir(REF_INC):
        pytest.skip("reference tree not present")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-c", "-w", "-I", os.path.join(ROOT, "oracle", "ref_stubs"), "-I",
           os.path.join(ROOT, "oracle", "eigen_shim2"), "-I", REF_INC, os.path.join(ROOT, "tests", "dropin_callsite.cpp"),
           "-o", str(tmp_path / "callsite.o")]
    if mode == "reference":
        cmd.insert(1, "-DUSE_REFERENCE_OPTIMIZER")
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]

"""The C++20 host shim (include/openfly_b200/fly.hpp) -- the reference's class surface for a compiled host.
CPU: the header compiles with -std=c++20, links against the C-ABI library and its host-only calls work.
GPU: examples/shim_demo.cpp (List / Generic / LBFGS / Dimer through the shim) against the CPU oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.common import ROOT, WEN, bcc_vacancy

DEMO = os.path.join(ROOT, "examples", "shim_demo")
MGPU = os.path.join(ROOT, "examples", "multi_gpu_batch")


@pytest.fixture(scope="module")
def demo():
    from openfly_b200 import capi

    capi.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)
    return DEMO


def test_shim_compiles_links_and_runs_host_only(demo):
    out = subprocess.run([demo, "--host-only", WEN], check=True, capture_output=True, text=True).stdout
    d = json.loads(out)
    assert d["version"].startswith("openfly_b200")
    assert (d["lbfgs_n"], d["dimer_n"], d["rotor_n"], d["types"], d["knots_r"], d["r_cut"]) == (10, 10, 6, 2, 10004, 5.3)


def test_shim_reports_errors_as_exceptions(demo):
    res = subprocess.run([demo, "--host-only", "/nonexistent.eam.fs"], capture_output=True, text=True)
    assert res.returncode == 1 and "error:" in res.stderr


@pytest.mark.gpu
def test_shim_demo_matches_oracle(demo, oracle_lib, tmp_path):
    cells, R = 4, 4
    base, ex, ids = bcc_vacancy(cells, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    tab = oracle_lib.Tables(WEN)
    osys = oracle_lib.System(tab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=4)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(R)]
    path = tmp_path / "starts.bin"
    np.stack([np.stack([s[0], s[1]]) for s in starts]).astype(np.float64).tofile(path)

    out = subprocess.run([demo, WEN, str(cells), str(R), str(path)], check=True, capture_output=True, text=True).stdout
    d = json.loads(out)
    assert d["atoms"] == len(ids) and d["launches"] > 0

    osys.rebuild(tab.rcut, base)
    E = osys.energy(ids)
    assert abs(d["energy"] - E) <= 1e-10 * abs(E)
    assert abs(d["grad_max"] - np.abs(osys.gradient(ids, fz)).max()) < 1e-9
    assert d["neigh0"] == len(osys.for_neighbours(0, tab.rcut)[0])
    assert d["min_failed"] == 0
    osys.rebuild(tab.rcut, r0)
    assert abs(d["energy_min"] - osys.energy(ids)) < 1e-9
    res = np.fromfile(str(path) + ".out", dtype=np.float64)
    n3 = 3 * len(ids)
    assert np.abs(res[:n3].reshape(-1, 3) - r0).max() < 1e-6
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    o_r, _, o_code, _ = osys.find_sp_batch(P, A, ids, fz, r0, threads=4)
    assert d["exits"] == o_code.tolist()
    sp = res[n3:].reshape(R, -1, 3)
    for k in range(R):
        if o_code[k] == 0:
            assert np.abs(sp[k] - o_r[k]).max() < 1e-5


@pytest.mark.gpu
def test_cpp_host_shards_replicas_over_devices(demo, oracle_lib, tmp_path):
    """examples/multi_gpu_batch.cpp: one host thread + Context per device, replica r on device r mod G, merge in replica
    order. With every visible device (one is enough to run it) the merged result equals the single-device one."""
    import torch

    cells, R = 4, 6
    base, ex, ids = bcc_vacancy(cells, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    tab = oracle_lib.Tables(WEN)
    osys = oracle_lib.System(tab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=4)
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(R)]
    path = tmp_path / "starts.bin"
    np.stack([np.stack([s[0], s[1]]) for s in starts]).astype(np.float64).tofile(path)
    runs = {}
    for g in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        out = subprocess.run([MGPU, WEN, str(cells), str(R), str(path), str(g)], check=True, capture_output=True, text=True).stdout
        d = json.loads(out)
        assert d["devices"] == g and d["replicas"] == R
        runs[g] = (d["exits"], np.fromfile(str(path) + ".mgpu", dtype=np.float64))
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    o_r, _, o_code, _ = osys.find_sp_batch(P, A, ids, fz, r0, threads=4)
    for g, (exits, sp) in runs.items():
        assert exits == o_code.tolist()
        sp = sp.reshape(R, -1, 3)
        for k in range(R):
            if o_code[k] == 0:
                assert np.abs(sp[k] - o_r[k]).max() < 1e-5


SURF = os.path.join(ROOT, "examples", "libfly_surface")


def test_libfly_surface_compiles_with_reference_signatures(demo):
    """include/openfly_b200/libfly.hpp: namespace fly, the reference's own signatures over system::SoA views
    (generic.hpp:79-136, list.hpp:86-161, lbfgs.hpp:96-118, dimer.hpp:148-155). The forwarding bodies INTEGRATION.md quotes are
    the ones compiled here."""
    d = json.loads(subprocess.run([SURF, "--compile-only"], check=True, capture_output=True, text=True).stdout)
    assert d == {"lbfgs_n": 10, "dimer_n": 10, "rotor_n": 6, "view_sees_owner": 1}
    text = open(os.path.join(ROOT, "include", "openfly_b200", "libfly.hpp")).read()
    for sig in ("auto energy(system::SoA<TypeID const &, Frozen const &> in, neigh::List const &nl, int /*threads*/ = 1) -> double",
                "auto rebuild(system::SoA<Position const &> positions, int num_threads = 1) -> void",
                "auto update(system::SoA<Delta const &> x) -> void",
                "auto minimise(system::SoA<Position &, PotentialGradient &> out, system::SoA<Position const &, TypeID const &, Frozen const &> in,",
                "auto find_sp(system::SoA<Position &, Axis &> out, system::SoA<Position const &, Axis const &, TypeID const &, Frozen const &> in,"):
        assert sig in text, sig


@pytest.mark.gpu
def test_libfly_surface_matches_oracle(demo, oracle_lib, tmp_path):
    """A libfly-style program through the reference signatures: energy / gradient / Hessian / for_neighbours / minimise /
    find_sp (single and batched) against the CPU oracle; two Lists do not disturb each other."""
    cells, R = 4, 4
    base, ex, ids = bcc_vacancy(cells, a=2.855, remove=(1,), hydrogen=True)
    fz = np.zeros(len(ids), np.uint8)
    tab = oracle_lib.Tables(WEN)
    osys = oracle_lib.System(tab, ex)
    r0, _, failed, _ = osys.minimise(base, ids, fz, threads=4)
    assert not failed
    starts = [osys.perturb(r0, fz, len(ids) - 1, (1, 3, 3, 4), rep) for rep in range(R)]
    path = tmp_path / "starts.bin"
    np.stack([np.stack([s[0], s[1]]) for s in starts]).astype(np.float64).tofile(path)
    res = subprocess.run([SURF, WEN, str(cells), str(R), str(path)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    d = json.loads(res.stdout)
    osys.rebuild(tab.rcut, base)
    E = osys.energy(ids)
    assert d["atoms"] == len(ids)
    assert abs(d["energy"] - E) <= 1e-10 * abs(E)
    assert abs(d["grad_max"] - np.abs(osys.gradient(ids, fz)).max()) < 1e-9
    counts = osys.export()[0]
    assert d["neigh0"] == len(osys.for_neighbours(0, tab.rcut)[0]) and d["pairs"] == int(counts.sum())
    H = osys.hessian(ids, fz, 4)
    assert abs(d["h00"] - H[0, 0]) < 1e-9 and abs(d["h30"] - H[3, 0]) < 1e-9
    assert d["energy_again"] == d["energy"] and d["energy_other_list"] != d["energy"]  # Lists own their state
    assert d["energy_updated"] != d["energy"] and abs(d["energy_back"] - d["energy"]) < 1e-9
    assert d["min_failed"] == 0
    osys.rebuild(tab.rcut, r0)
    assert abs(d["energy_min"] - osys.energy(ids)) < 1e-9
    P, A = np.stack([s[0] for s in starts]), np.stack([s[1] for s in starts])
    # the program searches from ITS relaxed minimum (1e-6 A from the oracle's): same exits, same saddles
    o_r, _, o_code, _ = osys.find_sp_batch(P, A, ids, fz, r0, threads=4)
    assert d["exits"] == o_code.tolist() and d["exit0"] == o_code[0]
    assert d["single_vs_batch_sp"] == 0.0  # the single-system signature and the batch give the same bits
    out = np.fromfile(str(path) + ".fly", dtype=np.float64)
    n3 = 3 * len(ids)
    assert np.abs(out[:n3].reshape(-1, 3) - r0).max() < 1e-6
    sp = out[n3:].reshape(R, -1, 3)
    for k in range(R):
        if o_code[k] == 0:
            assert np.abs(sp[k] - o_r[k]).max() < 1e-5

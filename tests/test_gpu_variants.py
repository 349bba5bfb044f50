"""Kernel variants of the hot path agree: the scheduled, batched and thread-per-row volume assembly, the real and the
complex chunk image, and the TMA-streamed and register-streamed CSR SpMV (each selected per process by an environment
variable, so every variant runs in its own subprocess)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, name, env_extra):
    out = str(tmp_path / (name + ".npz"))
    env = dict(os.environ)
    for k in ("EDGEFEM_B200_ASM_KERNEL", "EDGEFEM_B200_ASM_NO_REAL", "EDGEFEM_B200_SPMV_KERNEL"):
        env.pop(k, None)
    env.update(env_extra)
    r = subprocess.run([sys.executable, os.path.join(HERE, "variant_probe.py"), out], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out)


def test_assembly_and_spmv_variants_agree(tmp_path):
    ref = _run(tmp_path, "default", {})
    batch = _run(tmp_path, "batch", {"EDGEFEM_B200_ASM_KERNEL": "batch"})
    cplx = _run(tmp_path, "cplx", {"EDGEFEM_B200_ASM_NO_REAL": "1", "EDGEFEM_B200_SPMV_KERNEL": "regs"})
    row = _run(tmp_path, "row", {"EDGEFEM_B200_ASM_KERNEL": "row"})
    for other in (batch, cplx, row):
        assert np.array_equal(ref["rowptr"], other["rowptr"]) and np.array_equal(ref["colidx"], other["colidx"])
    scale = np.max(np.abs(ref["vals_lossy"]))
    for key in ("vals_real", "vals_lossy"):
        # same per-incidence arithmetic and the same ascending-tet summation order: bit-identical
        assert np.array_equal(ref[key], batch[key]), key
        assert np.array_equal(ref[key], cplx[key]), key
        # the first-generation kernel forms V/20 by a division and leaves the fma contraction to the compiler
        assert np.max(np.abs(ref[key] - row[key])) <= 1e-14 * scale, key
    assert np.all(ref["vals_real"].imag == 0.0) and np.any(ref["vals_lossy"].imag != 0.0)
    for key in ("y_real", "y_lossy"):
        # same chunks, products and in-order row sums in both SpMV kernels
        ys = np.max(np.abs(ref[key]))
        assert np.max(np.abs(ref[key] - cplx[key])) <= 1e-15 * ys, key


def test_resolve_with_single_rhs_tail_jobs(wr90):
    """A re-solve of a resident sweep system (more matrices than SMs, iteration history present) queues the matrices
    expected to finish last as single-rhs jobs (run_cocg_small, mixed queue): same S-parameters as the first solve,
    which runs every matrix as one two-rhs job."""
    import edgefem_oracle as orc
    import helpers as H
    from edgefem_b200 import cabi

    mesh, pec = wr90
    ctx = cabi.Ctx(0)
    freqs = np.linspace(8e9, 12e9, 160)
    ports = orc.wr90_ports(mesh, pec, 10e9)
    keep = {}
    S1, res1 = H.eigenmode_sweep_gpu(ctx, mesh, pec, ports, freqs, keep=keep)
    assert all(r["converged"] for r in res1)
    sysd, dports = keep["sys"], keep["ports"]
    res2 = sysd.solve(method=cabi.METHOD_AUTO, precond=cabi.PRECOND_AUX, tol=1e-10, symmetric=True)
    assert all(r["converged"] for r in res2)
    F, P = len(freqs), len(ports)
    S2 = np.zeros_like(S1)
    for fi in range(F):
        for a in range(P):
            for j in range(P):
                v = dports[j].project_mass(fi * P + a)
                S2[fi, j, a] = v - 1.0 if j == a else v
    assert np.max(np.abs(S2 - S1)) <= 1e-7
    it1 = np.array([r["iters"] for r in res1]); it2 = np.array([r["iters"] for r in res2])
    # same Krylov process per right-hand side; the one-rhs instantiation rounds differently (fma contraction), and COCG's
    # residual hovers around the tolerance at the end, so single systems stop up to a few percent earlier or later
    # (measured: most identical, a few +-20 of ~330; one run in six of the full suite saw a system beyond 15 % -- a restart
    # cycle taken by one instantiation and not by the other -- so the bounds leave room for that)
    assert np.max(np.abs(it1 - it2)) <= 0.35 * np.max(it1)
    assert abs(int(it1.sum()) - int(it2.sum())) <= 0.06 * it1.sum()
    # spot check against the oracle
    for fi in (0, 80, 159):
        S_ref = orc.wr90_sparams(mesh, pec, freqs[fi], ports)
        assert np.max(np.abs(S2[fi] - S_ref)) <= 1e-6
    for dp in dports:
        dp.close()
    sysd.close()
    keep["mesh"].close()
    ctx.close()

#!/usr/bin/env python
"""bench.py -- EdgeFEM hot path on B200: frequency points/s on the WR-90 sweep (8-12 GHz, 256
points, BASELINE.json configs[1]), plus assembly Mtets/s and SpMV HBM GB/s on a synthetic cube.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path over one batch: the full 256-point, 2-port eigenmode sweep
(512 complex sparse solves) -- volume assembly, port terms, right-hand sides, batched Krylov
solve, S-parameter projection.
  value : points/s with the mesh, CSR pattern and port operators already resident in HBM
  e2e   : points/s through the public API (pyedgefem.calculate_sparams_eigenmode_sweep) from host
          buffers: mesh upload, pattern build, port upload and S read-back inside the timed region
Multi-GPU (torchrun, one rank per GPU): the sweep shards by frequency with no data-path collective.
Default "strong" scaling = BASELINE.json configs[1] as written: the 256 points are dealt round-robin to the ranks
(freqs[rank::world]) and the P x P S-matrices are all-gathered (NCCL) INSIDE the timed region of every step;
--scaling weak gives every rank its own 256-point sub-band instead.
Other named workloads: --workload c1 | patch | unitcell | cube print their own JSON line (profiles/ keeps one each).
Timing: CUDA events on the library's stream (efb_timer_*), barrier + synchronize on both sides,
max over ranks.  Working set per step (349 MB of matrix values + 330 MB of Krylov vectors) is
larger than the 126 MB L2, so no explicit L2 flush is needed between timed iterations; shards of fewer than 64 points
(N = 8) would fit, so there the L2 is flushed (efb_l2_flush, untimed) before every timed step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C0 = 299792458.0
WR90_A, WR90_B = 0.02286, 0.01016
F_LO, F_HI, N_POINTS = 8e9, 12e9, 256


_LINE = []  # the one JSON line of this process (rank 0 only), printed by main() after stdout is restored

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--cube-n", type=int, default=150, help="synthetic cube (n^3 boxes x 6 tets) of the C5 extras: assembly / SpMV roofline and a converged "
                    "solve; 150 = BASELINE configs[4] (20.25 M tets); 0 = skip")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--workload", default="wr90", choices=["wr90", "c1", "patch", "unitcell", "cube"])
    ap.add_argument("--cube-freq", type=float, default=240e6, help="frequency of the C5 solve (between the 212.0 and 259.6 MHz modes of the 1 m cube)")
    ap.add_argument("--cpu-sample", type=int, default=12, help="frequency points of the CPU baseline sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, device: int):
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                parts = [x.strip() for x in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------ reference arm (CPU oracle)
def _oracle_points(args):
    """Worker: eigenmode S-parameters of a list of frequencies with the CPU oracle (SuperLU)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import edgefem_oracle as orc

    freqs, = args
    z = np.load(os.path.join(ROOT, "tests", "golden", "rect_waveguide.npz"))
    mesh = orc.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], node_ids=z["node_ids"])
    pec = orc.build_edge_pec(mesh, 1)
    ports = orc.wr90_ports(mesh, pec, 10e9)
    t0 = time.perf_counter()
    out = [orc.wr90_sparams(mesh, pec, f, ports) for f in freqs]
    return time.perf_counter() - t0, len(out)


def cpu_baseline(n_points: int, cores: int):
    """points/s of the oracle (restated reference path: per-frequency triplet assembly + SuperLU solves)
    on `cores` host processes; setup (mesh, port modes) excluded like the GPU arm's resident number."""
    import numpy as np

    freqs = list(np.linspace(F_LO, F_HI, n_points))
    import multiprocessing as mp

    # one BLAS/OpenMP thread per worker process: `cores` is then the number of threads really used
    # (and 8 workers x 8 spinning OpenBLAS threads on 8 cores would otherwise livelock)
    pinned = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in pinned:
        os.environ[k] = "1"
    try:
        chunks = [freqs[i::cores] for i in range(cores)]
        chunks = [c for c in chunks if c]
        with mp.get_context("spawn").Pool(len(chunks)) as pool:
            res = pool.map(_oracle_points, [(c,) for c in chunks])
    finally:
        for k, v in pinned.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    # worker start-up (imports, mesh, port eigen-solves) is excluded: use the slowest worker's compute time
    slowest = max(r[0] for r in res)
    return sum(r[1] for r in res) / slowest


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = max(cores, min(2 * cores, 64))
    vals = []
    for _ in range(a.warmup):
        cpu_baseline(sample, cores)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        vals.append(cpu_baseline(sample, cores))
    wall = time.perf_counter() - t0
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "wr90_sweep_freq_points_per_s", "value": v, "unit": "points/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * wall / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "WR-90 eigenmode sweep 8-12 GHz (rect_waveguide fixture, 4227 tets, 2 ports), bounded sample of %d points per step" % sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": "%d of 256 frequency points per step, %d processes; oracle = numpy/scipy restatement (Eigen is not installable here), "
                                   "per-frequency assembly + SuperLU factorisation + 2 solves" % (sample, cores)},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _LINE.append(json.dumps(line))


# ------------------------------------------------------------------------------------------ B200 arm
def mesh_arrays(hm):
    return dict(xyz=hm.xyz_array(), tet_nodes=hm.tet_nodes_array(), tet_edges=hm.tet_edges_array(), tet_orient=hm.tet_orient_array(),
                tet_phys=hm.tet_phys_array(), edge_nodes=hm.edge_nodes_array())


class ResidentSweep:
    """The hot path with every input already in HBM: one call = one full sweep step."""

    def __init__(self, ctx, pe, hm, bc, ports, freqs):
        import numpy as np
        from edgefem_b200 import cabi

        self.np, self.cabi, self.ctx = np, cabi, ctx
        arr = mesh_arrays(hm)
        self.dm = cabi.DeviceMesh(ctx, arr["xyz"], arr["tet_nodes"], arr["tet_edges"], arr["tet_orient"], arr["tet_phys"], arr["edge_nodes"])
        flags = np.zeros(hm.num_edges(), dtype=np.uint8)
        flags[np.asarray(bc.dirichlet_edges, dtype=np.int64)] = 1
        pe_idx = np.nonzero(flags)[0].astype(np.int32)
        self.F, self.P = len(freqs), len(ports)
        self.sys = cabi.DeviceSystem.from_mesh(self.dm, pe_idx, pe_idx, n_matrix=self.F, n_rhs=self.P)
        self.sys.set_dirichlet(flags)
        self.mats, self._keep = cabi.make_materials(len(self.dm.slot_tags))
        self.omegas = np.array([2 * math.pi * f for f in freqs])
        self.dports = []
        for port in ports:
            rp, ci, va = pe.assemble_port_surface_mass(hm, port.surface_tag, set(bc.dirichlet_edges)).to_csr()
            rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp)).astype(np.int32)
            dp = cabi.DevicePort(self.sys, port.edges, port.weights, rows, ci, va)
            dp.normalize_mass()
            self.dports.append(dp)
        k0 = self.omegas / C0
        self.betas = np.stack([np.sqrt((k0 * k0 - p.mode.kc ** 2).astype(complex)) for p in ports], axis=1)  # [F,P] vacuum ports
        self.idx = [np.arange(self.F, dtype=np.int32) * self.P + a for a in range(self.P)]
        self.all_idx = np.arange(self.F * self.P, dtype=np.int32)
        self.nnz, self.m = self.sys.nnz, self.sys.m
        # the persistent solver iterates on the FREE unknowns only (Dirichlet rows/columns dropped): its byte model
        rp, ci = self.sys.pattern()
        free = flags == 0
        rows = np.repeat(np.arange(self.m), np.diff(rp))
        self.m_free = int(free.sum())
        self.nnz_free = int(np.count_nonzero(free[rows] & free[ci]))
        self.last = None

    def close(self):
        for dp in self.dports:
            dp.close()
        self.sys.close()
        self.dm.close()

    def step(self):
        np = self.np
        self.sys.assemble_volume(self.omegas, self.mats)
        for a, dp in enumerate(self.dports):
            dp.add_mass(1j * self.betas[:, a])
        for a, dp in enumerate(self.dports):
            dp.rhs_batch(self.idx[a], 2j * self.betas[:, a], use_mass=True)
        res = self.sys.solve(precond=self.cabi.PRECOND_AUX, tol=1e-10, symmetric=True)
        S = np.zeros((self.F, self.P, self.P), dtype=complex)
        for j, dp in enumerate(self.dports):
            v = dp.project_batch(self.all_idx, use_mass=True).reshape(self.F, self.P)
            S[:, j, :] = v
        for a in range(self.P):
            S[:, a, a] -= 1.0
        self.last = (S, res)
        return S, res


def cube_extras(ctx, n: int, solve_freq: float, hbm_peak: float):
    """C5 (BASELINE configs[4]): refined PEC cube cavity (the reference's cavity test object, tests/test_cavity_eigenmodes.cpp,
    Kuhn split n^3 x 6 tets, jittered), single frequency: assembly Mtets/s, SpMV / Krylov-iteration GB/s at k0 h = 2 pi / 10
    (SURVEY 8d), and a CONVERGED solve (COCG + auxiliary-space Jacobi, tol 1e-10, random complex b, seed 1234) at
    `solve_freq`.  The solve frequency is fixed in Hz, not in k0 h: at k0 h = 2 pi / 10 the 20 M-tet lossless cavity is
    15 wavelengths across with ~28 000 resonances below k0 and no Krylov method without a multilevel preconditioner gets
    there (measured on n = 6..16: iterations grow like n^2.6)."""
    import numpy as np
    from edgefem_b200 import cabi, meshgen

    t0 = time.perf_counter()
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
    flags = cabi.pec_flags_from_tris(info["edges"].shape[0], info["tri_edges"], trp, 1)
    sysd = cabi.DeviceSystem.from_mesh(dm, n_matrix=1, n_rhs=1)
    sysd.set_dirichlet(flags)
    setup_s = time.perf_counter() - t0
    n_tet, n_node = int(tets.shape[0]), int(xyz.shape[0])
    del xyz, tets, tris, info
    h = 1.0 / n
    omega = (2 * math.pi / (10 * h)) * C0  # k0 h = 2 pi / 10
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    sysd.assemble_volume([omega], mats)
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(sysd.m) + 1j * rng.standard_normal(sysd.m)
    b[flags == 1] = 0
    sysd.rhs_set(0, b)
    sysd.x_set(0, b)
    m, nnz = sysd.m, sysd.nnz
    ms_asm = sysd.bench_kernel(3, 10)
    ms_spmv = sysd.bench_kernel(0, 20)
    ms_bicg = sysd.bench_kernel(1, 10)
    ms_cocg = sysd.bench_kernel(2, 10)
    ms_fp64 = sysd.bench_kernel(4, 5)
    fp64_tflops = 148 * 8 * 256 * 8 * 4096 * 2.0 / ms_fp64 / 1e9  # probe geometry is fixed in solve.cu (sm_count = 148 on B200)
    b_asm = 45.0 * n_tet + 24.0 * n_node + 16.0 * nnz
    b_spmv = nnz * 20.0 + m * 36.0
    b_bicg = 2 * b_spmv + 21 * 16.0 * m
    b_cocg = b_spmv + 10 * 16.0 * m
    out = {
        "mesh": {"n": n, "tets": n_tet, "nodes": n_node, "edges": m, "nnz": nnz, "free_unknowns": int((flags == 0).sum()),
                 "mesh_generation_s": round(t_gen, 2), "device_setup_s": round(setup_s, 2)},
        "fp64_fma_probe_tflops": fp64_tflops,
        "assembly": {"ms": ms_asm, "mtets_per_s": n_tet / ms_asm / 1e3, "algorithmic_gb": b_asm / 1e9, "gbs": b_asm / ms_asm / 1e6,
                     "fp64_gflop_model": 6 * n_tet * 300 / 1e9, "kernel": "k_assemble_volume_s (rank-major schedule, real chunk image)"},
        "spmv": {"ms": ms_spmv, "algorithmic_gb": b_spmv / 1e9, "gbs": b_spmv / ms_spmv / 1e6,
                 "kernel": "k_spmv_tma (CSR-stream, matrix stream by cp.async.bulk + mbarrier, bank-skewed products)"},
        "bicgstab_jacobi_iteration": {"ms": ms_bicg, "algorithmic_gb": b_bicg / 1e9, "gbs": b_bicg / ms_bicg / 1e6},
        "cocg_jacobi_iteration": {"ms": ms_cocg, "algorithmic_gb": b_cocg / 1e9, "gbs": b_cocg / ms_cocg / 1e6},
    }
    for k in ("assembly", "spmv", "bicgstab_jacobi_iteration", "cocg_jacobi_iteration"):
        out[k]["frac_of_hbm_peak"] = out[k]["gbs"] / hbm_peak
    if solve_freq > 0:
        om_s = 2 * math.pi * solve_freq
        sysd.assemble_volume([om_s], mats)
        sysd.rhs_set(0, b)
        ctx.sync()
        ctx.timer_start()
        res = sysd.solve(precond=cabi.PRECOND_AUX, tol=1e-10, max_iterations=40000, symmetric=True)[0]
        ms_solve = ctx.timer_stop()
        x = sysd.x_get(0)
        y = sysd.spmv(0, x)
        b_aux = b_spmv + (10 * 16.0 + 2 * 16.0) * m + 3 * 16.0 * n_node + 2 * m * 8.0  # + w gather/scatter, gradient index lists
        it = max(1, res["iters"])
        out["solve"] = {"frequency_hz": solve_freq, "k0h": om_s / C0 * h, "method": "COCG + auxiliary-space Jacobi (multi-kernel path, device-side scalars)",
                        "tolerance": 1e-10, "iters": res["iters"], "converged": bool(res["converged"]), "residual": res["residual"],
                        "true_residual_host_check": float(np.linalg.norm(b - y) / np.linalg.norm(b)), "seconds": ms_solve / 1e3,
                        "ms_per_iteration": ms_solve / it, "algorithmic_gb_per_iteration": b_aux / 1e9, "gbs": b_aux / (ms_solve / it) / 1e6,
                        "frac_of_hbm_peak": b_aux / (ms_solve / it) / 1e6 / hbm_peak}
    sysd.close()
    dm.close()
    return out


def cube_extras_dist(ctx, dist, rank, world, n: int, solve_freq: float, hbm_peak: float):
    """C5 row-partitioned over the ranks (second bullet): every rank assembles its row block (no communication),
    the SpMV input halo is read from the peers' memory inside the SpMV kernel (CUDA IPC over NVLink), NCCL carries the
    scalar all-reduces; converged COCG + auxiliary-space solve.  All ranks call this; rank 0 returns the record."""
    import numpy as np
    import torch
    from edgefem_b200 import cabi, meshgen, sharding

    uid = [cabi.dist_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.dist_init(rank, world, uid[0])

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    t0 = time.perf_counter()
    xyz, tets, tp, tris, trp = meshgen.cube_cavity(n, jitter=0.1)
    dm, info = cabi.device_mesh_from_conn(ctx, xyz, tets, tp, tris)
    m = int(info["edges"].shape[0])
    flags = cabi.pec_flags_from_tris(m, info["tri_edges"], trp, 1)
    r0, r1 = sharding.row_range(m, rank, world)
    sysd = cabi.DeviceSystem.from_mesh_rows(dm, r0, r1)
    sysd.set_dirichlet(flags)
    setup_s = time.perf_counter() - t0
    n_tet, n_node = int(tets.shape[0]), int(xyz.shape[0])
    del xyz, tets, tris, info
    mats, keep = cabi.make_materials(len(dm.slot_tags))
    om_s = 2 * math.pi * solve_freq
    sysd.assemble_volume([om_s], mats)  # first call builds the assembly schedule of the block
    ctx.timer_start()
    sysd.assemble_volume([om_s], mats)
    ms_asm = maxr(ctx.timer_stop())
    rng = np.random.default_rng(1234)
    b = rng.standard_normal(m) + 1j * rng.standard_normal(m)
    b[flags == 1] = 0
    sysd.rhs_set(0, b[r0:r1])
    nnz = sumr(sysd.nnz)
    b_asm = 45.0 * n_tet + 24.0 * n_node + 16.0 * nnz
    b_spmv = nnz * 20.0 + m * 36.0
    b_aux = b_spmv + (10 * 16.0 + 2 * 16.0) * m + 3 * 16.0 * n_node + 2 * m * 8.0
    ms_spmv = maxr(sysd.dist_bench(0, 20, 0))
    ms_it = maxr(sysd.dist_bench(2, 20, 0))
    dist.barrier()
    ctx.sync()
    t1 = time.perf_counter()
    res = sysd.dist_solve(tol=1e-10, max_iterations=40000, halo_mode=0, precond=cabi.PRECOND_AUX)
    ctx.sync()
    wall = maxr(time.perf_counter() - t1)
    it = max(1, res["iters"])
    out = {"mesh": {"n": n, "tets": n_tet, "nodes": n_node, "edges": m, "nnz": int(nnz), "rows_per_rank": r1 - r0, "setup_s_per_rank": round(setup_s, 2)},
           "partition": "contiguous row blocks of the global edge numbering; every rank holds the whole mesh; halo by peer loads inside the SpMV kernel",
           "assembly": {"ms": ms_asm, "mtets_per_s": n_tet / ms_asm / 1e3, "gbs_total": b_asm / ms_asm / 1e6, "frac_of_hbm_peak_x_gpus": b_asm / ms_asm / 1e6 / (hbm_peak * world)},
           "spmv": {"ms": ms_spmv, "gbs_total": b_spmv / ms_spmv / 1e6, "frac_of_hbm_peak_x_gpus": b_spmv / ms_spmv / 1e6 / (hbm_peak * world)},
           "cocg_aux_iteration": {"ms": ms_it, "gbs_total": b_aux / ms_it / 1e6, "frac_of_hbm_peak_x_gpus": b_aux / ms_it / 1e6 / (hbm_peak * world)},
           "solve": {"frequency_hz": solve_freq, "method": "COCG + auxiliary-space Jacobi, row-partitioned (3 scalar NCCL all-reduces per iteration)", "tolerance": 1e-10,
                     "iters": res["iters"], "converged": bool(res["converged"]), "residual": res["residual"], "seconds": wall, "ms_per_iteration": 1e3 * wall / it}}
    sysd.close()
    dm.close()
    return out


def ncu_traffic(kernel: str, n_matrix: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed `ncu --set full`
    capture of this same command (profiles/ncu_traffic.json); None when no capture matches the workload size."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f)[kernel]
        return float(rec["bytes_per_launch"]) if int(rec["matrices"]) == int(n_matrix) else None
    except Exception:
        return None


def ncu_limiter(kernel: str):
    """What the committed ncu capture of `kernel` names as the saturated unit (a quotation of profiles/, not a measurement of
    this run)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            rec = json.load(f)[kernel]
        return rec["limiter"] + " [" + rec["source"].split(" ")[0] + "]"
    except Exception:
        return None


def wr90_setup(pe):
    import numpy as np

    z = np.load(os.path.join(ROOT, "tests", "golden", "rect_waveguide.npz"))
    hm = pe.mesh_from_arrays(z["xyz"], z["tet_conn"], z["tet_phys"], z["tri_conn"], z["tri_phys"], z["node_ids"].tolist())
    bc = pe.build_edge_pec(hm, 1)
    dims = pe.RectWaveguidePort(WR90_A, WR90_B)
    kc_sq = (math.pi / WR90_A) ** 2
    ports = [pe.build_wave_port_2d(hm, tag, pe.solve_te10_mode(dims, 10e9), set(bc.dirichlet_edges), kc_sq) for tag in (2, 3)]
    return hm, bc, ports


def c1_latency(pe, hm, bc, ports, reps=20):
    """C1 (BASELINE configs[0], tests/benchmark_wr90.cpp:97-180): ONE frequency, 10 GHz, through the reference's own
    per-frequency API calculate_sparams_eigenmode (what waveguide.py:394-433 loops over): latency, iterations, us/iteration."""
    p = pe.MaxwellParams()
    p.omega = 2 * math.pi * 10e9
    for _ in range(3):
        S = pe.calculate_sparams_eigenmode(hm, p, bc, ports)
    t0 = time.perf_counter()
    for _ in range(reps):
        S = pe.calculate_sparams_eigenmode(hm, p, bc, ports)
    dt = (time.perf_counter() - t0) / reps
    S2, st = pe.calculate_sparams_eigenmode_sweep(hm, p, bc, ports, [10e9])
    it = max(1, max(st.iterations))
    return {"api": "pyedgefem.calculate_sparams_eigenmode (host buffers in, S out; device mesh cached after the first call)",
            "frequency_hz": 10e9, "ms_per_point": 1e3 * dt, "points_per_s": 1.0 / dt, "krylov_iterations": [int(v) for v in st.iterations],
            "device_ms": st.device_ms, "us_per_iteration_device": 1e3 * st.device_ms / it,
            "abs_s11": abs(S[0][0]), "abs_s21": abs(S[1][0]), "kernel_launches": int(st.kernel_launches)}


def run_b200(a):
    import numpy as np

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("EDGEFEM_B200_DEVICE", str(local))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from edgefem_b200 import cabi, load_pyedgefem, sharding

    pe = load_pyedgefem()  # fails loudly if the extension was not built
    ctx = cabi.Ctx(local)
    peaks = {}
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fjs:
            peaks = json.load(fjs)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    if a.workload != "wr90":
        return run_workload(a, ctx, pe, hbm_peak, peak_src, rank, world, dist)

    hm, bc, ports = wr90_setup(pe)
    strong = a.scaling == "strong"
    n_total = a.points if strong else a.points * world
    allf = np.linspace(F_LO, F_HI, n_total)
    freqs = list(allf[rank::world])  # == sharding.shard_indices(len(allf), rank, world): round-robin keeps iteration counts balanced
    dev = None
    if dist is not None:
        import torch

        dev = torch.device("cuda", local)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def step_resident():
        S, res = rs.step()
        # the one collective of the path: gather of the P x P S-matrices (16 P^2 bytes per point), inside the timed region
        S_all = sharding.gather_sweep(S, n_total, rank, world, dist=dist, device=dev) if dist is not None else S
        return S, S_all, res

    # ---------------- resident ("value") ----------------
    rs = ResidentSweep(ctx, pe, hm, bc, ports, freqs)
    for _ in range(a.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launch_count()
    # L2 rule: a shard of fewer than 64 points has a working set (matrix values + Krylov vectors) below the 126 MB L2, so
    # the cache is flushed (untimed) before every timed step; larger shards overflow it by themselves
    l2_flush = len(freqs) < 64
    ms_total = 0.0
    for _ in range(a.steps):
        if l2_flush:
            ctx.l2_flush()
        ctx.timer_start()
        S, S_all, res = step_resident()
        ms_total += ctx.timer_stop()
    launches = ctx.launch_count() - launches0
    barrier()
    clocks = sampler.stop()
    if dist is not None:
        import torch

        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        assert S_all.shape[0] == n_total and np.all(np.isfinite(S_all))
    ms_step = ms_total / a.steps
    value = n_total / (ms_step / 1000.0)
    iters = [r["iters"] for r in res]
    assert all(r["converged"] for r in res), "a solve did not converge"
    # roofline of the dominant kernel, timed live with CUDA events on the launching stream (the library's ctx stream).
    # The whole batched solve is ONE launch of a persistent kernel.  Algorithmic bytes per COCG iteration of one matrix with
    # P rhs (B_spmv + 10*16*m per system; the P systems of a matrix share the value/index stream):
    # nnz*20 + 4m + P*(32m + 160m), over the free unknowns the kernels iterate on.
    F, P, nnz, m = rs.F, rs.P, rs.nnz, rs.m
    nnz_f, m_f = rs.nnz_free, rs.m_free
    ms_kernel = rs.sys.last_solve_kernel_ms()
    cl_c, cl_nr, cl_n = rs.sys.last_solve_shape()
    it_per_matrix = [max(iters[f * P:(f + 1) * P]) for f in range(F)]
    if ms_kernel > 0:
        if cl_c:
            kname, kkey = ("k_cocg_cluster<%d> (persistent COCG + aux-space Jacobi, one cluster of %d CTAs per job, matrix slice resident in shared memory, "
                           "DSMEM halo/nodal exchange; %d resident clusters)" % (cl_nr, cl_c, cl_n)), "k_cocg_cluster"
            per_job = nnz_f * 20.0 + 4.0 * m_f + cl_nr * 192.0 * m_f
            jobs_it = float(sum(it_per_matrix)) if cl_nr == P else float(sum(iters))
            solve_bytes = jobs_it * per_job
            note = ("byte model of the same iteration streamed from memory; this kernel keeps matrix, vectors and index lists on chip (shared memory + "
                    "registers), so HBM sees only the per-job matrix load and b/x: it is bound by shared-memory wavefronts and cluster barriers, not HBM")
        else:
            kname, kkey = ("k_cocg_small (persistent COCG + aux-space Jacobi, one CTA per job, SELL-32 SpMV from smem-resident p; %d matrices x %d rhs in one launch; two-rhs jobs, or one-rhs jobs with r, q resident when every system gets its own SM)" % (F, P)), "k_cocg_small"
            solve_bytes = float(sum(it_per_matrix)) * (nnz_f * 20.0 + 4.0 * m_f + P * 192.0 * m_f)
            note = ("byte model over the %d free unknowns / %d free entries the kernel iterates on (Dirichlet rows and columns are dropped); vectors r,q,x stay "
                    "L2-resident per CTA and p lives in shared memory, so part of the algorithmic bytes never reaches HBM (see traffic)" % (m_f, nnz_f))
        roof = {"bound": "hbm", "kernel": kname, "achieved": solve_bytes / ms_kernel / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "frac": solve_bytes / ms_kernel / 1e6 / hbm_peak, "traffic": ncu_traffic(kkey, F), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": solve_bytes, "ms_per_launch": ms_kernel, "launches_per_step": 1, "share_of_step": ms_kernel / ms_step,
                "free_unknowns": m_f, "free_nnz": nnz_f, "note": note, "limiter_per_ncu": ncu_limiter(kkey)}
    else:  # multi-kernel path (EDGEFEM_B200_NO_PERSISTENT=1 and EDGEFEM_B200_CLUSTER=0): batched CSR SpMV dominates
        ms_spmv = rs.sys.bench_kernel(0, 50)
        spmv_bytes = F * nnz * 16.0 + nnz * 4.0 + (m + 1) * 4.0 + F * P * m * 32.0
        roof = {"bound": "hbm", "kernel": "k_spmv<16,2,*> (batched CSR complex128 SpMV, %d matrices x %d rhs)" % (F, P),
                "achieved": spmv_bytes / ms_spmv / 1e6, "peak": hbm_peak, "unit": "GB/s", "frac": spmv_bytes / ms_spmv / 1e6 / hbm_peak,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes, "ms_per_launch": ms_spmv}

    # ---------------- end to end through the public API ----------------
    p = pe.MaxwellParams()

    def step_e2e():
        pe.b200_clear_cache()
        S2, st = pe.calculate_sparams_eigenmode_sweep(hm, p, bc, ports, freqs)
        S2 = np.array(S2)
        S2_all = sharding.gather_sweep(S2, n_total, rank, world, dist=dist, device=dev) if dist is not None else S2
        return S2, S2_all, st

    for _ in range(max(1, min(a.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 3))
    h2d = d2h = 0
    for _ in range(e2e_steps):
        S2, S2_all, st = step_e2e()
        h2d, d2h = st.h2d_bytes, st.d2h_bytes
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        import torch

        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert np.max(np.abs(S2 - S)) < 1e-6, "public API and resident path disagree"

    extras = {}
    if rank == 0:
        try:
            extras["c1_single_frequency"] = c1_latency(pe, hm, bc, ports)
        except Exception as e:
            extras["c1_single_frequency"] = {"error": repr(e)}
    rs.close()
    if a.cube_n > 0 and world > 1:
        try:
            rec = cube_extras_dist(ctx, dist, rank, world, a.cube_n, a.cube_freq, hbm_peak)
            if rank == 0:
                extras["c5_cube_row_partitioned"] = rec
        except Exception as e:
            if rank == 0:
                extras["c5_cube_row_partitioned"] = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if a.cube_n > 0 and world == 1:
        try:
            extras["c5_cube"] = cube_extras(ctx, a.cube_n, a.cube_freq, hbm_peak)
        except Exception as e:  # extras never invalidate the headline number
            extras["c5_cube"] = {"error": repr(e)}
    cores = os.cpu_count() or 1
    cpu1 = cpu_baseline(a.cpu_sample, 1) if a.cpu_sample > 0 else None  # --cpu-sample 0: profiling runs skip the CPU leg
    line = {
        "metric": "wr90_sweep_freq_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "WR-90 eigenmode S-parameter sweep 8-12 GHz, %d points total x 2 ports, %s over %d GPU(s) (rect_waveguide fixture: 4227 tets, "
                               "5745 edges, nnz 85113), tol 1e-10, COCG + auxiliary-space Jacobi; S all-gather inside the timed region" %
                               (n_total, "sharded round-robin" if strong else "one 256-point sub-band per GPU", world),
                   "points_per_gpu": len(freqs), "total_points": n_total, "solves_per_step": n_total * 2,
                   "l2": ("flushed (256 MB written, untimed) before every timed step: the working set of %d systems is below the 126 MB L2" % (len(freqs) * 2))
                         if l2_flush else
                         ("no flush: the per-step working set (matrix values + Krylov vectors of %d systems, %.0f MB per GPU) is larger than the 126 MB L2" %
                          (len(freqs) * 2, (len(freqs) * nnz * 16.0 * 2 + len(freqs) * 2 * m * 16.0 * 9) / 1e6)),
                   "krylov_iterations": {"min": int(min(iters)), "median": int(sorted(iters)[len(iters) // 2]), "max": int(max(iters))},
                   "solver_launch": {"cluster_ctas": cl_c, "rhs_per_job": cl_nr, "resident_clusters": cl_n}},
        "roofline": roof,
        "cpu_baseline": {"value": cpu1, "unit": "points/s", "cores": 1, "kind": "port",
                         "sample": "%d of 256 points, 1 process: CPU oracle (numpy/scipy restatement of the reference path, SuperLU solves; a stand-in, "
                                   "not Eigen -- Eigen is not installable here); host has %d cores" % (a.cpu_sample, cores)},
        "e2e": {"value": n_total / e2e_s, "unit": "points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1000.0 * e2e_s, "api": "pyedgefem.calculate_sparams_eigenmode_sweep (mesh cache cleared every step) + S all-gather"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "extras": extras,
    }
    _LINE.append(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_workload(a, ctx, pe, hbm_peak, peak_src, rank, world, dist):
    """The other named workloads of BASELINE.json, one JSON line each (rank 0; replicas only -- these are single-GPU lines)."""
    import numpy as np
    from edgefem_b200 import meshgen

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    line = {"n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "cpu_baseline": None, "roofline": None}
    if a.workload == "c1":
        hm, bc, ports = wr90_setup(pe)
        r = c1_latency(pe, hm, bc, ports, reps=max(5, a.steps))
        line.update({"metric": "wr90_single_frequency_points_per_s", "value": r["points_per_s"], "unit": "points/s", "ms_per_step": r["ms_per_point"],
                     "config": {"workload": "C1: WR-90 S-parameters at 10 GHz (tests/benchmark_wr90.cpp) through calculate_sparams_eigenmode"},
                     "e2e": {"value": r["points_per_s"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 64}, "extras": r,
                     "gpu_launches": r["kernel_launches"]})
    elif a.workload == "cube":
        ex = cube_extras(ctx, a.cube_n, a.cube_freq, hbm_peak)
        line.update({"metric": "c5_assembly_mtets_per_s", "value": ex["assembly"]["mtets_per_s"], "unit": "Mtets/s", "ms_per_step": ex["assembly"]["ms"],
                     "config": {"workload": "C5: refined PEC cube cavity n=%d, single frequency: assembly + SpMV + converged Krylov solve" % a.cube_n},
                     "roofline": {"bound": "hbm", "kernel": ex["assembly"]["kernel"], "achieved": ex["assembly"]["gbs"], "peak": hbm_peak, "unit": "GB/s",
                                  "frac": ex["assembly"]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src},
                     "e2e": None, "extras": ex, "gpu_launches": 10})
    else:
        launches0 = pe.b200_launch_count()
        if a.workload == "patch":
            xyz, tets, tp, tris, trp, info = meshgen.patch_antenna(hmax_scale=0.7)
            hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
            bc = pe.BC()
            for tag in (1, 2, 3, 4, 10):
                bc.merge(pe.build_edge_pec(hm, tag))
            freqs = list(np.linspace(2.2e9, 2.7e9, 11))
            opts = pe.SolveOptions()
            opts.use_direct = True

            def point(f):
                ph = pe.MaxwellParams()
                ph.omega = 2 * math.pi * f
                ph.use_abc = True
                ph.abc_surface_tags = {50}
                ph.set_eps_r_region(110, complex(4.4, -4.4 * 0.02))
                cfg = pe.LumpedPortConfig()
                cfg.surface_tag, cfg.z0, cfg.e_direction = 5, 50.0, [1.0, 0.0, 0.0]
                ports = pe.normalize_port_weights(hm, ph, bc, [pe.build_lumped_port(hm, cfg)], opts)
                return pe.calculate_sparams(hm, ph, bc, ports, opts)[0][0]

            name = ("C3: probe-fed patch antenna on FR-4 (28.6 x 37.3 mm, h 1.6 mm, probe -9 mm), lumped port + ABC, 11 points 2.2-2.7 GHz, normalize_port_weights "
                    "+ calculate_sparams per point (stacked_patch.py:503-561); structured mesh %d tets, %d edges" % (hm.num_tets(), hm.num_edges()))
        else:
            h_air = C0 / 10e9 / 2
            xyz, tets, tp, tris, trp = meshgen.unit_cell(px=5e-3, py=5e-3, h_sub=0.5e-3, h_air=h_air, nx=8, ny=8, nz_sub=2, nz_air=8, patch=(4e-3, 4e-3))
            hm = pe.mesh_from_arrays(xyz, tets, tp, tris, trp)
            bc = pe.build_edge_pec(hm, 1)
            freqs = list(np.linspace(8e9, 12e9, 5))

            def point(f):
                omega = 2 * math.pi * f
                pbc = pe.build_periodic_pairs(hm, 5, 6, [5e-3, 0.0, 0.0])
                pe.set_floquet_phase(pbc, [0.0, 0.0])
                ph = pe.MaxwellParams()
                ph.omega = omega
                ph.eps_r_regions = {100: complex(3.5, 0.0), 101: complex(1.0, 0.0)}
                ph.use_port_abc = True
                mdl = pe.materials.DrudeLorentzMaterial(3.5, 2 * math.pi * 4e9, 2 * math.pi * 0.5e9)
                mdl.add_lorentz_pole(0.8, 2 * math.pi * 15e9, 2 * math.pi * 1e9)
                ph.set_eps_model(100, mdl)
                hs = pe.extract_surface_mesh(hm, 4)
                mode = pe.solve_port_eigens(hs.mesh, 1, omega, 1.0, 1.0, pe.ModePolarization.TE)[0]
                return pe.calculate_sparams_periodic(hm, ph, bc, pbc, [pe.build_wave_port(hm, hs, mode)])[0][0]

            name = ("C4: periodic unit cell 5 x 5 mm, Bloch phase along x (normal incidence), Drude-Lorentz substrate, modal top port + port ABC, 5 points "
                    "8-12 GHz (unit_cell.py:691-770); %d tets, %d edges" % (hm.num_tets(), hm.num_edges()))
        for _ in range(max(1, min(a.warmup, 2))):
            vals = [point(f) for f in freqs]
        t0 = time.perf_counter()
        for _ in range(a.steps):
            vals = [point(f) for f in freqs]
        dt = (time.perf_counter() - t0) / a.steps
        assert all(np.isfinite(v) for v in vals), "a solve did not converge"
        line.update({"metric": "%s_sweep_freq_points_per_s" % a.workload, "value": len(freqs) / dt, "unit": "points/s", "ms_per_step": 1e3 * dt,
                     "config": {"workload": name},
                     "e2e": {"value": len(freqs) / dt, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16 * len(freqs),
                             "api": "public per-frequency API from host buffers (wall clock)"},
                     "extras": {"abs_s11": [abs(v) for v in vals]}, "gpu_launches": int((pe.b200_launch_count() - launches0) / (a.steps + max(1, min(a.warmup, 2))))})
    line["clocks"] = sampler.stop()
    _LINE.append(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly one JSON line.  Libraries write banners to file descriptor 1 behind Python's back (NCCL prints
    # "NCCL version ..." when the box sets NCCL_DEBUG=VERSION, and reads that variable before Python can change it in every
    # rank), so fd 1 points at stderr while the benchmark runs and is restored for the final print.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        if a.impl == "reference":
            run_reference(a)
        else:
            run_b200(a)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    if _LINE:
        print(_LINE[0], flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- headline benchmark of the PQC hot path on B200.

Workload (BASELINE.json configs[2], the 16-qubit configuration the metric is quoted on):
TFIM Hamiltonian-variational circuit, 16 qubits x 16 layers (P = 32 parameters),
QFIM + effective quantum dimension (cutoff 1e-12) for S parameter sets per GPU.
One *step* = one pass of that path over the S x 32 synthetic angle batch
(np.random.default_rng(1), the stream the reference itself consumes).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --impl reference [...]                       reference algorithm on host cores

Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PQC samples/sec (statevector+measures) at 16q, 1/8 B200; % HBM roofline"
CUTOFF = 1e-12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--circuit", default="TFIM", choices=["TFIM", "XXZ"])
    ap.add_argument("--qubits", type=int, default=16)
    ap.add_argument("--layers", type=int, default=16)
    ap.add_argument("--samples", type=int, default=10000, help="parameter sets per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --samples is the TOTAL, split over the GPUs "
                         "(BASELINE config 3 as worded: 1e4 parameter sets sharded across 1/2/4/8)")
    return ap.parse_args()


def workload_name(a):
    return (f"{a.circuit} {a.qubits}q x {a.layers} layers: QFIM + effective quantum dimension "
            f"(cutoff {CUTOFF:g}) over {a.samples} parameter sets per GPU")


def angles_for(a, n_params, rank):
    """rows [rank*S, (rank+1)*S) of the one global stream default_rng(1).random((G*S, P))."""
    rng = np.random.default_rng(1)
    skip = rank * a.samples * n_params
    if skip:
        rng.random(skip)
    return rng.random((a.samples, n_params)) * 2 * np.pi


# ---------------------------------------------------------------------------------------
# CPU side: the oracle's literal restatement of update_state + get_QFI + EQD
# (circuit.py:127-130,149-192; measure.py:33-87): P full re-simulations per sample.
# ---------------------------------------------------------------------------------------
def _cpu_one_derivative(job):
    from oracle import pqc_oracle as orc
    specs, n, ang, init, idx = job
    if idx < 0:
        return orc.run(specs, n, ang, init)[0]
    return orc.gradients(specs, n, ang, init, only=[idx])[0, 0]


def cpu_qfim_eqd(specs, n, ang_row, init, pool=None):
    from oracle import pqc_oracle as orc
    P = orc.n_params(specs)
    jobs = [(specs, n, ang_row[None, :], init, i) for i in range(-1, P)]
    res = pool.map(_cpu_one_derivative, jobs) if pool is not None else \
        [_cpu_one_derivative(j) for j in jobs]
    F = orc.qfi(res[0], np.stack(res[1:]))
    return orc.eqd(F, CUTOFF), F


def cpu_baseline(a):
    """Single host thread, one parameter set of the very same workload."""
    from oracle import pqc_oracle as orc
    specs, init = orc.generate_circuit(a.circuit, a.qubits, a.layers)
    ang = angles_for(a, orc.n_params(specs), 0)
    t0 = time.perf_counter()
    _, F = cpu_qfim_eqd(specs, a.qubits, ang[0], init)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "_qfim_row0": F, "unit": "samples/s", "cores": 1, "kind": "port",
            "sample": f"1 of the {a.samples} parameter sets (row 0), full {a.qubits}q x "
                      f"{a.layers} layers, literal {orc.n_params(specs)} re-simulations + QFIM "
                      f"+ eigh; {dt:.1f} s of numpy on one core; QuTiP itself is not "
                      f"installable offline so the restated reference is timed"}


def run_reference(a):
    """--impl reference: the reference algorithm (oracle port) on all host cores; one
    step = one parameter set, its P+1 independent simulations spread over a process pool."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import pqc_oracle as orc
    specs, init = orc.generate_circuit(a.circuit, a.qubits, a.layers)
    P = orc.n_params(specs)
    ang = angles_for(a, P, 0)
    cores = max(1, min(os.cpu_count() or 1, P + 1))
    with mp.get_context("fork").Pool(cores) as pool:
        for i in range(a.warmup):
            cpu_qfim_eqd(specs, a.qubits, ang[i % a.samples], init, pool)
        t0 = time.perf_counter()
        for i in range(a.steps):
            cpu_qfim_eqd(specs, a.qubits, ang[(a.warmup + i) % a.samples], init, pool)
        dt = time.perf_counter() - t0
    value = a.steps / dt
    base = {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"each step = 1 parameter set of the workload; its {P}+1 independent "
                      f"simulations run on a {cores}-process pool (numpy oracle; QuTiP not "
                      f"installable offline)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_name(a), "step_unit": "1 parameter set"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons, power = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, algorithmic_bytes_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu
    --set full capture, rescaled to this run's launch size through the measured
    traffic / algorithmic-bytes ratio (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            r = json.load(f)[kernel]["ratio"]
        return float(np.mean(r)) * algorithmic_bytes_per_launch
    except Exception:
        return None


def exchange_leg(a, dev, rank, world):
    """N > 1 only, outside the timed region: the one measure with a data-path exchange
     -- expressibility of a sample set sharded over the ranks at BASELINE config 2's
    shape (hardware-efficient ansatz, 10 qubits x 10 layers): NCCL all-gather of the state shards
    overlapped with each rank's diagonal block, cross blocks, ONE int64 all-reduce of the
    histogram, KL on every rank.  Rank 0 also computes the whole histogram alone and the two
    must agree bin for bin."""
    import torch
    import torch.distributed as dist
    import pyramaterised_b200 as pyqc
    from pyramaterised_b200 import dist as pdist, engine

    S = 16384 * world
    qc = pyqc.templates.generate_circuit("generic_HE", 10, 10)
    ang = np.random.default_rng(2).random((S, qc.n_true_params)) * 2 * np.pi
    lo, hi = pdist.shard_bounds(S, rank, world)
    local = qc.program.run(torch.from_numpy(ang[lo:hi]).to(dev), init=qc.initial_state.tensor)
    tm = {}
    pdist.sharded_expressibility(local, S, 2.0 ** 10, timings={})          # warm-up (NCCL setup)
    kl = pdist.sharded_expressibility(local, S, 2.0 ** 10, timings=tm)
    pairs = S * (S - 1) // 2
    bins = engine.n_bins(pairs)
    ok = None
    if rank == 0:
        allst = qc.program.run(torch.from_numpy(ang).to(dev), init=qc.initial_state.tensor)
        h1 = engine.fidelity_hist(allst, bins=bins)[0]
        kl1 = float(engine.kl_haar(h1, 2.0 ** 10).item())
        ok = bool(int(h1.sum().item()) == pairs and kl1 == kl)
    t = torch.tensor([tm[k] for k in ("diag_block_s", "gather_wait_s", "cross_blocks_s",
                                      "all_reduce_s")], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    total = float(t.sum())
    return {"what": f"expressibility of {S} generic_HE 10q x 10 states sharded over {world} ranks: "
                    "all_gather_into_tensor (NCCL) overlapped with the diagonal block, cross "
                    "blocks, one int64 all-reduce",
            "pairs": pairs, "bins": bins, "kl": kl, "equals_single_gpu": ok,
            "seconds_max_over_ranks": {"diag_block_with_gather_in_flight": float(t[0]),
                                       "gather_wait": float(t[1]), "cross_blocks": float(t[2]),
                                       "all_reduce": float(t[3]), "total": total},
            "gather_bytes_per_rank": tm["gather_bytes_per_rank"], "all_reduce_bytes": tm["hist_bytes"],
            "pairs_per_s": pairs / total if total > 0 else None}


def run_b200(a):
    import torch
    import torch.distributed as dist
    import pyramaterised_b200 as pyqc
    from pyramaterised_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.strong:
        a.samples = (a.samples + world - 1) // world
    out_fd = 1
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: whatever the libraries print to fd 1 (NCCL's
        # version banner, INFO lines) is routed to stderr; the line itself goes to the saved fd
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_baseline(a)          # before the GPU run, on this box's host cores

    qc = pyqc.templates.generate_circuit(a.circuit, a.qubits, a.layers, shuffle=False)
    P = qc.n_true_params
    m = pyqc.measure.Measurements(qc)
    ang_host = torch.from_numpy(angles_for(a, P, rank)).pin_memory()
    ang_dev = ang_host.to(dev)
    eq_host = torch.empty((a.samples,), dtype=torch.int32).pin_memory()
    ev_host = torch.empty((a.samples, P), dtype=torch.float64).pin_memory()

    def step_resident():
        F = qc.qfim_batch(ang_dev)
        w = engine.eigvalsh(F)
        return engine.count_greater(w, CUTOFF), w

    def step_e2e():
        """Public API call with HOST buffers: pinned angles in, EQD + spectrum out."""
        F, eq, w = m.qfim_batch(ang_host, cutoff_eigvals=CUTOFF, want_eigvals=True)   # H2D inside
        eq_host.copy_(eq, non_blocking=True)
        ev_host.copy_(w, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return eq_host

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(max(3, a.warmup)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = engine.launch_count()
    engine.profile_begin()
    ms, (eq, w) = timed(step_resident, a.steps)
    prof = engine.profile_end()
    launches = engine.launch_count() - l0
    clocks = sampler.stop() if sampler else None

    # secondary: pure state generation (PQC.run over a batch) of the same circuit -- the
    # "gate apply" figure of SURVEY 8d: L x 2 x 16 x 2^n bytes per state
    S_apply = min(a.samples, 4096)
    st_buf = torch.empty((S_apply, 1 << a.qubits), dtype=torch.complex128, device=dev)
    init_t = qc.initial_state.tensor

    def step_apply():
        return qc.program.run(ang_dev[:S_apply], init=init_t, out=st_buf)
    for _ in range(3):
        step_apply()
    engine.profile_begin()
    ms_apply, _ = timed(step_apply, 3)
    prof_apply = engine.profile_end()
    del st_buf

    eq_spectrum0 = w[0].clone()
    step_e2e()
    ms_e2e, eq_h = timed(step_e2e, a.steps)
    assert np.array_equal(eq_h.numpy(), eq.cpu().numpy())

    lt = torch.tensor([launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
    total = a.samples * world
    value = total * a.steps / (ms / 1e3)
    e2e = total * a.steps / (ms_e2e / 1e3)
    peak, peak_src = measured_peak_gbs()
    # the dominant pass kernel of the timed region (the library times every pass-kernel launch
    # with an event pair on its stream and reports the kernels separately)
    by = prof["by_kernel"]
    kern = max(by, key=lambda k: by[k]["ms"]) if by else "none"
    kp = by.get(kern, {"ms": 0.0, "launches": 0, "bytes": 0.0})
    achieved = kp["bytes"] / (kp["ms"] / 1e3) / 1e9 if kp["ms"] > 0 else 0.0
    alg_per_launch = kp["bytes"] / max(1, kp["launches"])
    exchange = None
    strong = None
    if world > 1:
        try:
            exchange = exchange_leg(a, dev, rank, world)
        except Exception as e:            # the headline line must still be printed
            exchange = {"error": f"{type(e).__name__}: {e}"[:300]}
        if not a.strong:
            # BASELINE config 3 as worded: the SAME --samples parameter sets in total, split over
            # the ranks (strong scaling), outside the headline timed region
            try:
                per = (a.samples + world - 1) // world
                sl = ang_dev[:per]

                def step_strong():
                    F = qc.qfim_batch(sl)
                    return engine.count_greater(engine.eigvalsh(F), CUTOFF)
                for _ in range(3):
                    step_strong()
                ms_s, _ = timed(step_strong, a.steps)
                strong = {"what": f"the same workload with {per * world} parameter sets in total "
                                  f"({per} per GPU)", "samples_total": per * world,
                          "value": per * world * a.steps / (ms_s / 1e3), "unit": "samples/s",
                          "ms_per_step": ms_s / a.steps}
            except Exception as e:        # the headline line must still be printed
                strong = {"error": f"{type(e).__name__}: {e}"[:300]}
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "strong" if a.strong else "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": workload_name(a), "circuit": a.circuit, "n_qubits": a.qubits,
                   "layers": a.layers, "n_params": P, "samples_per_gpu": a.samples,
                   "sharding": f"samples x{world}, no data-path collective",
                   "l2": "no flush needed: each chunk's live vectors (33 MB per parameter "
                         "set x thousands of sets) exceed the 126 MB L2 many times over",
                   "eqd_histogram": np.bincount(eq.cpu().numpy()).tolist(),
                   # the 1e-12 cutoff sits inside the rounding noise of the 17th eigenvalue for
                   # ~6 % of the sets (profiles/r2_eqd_noise.json); at 1e-10 every path agrees
                   "eqd_histogram_cutoff_1e-10":
                       np.bincount(engine.count_greater(w, 1e-10).cpu().numpy()).tolist()},
        "e2e": {"value": e2e, "unit": "samples/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": int(ang_host.numel() * 8),
                "d2h_bytes_per_step": int(eq_host.numel() * 4 + ev_host.numel() * 8)},
        "gpu_launches": int(lt.item()),
        "roofline": {"kernel": kern, "bound": "hbm", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "launches": kp["launches"],
                     "avg_launch_ms": kp["ms"] / max(1, kp["launches"]),
                     "kernel_share_of_step": kp["ms"] / ms,
                     "pass_kernels_in_step": {k: {"launches": v["launches"],
                                                  "share_of_step": v["ms"] / ms,
                                                  "GBps": v["bytes"] / (v["ms"] / 1e3) / 1e9}
                                              for k, v in by.items()},
                     "algorithmic_bytes_per_launch": alg_per_launch,
                     "traffic": ncu_traffic(kern, alg_per_launch)},
        "roofline_apply_only": {
            "what": f"PQC.run_batch, {S_apply} states per GPU, no measures",
            "states_per_s": S_apply * world * 3 / (ms_apply / 1e3),
            "by_template_layers_GBps": (a.layers + (1 if a.circuit == "TFIM" else 0)) * 2 * 16 *
            (1 << a.qubits) * S_apply * 3 / (ms_apply / 1e3) / 1e9,
            "pass_kernel_GBps": prof_apply["bytes"] / (prof_apply["ms"] / 1e3) / 1e9
            if prof_apply["ms"] > 0 else 0.0,
            "passes": qc.program.n_passes, "peak": peak, "unit": "GB/s"},
        "clocks": clocks,
    }
    if exchange is not None:
        line["exchange"] = exchange
    if strong is not None:
        line["strong_scaling"] = strong
    ra = line["roofline_apply_only"]
    ra["frac_by_layers"] = ra["by_template_layers_GBps"] / peak
    ra["frac_pass_kernel"] = ra["pass_kernel_GBps"] / peak
    if cpu is not None:
        # the oracle's QFIM of row 0 (computed for the CPU baseline anyway) checks the GPU path
        F0 = qc.qfim_batch(ang_dev[:1])[0].cpu().numpy()
        R0 = cpu.pop("_qfim_row0")
        err = float(np.abs(F0 - R0).max() / np.abs(R0).max())
        line["parity_check"] = {"what": "QFIM of parameter set 0 vs the numpy oracle, max abs "
                                        "error / max |F|", "value": err, "tolerance": 1e-8}
        if not err < 1e-8:
            raise SystemExit(f"bench: GPU QFIM differs from the oracle (rel {err:.3e})")
        # EQD of the same set: GPU (Jacobi kernel on the GPU QFIM) vs oracle (scipy-style eigh of the
        # oracle QFIM, measure.py:77-87) at the reference's cutoff and at a noise-free one
        w0 = np.sort(eq_spectrum0.cpu().numpy())
        wr = np.linalg.eigvalsh(R0)
        line["parity_check"]["eqd_set0"] = {
            "gpu_1e-12": int((w0 > CUTOFF).sum()), "oracle_1e-12": int((wr > CUTOFF).sum()),
            "gpu_1e-10": int((w0 > 1e-10).sum()), "oracle_1e-10": int((wr > 1e-10).sum())}
        if (w0 > 1e-10).sum() != (wr > 1e-10).sum():
            raise SystemExit("bench: GPU EQD (cutoff 1e-10) differs from the oracle")
        line["cpu_baseline"] = cpu
    if rank == 0:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)

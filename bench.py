#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the element-loop hot path.

A "step" is one pass of the hot path over one batch of synthetic input:
    [Taylor-Hood only: LocalView::bind/index table of every element]
    interpolate(basis, coeff, f)          f evaluated at EVERY Lagrange node (interpolate.hh:151-173)
    [N > 1: one-plane exchange of interface DOFs with the z-neighbours]
    DiscreteGlobalBasisFunction evaluation at the tensor Gauss points of every element
Default workload "c5w" = the configuration BASELINE.json's target is quoted on: LagrangeBasis Q2 on a
3-d YaspGrid 512^3 PER GPU (z-slabs, weak scaling; the N = 1 case is configs[4] on one GPU), 3^3 Gauss
points.  At N = 1 the line also carries a `configs` array with the other BASELINE configs (c2, c3, c4),
each with its own roofline fractions and an in-line parity check against the CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c5w]

Timing: CUDA events on the launching stream around exactly K steps after W >= 3 warm-up steps, barrier +
synchronize on both sides, max over ranks.  Every workload but c1 moves far more than the 126 MB L2 per
step, so no flush is needed between iterations.  After the K steps the same step keeps running for about a
second ("sustained") so that the clock samples cover a long region.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: per-rank (weak) or global (strong) n, tree, gauss points per direction, jacobians, scaling
    "c1": dict(n=[64, 64], tree="lagrange1", m=2, jac=False, scaling="weak",
               desc="LagrangeBasis Q1, 2-d YaspGrid 64x64 per GPU, interpolate exp(-|x|^2) + eval at 2^2 Gauss points"),
    "c2": dict(n=[256, 256, 256], tree="lagrange2", m=3, jac=False, scaling="weak",
               desc="LagrangeBasis Q2, 3-d YaspGrid 256^3 per GPU (z-slabs), interpolate exp(-|x|^2) + "
                    "DiscreteGlobalBasisFunction eval at 3^3 Gauss points"),
    "c3": dict(n=[128, 128, 128], tree="taylorhood", m=3, jac=False, scaling="weak",
               desc="Taylor-Hood composite(power<3>(Lagrange<2>), Lagrange<1>), 3-d YaspGrid 128^3 per GPU: "
                    "global multi-index map + interpolation (velocity x, pressure exp(-|x|^2)) + eval"),
    "c4": dict(n=[192, 192, 192], tree="lagrangedg4", m=5, jac=True, scaling="weak",
               desc="LagrangeDGBasis order 4, 3-d YaspGrid 192^3 per GPU: values and gradients at 5^3 Gauss points of "
                    "(interpolant of a quadratic + seeded random vector); the step also interpolates the quadratic"),
    "c5": dict(n=[512, 512, 512], tree="lagrange2", m=3, jac=False, scaling="strong",
               desc="LagrangeBasis Q2, 3-d YaspGrid 512^3 slab-partitioned over the GPUs, interpolate + eval at 3^3"),
    "c5w": dict(n=[512, 512, 512], tree="lagrange2", m=3, jac=False, scaling="weak",
                desc="LagrangeBasis Q2, 3-d YaspGrid 512^3 per GPU (z-slabs; ~1.08e9 DOFs per GPU), interpolate exp(-|x|^2) "
                     "at every node + DiscreteGlobalBasisFunction eval at 3^3 Gauss points"),
}
WORKLOADS["q3"] = dict(n=[128, 128, 128], tree="lagrange3", m=4, jac=False, scaling="weak",
                       desc="(not a BASELINE config; SURVEY.md section 8(f) row 3's base case) LagrangeBasis Q3, 3-d YaspGrid 128^3 per GPU, "
                            "interpolate exp(-|x|^2) at every node + eval at 4^3 Gauss points")
HEADLINE = "c5w"
SUB_CONFIGS = ("c2", "c3", "c4", "q3")
METRIC = "DOFs/s through interpolate + DiscreteGlobalBasisFunction evaluation (FP64)"
TOL = 1e-12


def make_tree(name):
    import dune_functions_b200 as dfx
    return {"lagrange1": lambda: dfx.lagrange(1), "lagrange2": lambda: dfx.lagrange(2), "lagrange3": lambda: dfx.lagrange(3),
            "taylorhood": lambda: dfx.taylorHood(), "lagrangedg4": lambda: dfx.lagrangeDG(4)}[name]()


def global_grid(wl, world):
    n = list(wl["n"])
    if wl["scaling"] == "weak":
        n[-1] *= world
    return n


def workload_function(wl):
    """the analytic f of the workload: exp(-|x|^2); DG4: the quadratic of
    discreteglobalbasisfunctionderivativetest.cc:89-91"""
    import dune_functions_b200 as dfx
    if wl["tree"] == "lagrangedg4":
        return dfx.quadratic(0.0, [0, 0, 0], [42, 7, 0, 13, 0, 5]), "42 x0^2 + 13 x1^2 + 7 x0 x1 + 5 x2^2"
    return dfx.gaussian(1.0), "exp(-|x|^2)"


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.stamps, self.reasons, self.max_mhz = [], [], set(), None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.stamps.append(time.perf_counter())
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def report(self, t_timed_end):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        timed = [s for s, t in zip(self.samples, self.stamps) if t <= t_timed_end]
        r = {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
             "reasons": sorted(self.reasons), "samples": len(self.samples),
             "samples_in_timed_region": len(timed),
             "note": "sm_mhz = median over the K timed steps plus the sustained continuation of the same step"}
        if len(timed) >= 3:
            r["sm_mhz_timed_steps_only"] = float(np.median(timed))
        return r


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def bind_to_gpu_numa_node(phys_index):
    """run (and first-touch pinned host buffers) on the CPUs NVML reports as local to the GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(phys_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profiles(kernel, workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path)).get(workload, {}).get(kernel)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference loops) on a bounded box of the workload's grid
def th_node_fns():
    import dune_functions_b200 as dfx
    # expanded Taylor-Hood tree: 0 composite, 1 power, 2..4 Q2 leaves, 5 Q1 leaf
    return [(1, dfx.identity(3)), (5, dfx.gaussian(1.0))]


def cpu_step(ob, f, pts, jac, nthreads, node_fns=None):
    """one oracle step (interpolate + evaluate) on the sample basis; returns seconds (total, interpolate, evaluate)"""
    t0 = time.perf_counter()
    c = np.zeros(ob.dimension())
    if node_fns is None:
        ob.interpolate(f, coeff=c, nthreads=nthreads)
    else:
        for node, fn in node_fns:
            ob.interpolate(fn, coeff=c, node=node, nthreads=nthreads)
    t1 = time.perf_counter()
    ob.evaluate(c, pts, values=True, jacobians=jac, nthreads=nthreads)
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, t2 - t1


def cpu_sample(wl, budget_s, nthreads):
    """largest cube box of the workload's grid (same mesh width, same functions) whose oracle step fits the budget"""
    import orc
    import dune_functions_b200 as dfx
    d = len(wl["n"])
    tree = make_tree(wl["tree"])
    pts, _ = dfx.tensor_gauss(d, wl["m"])
    f, _ = workload_function(wl)
    node_fns = th_node_fns() if wl["tree"] == "taylorhood" else None

    def box(edge):
        n = [min(edge, x) for x in wl["n"]]
        return orc.OracleBasis(dfx.YaspGrid(wl["n"], local_begin=[0] * d, local_n=n), tree), n

    ob, n = box(12)
    t, _, _ = cpu_step(ob, f, pts, wl["jac"], nthreads, node_fns)
    per_elem = t / ob.numElements()
    edge = 12
    for cand in (16, 24, 32, 48, 64, 96, 128, 192, 256):
        if float(np.prod([min(cand, x) for x in wl["n"]])) * per_elem <= budget_s:
            edge = cand
    ob, n = box(edge)
    return ob, f, pts, node_fns, n


def run_reference_arm(args, wl, config, warmup):
    """the reference's own CPU path (restatement oracle: the reference cannot be built here, DESIGN.md
    section 5) on all host threads; each step a bounded box of the workload's grid"""
    import orc
    nthreads = orc.lib().orc_max_threads()
    ob, f, pts, node_fns, n = cpu_sample(wl, args.cpu_budget, nthreads)
    for _ in range(warmup):
        cpu_step(ob, f, pts, wl["jac"], nthreads, node_fns)
    t0 = time.perf_counter()
    ti = te = 0.0
    for _ in range(args.steps):
        _, a, b = cpu_step(ob, f, pts, wl["jac"], nthreads, node_fns)
        ti += a
        te += b
    dt = (time.perf_counter() - t0) / args.steps
    val = ob.dimension() / dt
    sample = f"{'x'.join(map(str, n))} element box of the workload grid per step ({ob.dimension()} DOFs)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "DOFs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "DOFs/s", "cores": nthreads, "kind": "port", "sample": sample,
                             "interpolate_dofs_per_s": ob.dimension() * args.steps / ti,
                             "basis_fn_evals_per_s": ob.numElements() * len(pts) * ob.maxNodeSize() * args.steps / te},
            "e2e": {"value": val, "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
class Runner:
    """one workload on this rank's slab: handles, buffers and the step"""

    def __init__(self, name, rank, world, local_rank, halo_mode="peer"):
        import torch
        import torch.distributed as dist
        import dune_functions_b200 as dfx
        from dune_functions_b200 import _capi as capi
        from dune_functions_b200.slab import HaloExchange, PeerHaloExchange
        self.torch, self.dist, self.dfx, self.capi = torch, dist, dfx, capi
        self.L = capi.lib()
        self.name, self.wl = name, WORKLOADS[name]
        wl = self.wl
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", local_rank)
        self.ng = global_grid(wl, world)
        self.grid = dfx.YaspGrid(self.ng).slab(rank, world)
        self.tree = make_tree(wl["tree"])
        self.basis = dfx.makeBasis(self.grid, self.tree, device=local_rank)
        basis = self.basis
        self.dim = len(self.ng)
        self.pts, _ = dfx.tensor_gauss(self.dim, wl["m"])
        self.nq = len(self.pts)
        self.f, self.f_name = workload_function(wl)
        self.node_fns = None
        if wl["tree"] == "taylorhood":
            self.node_fns = [(basis.child(0, 0), dfx.identity(3)), (basis.child(0, 1), dfx.gaussian(1.0))]
        self.ne = basis.numElements()
        self.ncomp = basis.numLeaves()
        self.ndof_local = basis.dimension()
        # whole-job DOF count = dimension of the same basis on the global grid (host bookkeeping)
        self.ndof_global = dfx.makeBasis(dfx.YaspGrid(self.ng), self.tree, device=local_rank).dimension() if world > 1 \
            else self.ndof_local
        self.coeff = basis.newVector(0.0)
        self.values = torch.empty((self.ne, self.nq, self.ncomp), dtype=torch.float64, device=self.dev)
        self.jac = torch.empty((self.ne, self.nq, self.ncomp, self.dim), dtype=torch.float64, device=self.dev) if wl["jac"] else None
        # configs[3] (DG4) is an EVALUATION config: its coefficients are the interpolant of the quadratic plus a
        # seeded random vector, prepared once; the step still times the interpolation of
        # the quadratic into `coeff`, and evaluates `coeff_eval`
        self.coeff_eval = self.coeff
        if wl["tree"] == "lagrangedg4":
            g = torch.Generator(device=self.dev)
            g.manual_seed(12345)
            self.coeff_eval = basis.newVector(0.0)
            dfx.interpolate(basis, self.coeff_eval, self.f)
            self.coeff_eval += torch.empty_like(self.coeff_eval).uniform_(-1.0, 1.0, generator=g)
        self.dgbf = dfx.makeDiscreteGlobalBasisFunction(basis, self.coeff_eval)
        self.index_table = None
        if wl["tree"] == "taylorhood":
            self.index_table = torch.empty((self.ne, basis.maxNodeSize()), dtype=torch.int32, device=self.dev)
        self.halo = HaloExchange(basis, rank, world, device=self.dev)
        self.peer, self.halo_note = None, halo_mode
        if world > 1 and halo_mode == "peer":
            # every rank must get its link (CUDA IPC between the processes of this node); otherwise all
            # ranks take the NCCL transport together and the JSON line says so
            ok = 1
            try:
                self.peer = PeerHaloExchange(basis, rank, world)
            except Exception as exc:          # noqa: BLE001
                ok, self.peer = 0, None
                print(f"rank {rank}: peer-memory halo link unavailable ({exc}); using NCCL", file=sys.stderr)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if self.peer is not None:
                    self.peer.close()
                self.peer, self.halo_note = None, "nccl (peer link unavailable)"
        # N > 1: the exchange overlaps the evaluation of the elements that do not touch the interface
        # plane (all z-layers but the top one); the top layer follows once the plane has arrived
        layer = self.ne // self.grid.local_n[-1]
        self.split = self.ne - layer if (world > 1 and rank < world - 1 and self.halo.n > 0 and self.grid.local_n[-1] > 1) else self.ne
        self.comm_stream = torch.cuda.Stream(device=self.dev) if world > 1 else None
        self.v_flat = self.values.view(self.ne, -1)
        self.j_flat = self.jac.view(self.ne, -1) if self.jac is not None else None
        self.ev = {k: [] for k in ("index", "interp", "halo", "eval")}

    def stream_ptr(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def evaluate_range(self, a, b, jac=None):
        jac = self.wl["jac"] if jac is None else jac
        if b > a:
            self.dgbf.evaluate(self.pts, values=True, jacobians=jac, elem_begin=a, elem_end=b, out_values=self.v_flat[a:b],
                               out_jacobians=None if (self.j_flat is None or not jac) else self.j_flat[a:b])

    def indices(self):
        if self.index_table is not None:
            self.capi.check(self.L.dfx_indices_flat(self.basis._h, 0, self.ne, C.c_void_p(self.index_table.data_ptr()), self.stream_ptr()))

    def interpolate(self, separable=False):
        if self.node_fns is None:
            self.dfx.interpolate(self.basis, self.coeff, self.f, separable=separable)
        elif separable:
            for node, fn in self.node_fns:
                self.capi.check(self.L.dfx_interpolate_separable(self.basis._h, node, C.byref(fn), C.c_void_p(self.coeff.data_ptr()),
                                                                 None, self.stream_ptr()))
        else:
            # velocity and pressure sub-spaces as one batched call (one launch for the two lattices)
            self.dfx.interpolateBatch(self.basis, self.coeff, self.node_fns)

    def step(self, timed):
        torch, capi, L = self.torch, self.capi, self.L
        marks = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        self.indices()
        mark()
        self.interpolate()
        mark()
        world, rank = self.world, self.rank
        if world > 1 and self.peer is not None:
            # push my bottom plane into the lower neighbour's buffer right away, evaluate everything that
            # does not need the interface plane, then take the upper neighbour's plane (by now it has
            # usually arrived: the wait is a counter check inside the unpack kernel) and finish with the top layer
            st = self.stream_ptr()
            self.peer.step += 1
            if rank > 0:
                capi.check(L.dfx_halo_push(self.basis._h, self.peer._link, C.c_void_p(self.coeff.data_ptr()), self.peer.step, st))
            self.evaluate_range(0, self.split)
            if rank < world - 1:
                capi.check(L.dfx_halo_pull(self.basis._h, self.peer._link, C.c_void_p(self.coeff.data_ptr()), self.peer.step, st))
            mark()
            self.evaluate_range(self.split, self.ne)
        elif world > 1:
            main = torch.cuda.current_stream()
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                self.halo(self.coeff)      # my top plane <- bottom plane of the slab above (NCCL send/recv)
            self.evaluate_range(0, self.split)
            main.wait_stream(self.comm_stream)
            mark()
            self.evaluate_range(self.split, self.ne)
        else:
            mark()
            self.evaluate_range(0, self.ne)
        mark()
        if timed:
            self.ev["index"].append((marks[0], marks[1]))
            self.ev["interp"].append((marks[1], marks[2]))
            self.ev["halo"].append((marks[2], marks[3]))      # N > 1: interior evaluation overlapped with the exchange
            self.ev["eval"].append((marks[2], marks[4]))      # the whole evaluation phase (N > 1: includes the exchange)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def phase_ms(self):
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) if v else 0.0 for k, v in self.ev.items()}

    def time_call(self, fn, reps):
        """median CUDA-event time of fn() in ms (after 2 warm-up calls)"""
        torch = self.torch
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    # ---- algorithmic bytes: each datum once
    def eval_bytes(self, jac=None):
        jac = self.wl["jac"] if jac is None else jac
        return self.ndof_local * 8 + self.ne * self.nq * self.ncomp * 8 * (1 + (self.dim if jac else 0))

    def interp_bytes(self):
        return self.ndof_local * 8

    def index_bytes(self):
        return 0 if self.index_table is None else self.index_table.numel() * 4

    # ---- parity against the CPU oracle on element ranges of THIS grid, using the buffers the last step wrote
    # (the caller runs one untimed step on every rank first, so coeff / values / jac are what a step leaves)
    def parity(self, n_ranges=3, count=None):
        import orc
        torch = self.torch
        basis, ne = self.basis, self.ne
        if count is None:
            count = {"lagrangedg4": 4096, "taylorhood": 8192, "lagrange3": 8192}.get(self.wl["tree"], 32768)
        count = min(count, ne)
        count -= count % 128 if count >= 128 else 0
        starts = sorted({0, ((ne - count) // 2) // 128 * 128, ne - count})[:n_ranges]
        ob = orc.OracleBasis(self.grid, self.tree)
        nthreads = max(1, min(16, orc.lib().orc_max_threads()))
        jac = self.wl["jac"]
        res = {"indices_equal": True, "interpolate_max_rel": 0.0, "evaluate_max_rel": 0.0,
               "jacobian_max_rel": 0.0 if jac else None, "evaluate_random_coefficients_max_rel": 0.0,
               "elements_checked": 0, "tolerance": TOL,
               "against": "CPU oracle (oracle/dfx_oracle.cpp) on element ranges of the same grid: index table, interpolated "
                          "coefficients, the point values the last step left in the output buffers, and a second evaluation "
                          "of seeded random coefficients"}
        torch.cuda.synchronize()
        multi = basis.maxMultiIndexSize() > 1
        g = torch.Generator(device=self.dev)
        g.manual_seed(12345)
        x_rand = torch.empty_like(self.coeff).uniform_(-1.0, 1.0, generator=g)
        f_rand = self.dfx.makeDiscreteGlobalBasisFunction(basis, x_rand)

        def rel(a, b):
            return float(np.abs(a - b).max() / max(1.0, np.abs(b).max())) if a.size else 0.0
        for e0 in starts:
            e1 = e0 + count
            idx = basis.flatIndices(e0, e1, dtype=torch.int64)
            idx_g = idx.cpu().numpy()
            idx_o = ob.flatIndices(e0, e1).astype(np.int64)
            res["indices_equal"] = res["indices_equal"] and bool(np.array_equal(idx_g, idx_o))
            if self.index_table is not None:      # the table the step itself wrote (u32)
                tab = self.index_table[e0:e1].cpu().numpy().view(np.uint32).astype(np.int64)
                res["indices_equal"] = res["indices_equal"] and bool(np.array_equal(tab, idx_o))
            if multi:
                dig = basis.indices(e0, e1).cpu().numpy().astype(np.uint64)
                res["indices_equal"] = res["indices_equal"] and bool(np.array_equal(dig, ob.indices(e0, e1)[0]))
            c_gpu = self.coeff[idx].cpu().numpy()
            c_orc = np.zeros(ob.dimension())          # calloc: only the touched pages exist
            if self.node_fns is None:
                ob.interpolate(self.f, coeff=c_orc, e0=e0, e1=e1, nthreads=nthreads)
            else:
                for (node, fn) in th_node_fns():
                    ob.interpolate(fn, coeff=c_orc, node=node, e0=e0, e1=e1, nthreads=nthreads)
            res["interpolate_max_rel"] = max(res["interpolate_max_rel"], rel(c_gpu, c_orc[idx_o]))
            # evaluation: the oracle contracts the GPU's own coefficients, so the comparison isolates the kernel
            h = np.zeros(ob.dimension())
            h[idx_o] = self.coeff_eval[idx].cpu().numpy()
            out = ob.evaluate(h, self.pts, values=True, jacobians=jac, e0=e0, e1=e1, nthreads=nthreads)
            v_o, j_o = out if jac else (out, None)
            res["evaluate_max_rel"] = max(res["evaluate_max_rel"], rel(self.values[e0:e1].cpu().numpy(), v_o))
            if j_o is not None:
                res["jacobian_max_rel"] = max(res["jacobian_max_rel"], rel(self.jac[e0:e1].cpu().numpy(), j_o))
            # the same kernels on non-smooth data: seeded random coefficients
            h[idx_o] = x_rand[idx].cpu().numpy()
            got = f_rand.evaluate(self.pts, values=True, jacobians=jac, elem_begin=e0, elem_end=e1)
            out = ob.evaluate(h, self.pts, values=True, jacobians=jac, e0=e0, e1=e1, nthreads=nthreads)
            if jac:
                r = max(rel(got[0].cpu().numpy(), out[0]), rel(got[1].cpu().numpy(), out[1]))
            else:
                r = rel(got.cpu().numpy(), out)
            res["evaluate_random_coefficients_max_rel"] = max(res["evaluate_random_coefficients_max_rel"], r)
            if self.wl["tree"] == "lagrangedg4":
                # ill-conditioned on purpose: the gradient of the SMOOTH interpolant alone (values up to 67 on
                # elements of width 1/192: the terms of the sum are 5700 times the result).  The analytic gradient is
                # known, so the yardstick is how far the reference algorithm itself lands from it.
                f_s = self.dfx.makeDiscreteGlobalBasisFunction(basis, self.coeff)
                _, j_g = f_s.evaluate(self.pts, values=True, jacobians=True, elem_begin=e0, elem_end=e1)
                h[idx_o] = self.coeff[idx].cpu().numpy()
                _, j_s = ob.evaluate(h, self.pts, values=True, jacobians=True, e0=e0, e1=e1, nthreads=nthreads)
                e = np.arange(e0, e1)
                N = self.grid.local_n
                ec = np.stack([e % N[0], (e // N[0]) % N[1], e // (N[0] * N[1])], 1).astype(np.float64)
                P = (ec[:, None, :] + self.pts[None]) / np.array(self.ng, dtype=np.float64)
                exact = np.stack([84 * P[..., 0] + 7 * P[..., 1], 26 * P[..., 1] + 7 * P[..., 0], 10 * P[..., 2]], -1)
                sc = max(1.0, np.abs(exact).max())
                sm = res.setdefault("jacobian_of_smooth_interpolant", {"gpu_vs_oracle": 0.0, "oracle_vs_analytic": 0.0, "gpu_vs_analytic": 0.0})
                j_g = j_g.cpu().numpy()[:, :, 0, :]
                sm["gpu_vs_oracle"] = max(sm["gpu_vs_oracle"], float(np.abs(j_g - j_s[:, :, 0, :]).max() / sc))
                sm["oracle_vs_analytic"] = max(sm["oracle_vs_analytic"], float(np.abs(j_s[:, :, 0, :] - exact).max() / sc))
                sm["gpu_vs_analytic"] = max(sm["gpu_vs_analytic"], float(np.abs(j_g - exact).max() / sc))
            res["elements_checked"] += count
        del x_rand, f_rand
        res["ranges"] = [[int(s), int(s + count)] for s in starts]
        res["ok"] = bool(res["indices_equal"] and res["interpolate_max_rel"] <= TOL and res["evaluate_max_rel"] <= TOL and
                         res["evaluate_random_coefficients_max_rel"] <= TOL and
                         (res["jacobian_max_rel"] is None or res["jacobian_max_rel"] <= TOL))
        sm = res.get("jacobian_of_smooth_interpolant")
        if sm is not None:
            sm["criterion"] = "gpu_vs_analytic <= 3 x max(oracle_vs_analytic, 1e-12): as close to the true gradient as the reference algorithm"
            sm["ok"] = bool(sm["gpu_vs_analytic"] <= 3 * max(sm["oracle_vs_analytic"], TOL))
            res["ok"] = res["ok"] and sm["ok"]
        return res

    # ---- N > 1: is the interface plane really the neighbour's?  (interpolate writes the same values on both
    # sides, so the timed steps cannot tell): rank-dependent random coefficients, top plane poisoned with NaN,
    # one exchange over the transport under test, then bit-for-bit comparison with the upper neighbour's
    # bottom plane sent over NCCL
    def halo_check(self):
        torch, dist, capi, L = self.torch, self.dist, self.capi, self.L
        if self.world == 1 or self.halo.n == 0:
            return None
        n, rank, world = self.halo.n, self.rank, self.world
        g = torch.Generator(device=self.dev)
        g.manual_seed(4242 + rank)
        x = torch.empty_like(self.coeff).uniform_(-1.0, 1.0, generator=g)
        st = self.stream_ptr()
        bottom = torch.empty(n, dtype=torch.float64, device=self.dev)
        expect = torch.empty(n, dtype=torch.float64, device=self.dev)
        top = torch.empty(n, dtype=torch.float64, device=self.dev)
        capi.check(L.dfx_halo_pack(self.basis._h, 0, C.c_void_p(x.data_ptr()), C.c_void_p(bottom.data_ptr()), st))
        if rank < world - 1:
            nan = torch.full((n,), float("nan"), dtype=torch.float64, device=self.dev)
            capi.check(L.dfx_halo_unpack(self.basis._h, 1, C.c_void_p(nan.data_ptr()), C.c_void_p(x.data_ptr()), st))
        ops = []
        if rank > 0:
            ops.append(dist.P2POp(dist.isend, bottom, rank - 1))
        if rank < world - 1:
            ops.append(dist.P2POp(dist.irecv, expect, rank + 1))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        # the transport under test
        if self.peer is not None:
            self.peer(x)
        else:
            self.halo(x)
        ok = 1
        if rank < world - 1:
            capi.check(L.dfx_halo_pack(self.basis._h, 1, C.c_void_p(x.data_ptr()), C.c_void_p(top.data_ptr()), st))
            torch.cuda.synchronize()
            ok = int(torch.equal(top, expect))
        if self.peer is not None and self.peer.error():
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return "ok" if int(flag.item()) == 1 else "fail"

    def free(self):
        torch = self.torch
        if self.peer is not None:
            torch.cuda.synchronize()
            self.dist.barrier()
            self.peer.close()
            self.peer = None
        for k in ("values", "jac", "v_flat", "j_flat", "coeff", "coeff_eval", "dgbf", "index_table", "halo"):
            setattr(self, k, None)
        self.basis = None
        torch.cuda.empty_cache()


def pcie_ceiling(torch, dev, world, dist, nbytes=1 << 30, reps=3):
    """pinned-memory copies, both directions at once, on every rank at the same time: the bound the
    end-to-end number lives under"""
    h_in = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    h_out = torch.empty(nbytes // 8, dtype=torch.float64, pin_memory=True)
    d_in = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    d_out = torch.zeros(nbytes // 8, dtype=torch.float64, device=dev)
    h_in.zero_()
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    res = {}
    for label, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("duplex", True, True)):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[label] = nbytes * reps * (int(do_in) + int(do_out)) / dt / 1e9
    t = torch.tensor([res["h2d"], res["d2h"], res["duplex"]], dtype=torch.float64, device=dev)
    if world > 1:
        tmin = t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {"per_rank_min_gbs": {"h2d": tmin[0].item(), "d2h": tmin[1].item(), "duplex_sum": tmin[2].item()},
                "all_ranks_sum_gbs": {"h2d": t[0].item(), "d2h": t[1].item(), "duplex_sum": t[2].item()},
                "how": f"{nbytes >> 20} MiB pinned copies x{reps}, all {world} ranks concurrently"}
    return {"per_rank_min_gbs": {"h2d": res["h2d"], "d2h": res["d2h"], "duplex_sum": res["duplex"]},
            "how": f"{nbytes >> 20} MiB pinned copies x{reps}"}


def run_e2e(run, steps):
    """the same step through the host-buffer C ABI (what a CPU caller of the reference uses): pinned HOST
    buffers, PCIe copies inside the timed region.  Results larger than 4 GiB per rank are taken in bands of
    z-layers through one pinned result buffer (every byte still crosses PCIe and is counted)."""
    torch, capi, L, wl = run.torch, run.capi, run.L, run.wl
    ne, nq, ncomp, dim = run.ne, run.nq, run.ncomp, run.dim
    per_elem = nq * ncomp * 8 * (1 + (dim if wl["jac"] else 0))
    layer = ne // run.grid.local_n[-1]
    cap = 4 << 30
    band = ne if ne * per_elem <= cap else max(layer, (cap // per_elem) // layer * layer)
    h_coeff = torch.empty(run.ndof_local, dtype=torch.float64, pin_memory=True)
    h_vals = torch.empty((band, nq, ncomp), dtype=torch.float64, pin_memory=True)
    h_jac = torch.empty((band, nq, ncomp, dim), dtype=torch.float64, pin_memory=True) if wl["jac"] else None
    xh = np.ascontiguousarray(run.pts)
    h = run.basis._h

    def e2e_step():
        if run.node_fns is None:
            capi.check(L.dfx_interpolate_host(h, 0, C.byref(run.f), C.c_void_p(h_coeff.data_ptr()), None))
        else:
            for node, fn in run.node_fns:
                capi.check(L.dfx_interpolate_host(h, node, C.byref(fn), C.c_void_p(h_coeff.data_ptr()), None))
        for a in range(0, ne, band):
            b = min(ne, a + band)
            capi.check(L.dfx_evaluate_host(h, 0, C.c_void_p(h_coeff.data_ptr()), xh.ctypes.data_as(C.POINTER(C.c_double)), nq, a, b,
                                           C.c_void_p(h_vals.data_ptr()), C.c_void_p(h_jac.data_ptr()) if h_jac is not None else None))
    n_calls = 1 if run.node_fns is None else len(run.node_fns)
    # interpolate_host: a sub-tree call first brings the vector over (entries outside the subtree keep their
    # value), every call sends it back; evaluate_host: coefficients in, results out
    h2d = run.ndof_local * 8 + (run.ndof_local * 8 * n_calls if run.node_fns is not None else 0)
    d2h = run.ndof_local * 8 * n_calls + ne * per_elem
    e2e_step()
    run.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=run.dev)
    if run.world > 1:
        run.dist.all_reduce(dt, op=run.dist.ReduceOp.MAX)
    out = {"value": run.ndof_global / dt.item(), "unit": "DOFs/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": dt.item() * 1e3, "steps": steps,
           "api": "dfx_interpolate_host + dfx_evaluate_host (pinned host buffers"
                  + ("" if band == ne else f"; results taken in bands of {band} elements through one pinned buffer") + ")"}
    del h_coeff, h_vals, h_jac
    return out


def copy_bandwidth_here(run, seconds=0.3):
    """device-to-device copy bandwidth (read + write bytes) measured on this GPU right after the sustained loop, i.e.
    under the clocks and the power state the step itself ran at — MEASURED_PEAKS.json's figure is a best-of-10 burst.
    Uses two disjoint 1 GiB pieces of the step's own output buffer (recomputed by the next step), larger than the L2."""
    torch = run.torch
    flat = run.values.view(-1)
    n = int(min(1 << 27, flat.numel() // 2))
    if n * 8 < (256 << 20):
        return None
    a, b = flat[:n], flat[n:2 * n]
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    per = 2.0 * n * 8
    reps = int(max(10, min(4000, seconds * 6.0e12 / per)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"gbs": per * reps / ms / 1e6, "seconds": ms / 1e3, "copies": reps, "bytes_per_copy": int(per),
            "how": "torch copy_ between two disjoint 1 GiB pieces of the output buffer, back to back right after the sustained "
                   "loop (same clocks / power state as the step); read + write bytes"}


def time_workload(run, steps, warmup, sustain_s=0.0, sampler=None):
    """W warm-up steps, exactly K timed steps (CUDA events, barrier + synchronize on both sides), optional
    sustained continuation; returns (ms_per_step max over ranks, launches, sustained dict, t_end of the timed steps)"""
    torch, dist = run.torch, run.dist
    for _ in range(warmup):
        run.step(False)
    run.barrier()
    if sampler is not None and run.rank == 0:
        sampler.start()
    launches0 = run.L.dfx_kernel_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        run.step(True)
    stop.record()
    run.barrier()
    t_end = time.perf_counter()
    launches = int(run.L.dfx_kernel_launch_count() - launches0)
    t_ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=run.dev)
    if run.world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / steps
    sustained = None
    if sustain_s > 0:
        n_more = int(min(2000, max(steps, np.ceil(sustain_s * 1e3 / max(ms_per_step, 1e-3)))))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_more):
            run.step(False)
        b.record()
        run.barrier()
        t2 = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=run.dev)
        if run.world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        sustained = {"steps": n_more, "ms_per_step": t2.item() / n_more, "seconds": t2.item() / 1e3}
    return ms_per_step, launches, sustained, t_end


def phase_report(run, peak, per_rank=False):
    """per-phase times (max over ranks), roofline fractions, and — N > 1 — min / median / max over ranks"""
    torch, dist = run.torch, run.dist
    parts = run.phase_ms()
    keys = ("index", "interp", "halo", "eval")
    pt = torch.tensor([parts[k] for k in keys], dtype=torch.float64, device=run.dev)
    ranks = None
    if run.world > 1:
        allp = [torch.empty_like(pt) for _ in range(run.world)]
        dist.all_gather(allp, pt)
        mat = torch.stack(allp).cpu().numpy()          # [rank, phase]
        ranks = {k: {"min": float(mat[:, i].min()), "median": float(np.median(mat[:, i])), "max": float(mat[:, i].max()),
                     "argmax_rank": int(mat[:, i].argmax())} for i, k in enumerate(keys)}
        pt = torch.from_numpy(mat.max(axis=0))
    t_index, t_interp, t_halo, t_eval = [x / 1e3 for x in pt.tolist()]

    def frac(nbytes, t):
        return None if t <= 0 or nbytes == 0 else nbytes / t / 1e9 / peak
    ms = {"interpolate": t_interp * 1e3, "evaluate": t_eval * 1e3, "indices": t_index * 1e3}
    fr = {"evaluate": frac(run.eval_bytes(), t_eval), "interpolate": frac(run.interp_bytes(), t_interp),
          "indices": frac(run.index_bytes(), t_index)}
    return ms, fr, (t_index, t_interp, t_halo, t_eval), ranks


def sub_config(name, local_rank, steps, warmup, peak):
    """one of the other BASELINE configs on this GPU: short timed loop, per-phase roofline fractions,
    extra kernel variants, parity against the oracle"""
    run = Runner(name, 0, 1, local_rank)
    ms_per_step, launches, _, _ = time_workload(run, steps, warmup)
    ms, fr, _, _ = phase_report(run, peak)
    wl = run.wl
    out = {"id": name, "workload": wl["desc"], "function": run.f_name, "dofs": run.ndof_local, "elements": run.ne,
           "steps": steps, "ms_per_step": ms_per_step, "value": run.ndof_local / (ms_per_step / 1e3), "unit": "DOFs/s",
           "gpu_launches": launches, "ms": ms,
           "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "frac": fr["evaluate"], "kernel": "evaluate",
                        "achieved": None if fr["evaluate"] is None else fr["evaluate"] * peak,
                        "algorithmic_bytes": run.eval_bytes(), "traffic": traffic_from_profiles("evaluate", name),
                        "interpolate": {"frac": fr["interpolate"], "algorithmic_bytes": run.interp_bytes(),
                                        "traffic": traffic_from_profiles("interpolate", name)}}}
    if name == "c4":
        # the config's synthetic f costs 24 FP64 operations per node in the reference's summation order (kept for
        # bit-exactness): the per-node interpolation of this config is bound by the FP64 pipe, not by HBM
        out["roofline"]["interpolate"]["second_bound"] = {
            "bound": "fp64 pipe", "sm__pipe_fp64_cycles_active_pct": 80.0,
            "source": "profiles/r02_ncu_configs.csv, k_interp_cell<RegistryFnT<QUADRATIC>,0,4,1>"}
    if run.index_table is not None:
        out["roofline"]["indices"] = {"frac": fr["indices"], "algorithmic_bytes": run.index_bytes(),
                                      "traffic": traffic_from_profiles("indices", name)}
        mi = run.torch.empty((run.ne, run.basis.maxNodeSize(), 2), dtype=run.torch.int32, device=run.dev)
        t = run.time_call(lambda: run.capi.check(run.L.dfx_indices_multi(run.basis._h, 0, run.ne, C.c_void_p(mi.data_ptr()), 2,
                                                                          run.stream_ptr())), 5)
        out["ms"]["indices_multi_2_digits"] = t
        out["roofline"]["indices_multi"] = {"frac": mi.numel() * 4 / (t / 1e3) / 1e9 / peak, "algorithmic_bytes": mi.numel() * 4}
        del mi
    # interpolation with the opt-in per-direction tables (dfx_interpolate_separable), for comparison
    t = run.time_call(lambda: run.interpolate(separable=True), 5)
    out["ms"]["interpolate_separable_tables"] = t
    out["roofline"]["interpolate"]["frac_separable_tables"] = run.interp_bytes() / (t / 1e3) / 1e9 / peak
    if wl["jac"]:
        t = run.time_call(lambda: run.evaluate_range(0, run.ne, jac=False), 3)
        out["ms"]["evaluate_values_only"] = t
        out["roofline"]["evaluate_values_only"] = {"frac": run.eval_bytes(False) / (t / 1e3) / 1e9 / peak,
                                                   "algorithmic_bytes": run.eval_bytes(False)}
    run.step(False)
    out["parity"] = run.parity()
    run.free()
    return out


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: interface-plane exchange over NVLink peer memory (P2P stores + on-stream counters) or NCCL send/recv")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the sub-results for the other BASELINE configs")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--sustain", type=float, default=1.0, help="seconds the step keeps running after the K timed steps")
    ap.add_argument("--cpu-budget", type=float, default=6.0, help="seconds of CPU work per oracle step")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)
    _, f_name = workload_function(wl)

    config = {"workload": wl["desc"], "workload_id": args.workload,
              "grid_global": global_grid(wl, world), "partition": f"z-slabs x{world}",
              "gauss_points_per_element": wl["m"] ** len(wl["n"]), "function": f_name,
              "interpolation": "f evaluated at every Lagrange node (dfx_interpolate)",
              "cache": "inputs and outputs larger than L2 (no flush needed)" if args.workload != "c1"
              else "L2-resident (tiny config)"}

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args, wl, config, warmup)
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa_node(physical_gpu_index(local_rank))
    if world > 1:
        # NCCL writes its version banner to stdout when the first communicator comes up; stdout must
        # carry exactly one JSON line, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    peak, peak_src = peaks()
    run = Runner(args.workload, rank, world, local_rank, args.halo)
    # NVML is polled by rank 0 only: queries from every process contend in the driver and delay launches
    sampler = ClockSampler(physical_gpu_index(local_rank))
    ms_per_step, launches, sustained, t_end = time_workload(run, args.steps, warmup, args.sustain, sampler)
    copy_here = copy_bandwidth_here(run) if args.sustain > 0 else None
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join()
    ms, fr, (t_index, t_interp, t_halo, t_eval), ranks = phase_report(run, peak)

    # ---- other interpolation variant, for the record (not part of `value`)
    t_sep = run.time_call(lambda: run.interpolate(separable=True), 5)

    # ---- correctness of what was just timed
    parity = None
    if not args.no_parity:
        run.step(False)          # every rank: coeff / values hold exactly what a step leaves behind
        run.barrier()
        if rank == 0:
            parity = run.parity()
    halo_check = run.halo_check() if world > 1 else None
    peer_error = run.peer.error() if run.peer is not None else 0

    # ---- gather of the distributed vector (north-star: "NCCL ... gather the result vector"), timed on its own
    gather = None
    if world > 1 and not args.no_gather:
        from dune_functions_b200.slab import GlobalGather
        run.interpolate()
        run.values = run.v_flat = run.jac = run.j_flat = None      # make room: the point values are not needed any more
        torch.cuda.empty_cache()
        free_b, _ = torch.cuda.mem_get_info()
        need = run.ndof_global * 8 + (2 << 30)
        ok_all = torch.tensor([1 if free_b >= need else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        root = None if int(ok_all.item()) == 1 else 0
        try:
            gg = GlobalGather(run.basis, run.ndof_global, root=root)
            out = torch.empty(run.ndof_global, dtype=torch.float64, device=dev) if (root is None or rank == root) else None
            gg(run.coeff, out=out)                                  # warm-up (NCCL connections)
            run.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            gg(run.coeff, out=out)
            b.record()
            run.barrier()
            tg = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            # what arrived from rank (r + 1) % N against that slab's interpolation recomputed here (interpolate
            # is deterministic and needs no exchange), bit for bit; my own ranges against my own vector
            gather_ok = None
            if out is not None:
                import dune_functions_b200 as dfx
                okg = all(bool(torch.equal(out[gb:gb + c], run.coeff[lb:lb + c])) for lb, c, gb in gg.mine)
                nb = (rank + 1) % world
                free_b, _ = torch.cuda.mem_get_info()
                if free_b >= run.ndof_local * 8 + (1 << 30):
                    nbasis = dfx.makeBasis(dfx.YaspGrid(run.ng).slab(nb, world), run.tree, device=local_rank)
                    ref = nbasis.newVector(0.0)
                    if run.node_fns is None:
                        dfx.interpolate(nbasis, ref, run.f)
                    else:
                        for (node, fn) in run.node_fns:
                            run.capi.check(run.L.dfx_interpolate(nbasis._h, node, C.byref(fn), C.c_void_p(ref.data_ptr()), None,
                                                                 run.stream_ptr()))
                    okg = okg and all(bool(torch.equal(out[gb:gb + c], ref[lb:lb + c])) for lb, c, gb in gg.ranges[nb])
                    del ref, nbasis
                gather_ok = okg
            flag = torch.tensor([0 if gather_ok is False else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            gather_ok = bool(flag.item())
            rb = torch.tensor([gg.bytes_received()], dtype=torch.float64, device=dev)
            dist.all_reduce(rb, op=dist.ReduceOp.MAX)
            gather = {"ms": tg.item(), "mode": "all ranks" if root is None else "root 0 only (no room for all-gather)",
                      "bytes_received_per_rank": int(rb.item()), "gbs_per_rank": rb.item() / (tg.item() / 1e3) / 1e9,
                      "check": "ok" if gather_ok else "fail",
                      "check_how": "own ranges + the ranges received from rank (r+1)%N against that slab's interpolation recomputed locally, bit for bit"}
            del out
        except Exception as exc:          # noqa: BLE001
            gather = {"ms": None, "error": str(exc)[:200]}
        torch.cuda.empty_cache()

    # ---- end to end through the host-buffer C ABI
    e2e = None
    if not args.no_e2e:
        pcie = pcie_ceiling(torch, dev, world, dist)
        e2e = run_e2e(run, args.e2e_steps)
        e2e["pcie_ceiling"] = pcie
        t = e2e["ms_per_step"] / 1e3
        lim = max(e2e["h2d_bytes_per_step"] / (pcie["per_rank_min_gbs"]["h2d"] * 1e9),
                  e2e["d2h_bytes_per_step"] / (pcie["per_rank_min_gbs"]["d2h"] * 1e9))
        e2e["frac_of_pcie"] = lim / t
        e2e["frac_of_pcie_note"] = "time the busier PCIe direction needs at the measured pinned-copy rate / step time"

    nd_g, ne, nq = run.ndof_global, run.ne, run.nq
    max_node = run.basis.maxNodeSize()
    eval_bytes, interp_bytes = run.eval_bytes(), run.interp_bytes()
    halo_bytes = int(run.halo.n * 8) if world > 1 else 0
    halo_note = run.halo_note
    run.free()

    # ---- the other BASELINE configs (N = 1 only: one GPU each)
    configs = []
    if world == 1 and not args.no_configs and args.workload == HEADLINE:
        for name in SUB_CONFIGS:
            try:
                configs.append(sub_config(name, local_rank, max(5, min(args.steps, 20)), warmup, peak))
            except Exception as exc:          # noqa: BLE001
                configs.append({"id": name, "error": str(exc)[:300]})

    rc = 0
    if rank == 0:
        achieved = eval_bytes / t_eval / 1e9 if t_eval > 0 else 0.0
        line = {
            "metric": METRIC, "value": nd_g / (ms_per_step / 1e3), "unit": "DOFs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "dofs_global": nd_g, "elements_per_gpu": ne,
            "interpolate_dofs_per_s": nd_g / t_interp if t_interp > 0 else None,
            "basis_fn_evals_per_s": world * ne * nq * max_node / t_eval if t_eval > 0 else None,
            "eval_points_per_s": world * ne * nq / t_eval if t_eval > 0 else None,
            "ms": {"interpolate": t_interp * 1e3, "interpolate_separable_tables": t_sep, "evaluate": t_eval * 1e3,
                   "indices": t_index * 1e3,
                   "evaluate_interior_overlapped_with_halo_exchange": t_halo * 1e3 if world > 1 else None,
                   "halo_bytes_per_rank": halo_bytes, "halo_transport": (halo_note if world > 1 else None),
                   "per_rank": ranks},
            "sustained": sustained,
            "roofline": {"bound": "hbm", "kernel": "evaluate (DiscreteGlobalBasisFunction at Gauss points)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes": eval_bytes,
                         "traffic": traffic_from_profiles("evaluate", args.workload),
                         "copy_bandwidth_after_sustained_loop": copy_here,
                         # evaluate inside the sustained loop (its share of the step as in the timed steps) against the copy
                         # bandwidth measured right after that loop: both under the settled clocks
                         "frac_of_copy_bandwidth_after_sustained_loop":
                             (achieved * ms_per_step / sustained["ms_per_step"] / copy_here["gbs"]) if (copy_here and sustained) else None,
                         "interpolate": {"achieved": interp_bytes / t_interp / 1e9 if t_interp > 0 else None,
                                         "frac": interp_bytes / t_interp / 1e9 / peak if t_interp > 0 else None,
                                         "frac_separable_tables": interp_bytes / (t_sep / 1e3) / 1e9 / peak if t_sep > 0 else None,
                                         "algorithmic_bytes": interp_bytes,
                                         "traffic": traffic_from_profiles("interpolate", args.workload)}},
            "clocks": sampler.report(t_end),
            "gpu_launches": launches, "e2e": e2e, "parity": parity,
        }
        if world > 1:
            line["halo_check"] = halo_check
            line["halo_link_error"] = int(peer_error)
            line["gather"] = gather
            line["gather_ms"] = None if gather is None else gather.get("ms")
        if configs:
            line["configs"] = configs
        if not args.no_cpu and world == 1:
            import orc
            nthreads = orc.lib().orc_max_threads()
            ob, cf, cpts, cnf, n = cpu_sample(wl, args.cpu_budget, nthreads)
            dt, ti, te = cpu_step(ob, cf, cpts, wl["jac"], nthreads, cnf)
            line["cpu_baseline"] = {
                "value": ob.dimension() / dt, "unit": "DOFs/s", "cores": nthreads, "kind": "port",
                "sample": f"{'x'.join(map(str, n))} element box of the workload grid ({ob.dimension()} DOFs), "
                          f"oracle restatement of the reference loops on {nthreads} std::threads",
                "interpolate_dofs_per_s": ob.dimension() / ti,
                "basis_fn_evals_per_s": ob.numElements() * len(cpts) * ob.maxNodeSize() / te}
            # the reference's element loop is serial (interpolate.hh:254): the same port on ONE thread
            ob1, cf, cpts, cnf, n1 = cpu_sample(wl, max(1.0, args.cpu_budget / 2), 1)
            dt1, ti1, te1 = cpu_step(ob1, cf, cpts, wl["jac"], 1, cnf)
            line["cpu_baseline"]["single_thread_dofs_per_s"] = ob1.dimension() / dt1
            line["cpu_baseline"]["single_thread_sample"] = f"{'x'.join(map(str, n1))} element box ({ob1.dimension()} DOFs), 1 thread"
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            print(f"PARITY FAILURE: {parity}", file=sys.stderr)
            rc = 3
        for c in configs:
            if "parity" in c and not c["parity"]["ok"]:
                print(f"PARITY FAILURE in {c['id']}: {c['parity']}", file=sys.stderr)
                rc = 3
    if world > 1:
        bad = torch.tensor([1 if (halo_check == "fail" or peer_error or (gather or {}).get("check") == "fail") else 0],
                           dtype=torch.int32, device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()):
            if peer_error:
                print(f"rank {rank}: halo link wait timed out at step {peer_error}", file=sys.stderr)
            rc = 4
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(rc)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the element-loop hot path.

A "step" is one pass of the hot path over one batch of synthetic input:
    interpolate(basis, coeff, f)                      f(x) = exp(-|x|^2)
    [N > 1: one-plane halo exchange of interface DOFs with the z-neighbours]
    DiscreteGlobalBasisFunction evaluation at the tensor Gauss points of every element
Default workload (BASELINE.json configs[1]): LagrangeBasis Q2 on a 3-d YaspGrid 256^3 per
GPU, 3^3 Gauss points.  With N GPUs the grid is 256 x 256 x (256 N), slab-partitioned in z
(weak scaling); `--workload c5` is configs[4]: 512^3 split over the N ranks (strong).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2]
"""
import argparse
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
    # name: (per-rank or global n, tree, gauss points per direction, jacobians, scaling)
    "c1": dict(n=[64, 64], tree="lagrange1", m=2, jac=False, scaling="weak",
               desc="LagrangeBasis Q1, 2-d YaspGrid 64x64 per GPU, interpolate exp(-|x|^2) + eval at 2^2 Gauss points"),
    "c2": dict(n=[256, 256, 256], tree="lagrange2", m=3, jac=False, scaling="weak",
               desc="LagrangeBasis Q2, 3-d YaspGrid 256^3 per GPU (z-slabs), interpolate exp(-|x|^2) + "
                    "DiscreteGlobalBasisFunction eval at 3^3 Gauss points"),
    "c3": dict(n=[128, 128, 128], tree="taylorhood", m=3, jac=False, scaling="weak",
               desc="Taylor-Hood composite(power<3>(Lagrange<2>), Lagrange<1>), 3-d YaspGrid 128^3 per GPU: "
                    "global multi-index map + interpolation (velocity x, pressure exp(-|x|^2)) + eval"),
    "c4": dict(n=[192, 192, 192], tree="lagrangedg4", m=5, jac=True, scaling="weak",
               desc="LagrangeDGBasis order 4, 3-d YaspGrid 192^3 per GPU, values + gradients at 5^3 Gauss points"),
    "c5": dict(n=[512, 512, 512], tree="lagrange2", m=3, jac=False, scaling="strong",
               desc="LagrangeBasis Q2, 3-d YaspGrid 512^3 slab-partitioned over the GPUs, interpolate + eval at 3^3"),
}


def make_tree(name):
    import dune_functions_b200 as dfx
    return {"lagrange1": lambda: dfx.lagrange(1), "lagrange2": lambda: dfx.lagrange(2),
            "taylorhood": lambda: dfx.taylorHood(), "lagrangedg4": lambda: dfx.lagrangeDG(4)}[name]()


def global_grid(wl, world):
    n = list(wl["n"])
    if wl["scaling"] == "weak":
        n[-1] *= world
    return n


# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
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
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def clocks_report(sampler, n_timed):
    """sm_mhz = median over the samples taken DURING the timed region when there are at least 3 of them,
    otherwise over all samples; sm_mhz_sustained = median over everything (timed region + the untimed
    continuation of the same steps)"""
    r = sampler.result()
    if sampler.ok and sampler.samples:
        r["sm_mhz_sustained"] = float(np.median(sampler.samples))
        if n_timed >= 3:
            r["sm_mhz"] = float(np.median(sampler.samples[:n_timed]))
        r["samples_in_timed_region"] = n_timed
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
def cpu_step(ob, f, pts, jac, nthreads, node_fns=None):
    """one oracle step (interpolate + evaluate) on the sample basis; returns seconds"""
    import dune_functions_b200 as dfx
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
    del dfx
    return t2 - t0, t1 - t0, t2 - t1


def cpu_sample_grid(wl, edge):
    import dune_functions_b200 as dfx
    d = len(wl["n"])
    n = [min(edge, x) for x in wl["n"]]
    # the sample is a box of the workload's grid: same mesh width, same functions
    return dfx.YaspGrid(wl["n"], local_begin=[0] * d, local_n=n), n


def th_node_fns():
    import dune_functions_b200 as dfx
    # expanded Taylor-Hood tree: 0 composite, 1 power, 2..4 Q2 leaves, 5 Q1 leaf
    return [(1, dfx.identity(3)), (5, dfx.gaussian(1.0))]


def run_cpu(wl, wl_name, budget_s, nthreads):
    """oracle timing on a bounded sample; returns dict for cpu_baseline"""
    import orc
    import dune_functions_b200 as dfx
    tree = make_tree(wl["tree"])
    pts, _ = dfx.tensor_gauss(len(wl["n"]), wl["m"])
    f = dfx.gaussian(1.0)
    node_fns = th_node_fns() if wl["tree"] == "taylorhood" else None
    edge = 16
    grid, n = cpu_sample_grid(wl, edge)
    ob = orc.OracleBasis(grid, tree)
    t, _, _ = cpu_step(ob, f, pts, wl["jac"], nthreads, node_fns)
    per_elem = t / ob.numElements()
    # largest cube sample that fits the budget
    for cand in (24, 32, 48, 64, 96, 128, 192, 256):
        ne = float(np.prod([min(cand, x) for x in wl["n"]]))
        if ne * per_elem <= budget_s:
            edge = cand
    grid, n = cpu_sample_grid(wl, edge)
    ob = orc.OracleBasis(grid, tree)
    return ob, f, pts, node_fns, n


# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"],
                    help="N > 1: interface-plane exchange over NVLink peer memory (P2P stores + on-stream counters) or NCCL send/recv")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work per oracle step")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)

    config = {"workload": wl["desc"], "workload_id": args.workload,
              "grid_global": global_grid(wl, world), "partition": f"z-slabs x{world}",
              "gauss_points_per_element": wl["m"] ** len(wl["n"]), "function": "exp(-|x|^2)",
              "cache": "inputs and outputs larger than L2 (no flush needed)" if args.workload != "c1"
              else "L2-resident (tiny config)"}
    metric = "DOFs/s through interpolate + DiscreteGlobalBasisFunction evaluation (FP64)"

    if args.impl == "reference":
        # the reference's own CPU path (restatement oracle: the reference cannot be built here,
        # see DESIGN.md) on all host threads, rank 0 only
        if rank != 0:
            return
        import orc
        nthreads = orc.lib().orc_max_threads()
        ob, f, pts, node_fns, n = run_cpu(wl, args.workload, args.cpu_budget, nthreads)
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
        nq = len(pts)
        sample = f"{'x'.join(map(str, n))} element box of the workload grid per step ({ob.dimension()} DOFs)"
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "DOFs/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "DOFs/s", "cores": nthreads, "kind": "port", "sample": sample,
                                 "interpolate_dofs_per_s": ob.dimension() * args.steps / ti,
                                 "basis_fn_evals_per_s": ob.numElements() * nq * ob.maxNodeSize() * args.steps / te},
                "e2e": {"value": val, "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    import dune_functions_b200 as dfx
    from dune_functions_b200 import _capi as capi
    import ctypes as C

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
    L = capi.lib()

    ng = global_grid(wl, world)
    grid = dfx.YaspGrid(ng).slab(rank, world)
    tree = make_tree(wl["tree"])
    basis = dfx.makeBasis(grid, tree, device=local_rank)
    dim = len(ng)
    pts, _ = dfx.tensor_gauss(dim, wl["m"])
    nq = len(pts)
    f = dfx.gaussian(1.0)
    node_fns = None
    if wl["tree"] == "taylorhood":
        node_fns = [(basis.child(0, 0), dfx.identity(3)), (basis.child(0, 1), dfx.gaussian(1.0))]
    ne = basis.numElements()
    ncomp = basis.numLeaves()
    ndof_local = basis.dimension()
    # whole-job DOF count = dimension of the same basis on the global grid
    ndof_global = dfx.makeBasis(dfx.YaspGrid(ng), tree, device=local_rank).dimension()
    coeff = basis.newVector(0.0)
    values = torch.empty((ne, nq, ncomp), dtype=torch.float64, device=dev)
    jac = torch.empty((ne, nq, ncomp, dim), dtype=torch.float64, device=dev) if wl["jac"] else None
    dgbf = dfx.makeDiscreteGlobalBasisFunction(basis, coeff)
    from dune_functions_b200.slab import HaloExchange, PeerHaloExchange
    halo = HaloExchange(basis, rank, world, device=dev)
    peer, halo_note = None, args.halo
    if world > 1 and args.halo == "peer":
        # every rank must get its link (CUDA IPC between the processes of this node); otherwise all
        # ranks take the NCCL transport together and the JSON line says so
        ok = 1
        try:
            peer = PeerHaloExchange(basis, rank, world)
        except Exception as exc:          # noqa: BLE001
            ok, peer = 0, None
            print(f"rank {rank}: peer-memory halo link unavailable ({exc}); using NCCL", file=sys.stderr)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if peer is not None:
                peer.close()
            peer, halo_note = None, "nccl (peer link unavailable)"
    index_table = None
    if args.workload == "c3":
        index_table = torch.empty((ne, basis.maxNodeSize()), dtype=torch.int32, device=dev)

    ev = {k: [] for k in ("interp", "halo", "eval", "index")}

    def stream_ptr():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # N > 1: the halo exchange runs on a side stream while the main stream evaluates the elements
    # that do not touch the interface plane (all z-layers but the top one); the top layer follows
    # once the plane has arrived.
    layer = ne // grid.local_n[-1]
    split = ne - layer if (world > 1 and rank < world - 1 and halo.n > 0 and grid.local_n[-1] > 1) else ne
    comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
    v_flat = values.view(ne, -1)
    j_flat = jac.view(ne, -1) if jac is not None else None

    def evaluate_range(a, b):
        if b > a:
            dgbf.evaluate(pts, values=True, jacobians=wl["jac"], elem_begin=a, elem_end=b,
                          out_values=v_flat[a:b], out_jacobians=None if j_flat is None else j_flat[a:b])

    def step(timed):
        marks = []

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append(e)
        mark()
        if index_table is not None:
            capi.check(L.dfx_indices_flat(basis._h, 0, ne, C.c_void_p(index_table.data_ptr()), stream_ptr()))
        mark()
        if node_fns is None:
            dfx.interpolate(basis, coeff, f)
        else:
            for node, fn in node_fns:
                capi.check(L.dfx_interpolate(basis._h, node, C.byref(fn), C.c_void_p(coeff.data_ptr()), None, stream_ptr()))
        mark()
        if world > 1 and peer is not None:
            # push my bottom plane into the lower neighbour's buffer right away, evaluate everything that
            # does not need the interface plane, then take the upper neighbour's plane (by now it has
            # usually arrived: the wait is a counter check on the stream) and finish with the top layer
            st = stream_ptr()
            peer.step += 1
            if rank > 0:
                capi.check(L.dfx_halo_push(basis._h, peer._link, C.c_void_p(coeff.data_ptr()), peer.step, st))
            evaluate_range(0, split)
            if rank < world - 1:
                capi.check(L.dfx_halo_pull(basis._h, peer._link, C.c_void_p(coeff.data_ptr()), peer.step, st))
            mark()
            evaluate_range(split, ne)
        elif world > 1:
            main = torch.cuda.current_stream()
            comm_stream.wait_stream(main)
            with torch.cuda.stream(comm_stream):
                halo(coeff)      # my top plane <- bottom plane of the slab above (NCCL send/recv)
            evaluate_range(0, split)
            main.wait_stream(comm_stream)
            mark()
            evaluate_range(split, ne)
        else:
            mark()
            evaluate_range(0, ne)
        mark()
        if timed:
            ev["index"].append((marks[0], marks[1]))
            ev["interp"].append((marks[1], marks[2]))
            ev["halo"].append((marks[2], marks[3]))      # N > 1: interior evaluation overlapped with the exchange
            ev["eval"].append((marks[2], marks[4]))      # the whole evaluation phase (N > 1: includes the exchange)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(False)
    barrier()
    # NVML is polled by rank 0 only: queries from every process contend in the driver and delay launches
    sampler = ClockSampler(physical_gpu_index(local_rank))
    if rank == 0:
        sampler.start()
    launches0 = L.dfx_kernel_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step(True)
    stop.record()
    barrier()
    launches = int(L.dfx_kernel_launch_count() - launches0)
    # the timed region of a short run (20 steps = 18 ms) gives NVML time for very few samples: the same
    # steps keep running, untimed, until the sampler has seen the load for a while
    samples_timed = len(sampler.samples)
    t_extra = time.perf_counter()
    while len(sampler.samples) < 25 and time.perf_counter() - t_extra < 0.5 and sampler.ok and world == 1:
        step(False)
        torch.cuda.synchronize()
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join()
    t_ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    parts = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) if v else 0.0 for k, v in ev.items()}
    pt = torch.tensor([parts["interp"], parts["halo"], parts["eval"], parts["index"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
    ms_per_step = t_ms.item() / args.steps
    t_interp, t_halo, t_eval, t_index = [x / 1e3 for x in pt.tolist()]

    # ---- end to end through the host-buffer C ABI (what a CPU caller of the reference uses)
    e2e = None
    if not args.no_e2e:
        h_coeff = torch.empty(ndof_local, dtype=torch.float64, pin_memory=True)
        h_vals = torch.empty((ne, nq, ncomp), dtype=torch.float64, pin_memory=True)
        h_jac = torch.empty((ne, nq, ncomp, dim), dtype=torch.float64, pin_memory=True) if wl["jac"] else None
        xh = np.ascontiguousarray(pts)

        def e2e_step():
            if node_fns is None:
                capi.check(L.dfx_interpolate_host(basis._h, 0, C.byref(f), C.c_void_p(h_coeff.data_ptr()), None))
            else:
                for node, fn in node_fns:
                    capi.check(L.dfx_interpolate_host(basis._h, node, C.byref(fn), C.c_void_p(h_coeff.data_ptr()), None))
            capi.check(L.dfx_evaluate_host(basis._h, 0, C.c_void_p(h_coeff.data_ptr()),
                                           xh.ctypes.data_as(C.POINTER(C.c_double)), nq, 0, ne,
                                           C.c_void_p(h_vals.data_ptr()),
                                           C.c_void_p(h_jac.data_ptr()) if h_jac is not None else None))
        n_interp_calls = 1 if node_fns is None else len(node_fns)
        h2d = ndof_local * 8 + (ndof_local * 8 * n_interp_calls if node_fns is not None else 0)
        d2h = ndof_local * 8 * n_interp_calls + h_vals.numel() * 8 + (h_jac.numel() * 8 if h_jac is not None else 0)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": ndof_global / dt.item(), "unit": "DOFs/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt.item() * 1e3, "steps": args.e2e_steps,
               "api": "dfx_interpolate_host + dfx_evaluate_host (pinned host buffers)"}
        del h_coeff, h_vals, h_jac

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes of the dominant kernel (evaluate): coefficients read once + outputs written once
        eval_bytes = ndof_local * 8 + ne * nq * ncomp * 8 * (1 + (dim if wl["jac"] else 0))
        interp_bytes = ndof_local * 8
        achieved = eval_bytes / t_eval / 1e9 if t_eval > 0 else 0.0
        line = {
            "metric": metric, "value": ndof_global / (ms_per_step / 1e3), "unit": "DOFs/s", "n_gpus": world,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "dofs_global": ndof_global, "elements_per_gpu": ne,
            "interpolate_dofs_per_s": ndof_global / t_interp if t_interp > 0 else None,
            "basis_fn_evals_per_s": world * ne * nq * basis.maxNodeSize() / t_eval if t_eval > 0 else None,
            "eval_points_per_s": world * ne * nq / t_eval if t_eval > 0 else None,
            "ms": {"interpolate": t_interp * 1e3, "evaluate": t_eval * 1e3, "indices": t_index * 1e3,
                   "evaluate_interior_overlapped_with_halo_exchange": t_halo * 1e3 if world > 1 else None,
                   "halo_bytes_per_rank": int(halo.n * 8) if world > 1 else 0,
                   "halo_transport": (halo_note if world > 1 else None)},
            "roofline": {"bound": "hbm", "kernel": "evaluate (DiscreteGlobalBasisFunction at Gauss points)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes": eval_bytes,
                         "traffic": traffic_from_profiles("evaluate", args.workload),
                         "interpolate": {"achieved": interp_bytes / t_interp / 1e9 if t_interp > 0 else None,
                                         "frac": interp_bytes / t_interp / 1e9 / peak if t_interp > 0 else None,
                                         "algorithmic_bytes": interp_bytes,
                                         "traffic": traffic_from_profiles("interpolate", args.workload)}},
            "clocks": clocks_report(sampler, samples_timed), "gpu_launches": launches, "e2e": e2e,
        }
        if not args.no_cpu and world == 1:
            import orc
            nthreads = orc.lib().orc_max_threads()
            ob, cf, cpts, cnf, n = run_cpu(wl, args.workload, args.cpu_budget, nthreads)
            dt, ti, te = cpu_step(ob, cf, cpts, wl["jac"], nthreads, cnf)
            line["cpu_baseline"] = {
                "value": ob.dimension() / dt, "unit": "DOFs/s", "cores": nthreads, "kind": "port",
                "sample": f"{'x'.join(map(str, n))} element box of the workload grid ({ob.dimension()} DOFs), "
                          f"oracle restatement of the reference loops on {nthreads} std::threads",
                "interpolate_dofs_per_s": ob.dimension() / ti,
                "basis_fn_evals_per_s": ob.numElements() * len(cpts) * ob.maxNodeSize() / te}
            t1 = cpu_step(ob, cf, cpts, wl["jac"], 1, cnf)[0] if ob.numElements() <= 64 ** 3 else None
            if t1:
                line["cpu_baseline"]["single_thread_dofs_per_s"] = ob.dimension() / t1
        print(json.dumps(line))
    if peer is not None:
        err = peer.error()
        if err:
            print(f"rank {rank}: halo link wait timed out at step {err}", file=sys.stderr)
        torch.cuda.synchronize()
        dist.barrier()
        peer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- certified fronts/sec on the bifurcation-diagram parameter sweep.

One "step" = one pass of the hot path (frame propagation + conjugate-point count, the reference's
E_u.frameDet call, heteroclinic.cpp:334) over this rank's shard of the sweep.  Workload at N GPUs =
BASELINE.json configs[1] per GPU: the reference's Construct_ParameterList sweep (heteroclinic.cpp:386-427)
with 128*N angles x 8 radii = 1,024*N parameter values, sharded front_id mod N (weak scaling), one
NCCL all-gather of the result records per step when N > 1.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--fronts-per-gpu M | --total-fronts T]
                  [--workload sweep|subbox4096] [--n 2|3|4]
  (N > 1: launched by torchrun, one rank per GPU)

  --total-fronts T      STRONG scaling: the sweep has T parameter values in total whatever N is (BASELINE configs[2]: T = 65536),
                        each rank takes the fronts f with f mod N == rank
  --workload subbox4096 BASELINE configs[3]: per parameter the x-part of Uxy cut into 16^3 = 4,096 sub-boxes, fused propagate + count,
                        hull / tally on the device (ccp_frame_count_subdivided); --fronts-per-gpu parameters per step (default 4)
  --n 2|4               BASELINE configs[4]: the two-component system / the four-component chain instead of n = 3

`--impl reference` times the CPU restatement of the reference's own path (oracle mode "ref"; the reference
itself needs CAPD and cannot be built here) on the box's host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "certified fronts/sec (param sweep)"
UNIT = "fronts/s"


def sweep_params(pkg, n_total, n=3):
    """Construct_ParameterList with N_c*N_r = n_total (N_r = 8 radii as in config 2 = 128 x 8).  n = 3: the reference's list
    (b1,b2,b3,c12,c23).  BASELINE configs[4]: n = 4 is the chain extension (1,.98,.96,.94,c12,c23,.03) over the same (c12,c23);
    n = 2 is (a,b,c) = (1,.97,c) with c = the list's c12 kept away from the decoupled system (|c| >= 0.02)."""
    import numpy as np
    n_r = 8 if n_total % 8 == 0 else 1
    P = pkg.parameter_list(n_total // n_r, n_r, 0.02, 0.06)
    if n == 3:
        return P
    if n == 4:
        return np.column_stack([np.full(len(P), 1.0), np.full(len(P), .98), np.full(len(P), .96), np.full(len(P), .94), P[:, 3], P[:, 4], np.full(len(P), .03)])
    r = np.hypot(P[:, 3], P[:, 4])
    return np.column_stack([np.full(len(P), 1.0), np.full(len(P), .97), np.where(P[:, 3] >= 0, r, -r)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = False
        self.th = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][2]), "power_w_max": max(float(r[3]) for r in self.rows),
                "reasons": reasons, "samples": len(self.rows)}


def ncu_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one front_kernel<3> launch of 1,024 fronts, from the newest committed
    `ncu --set full` summary under profiles/ (bytes; None if there is none).  Mostly instruction fetch: the algorithmic traffic is
    ~2.9 MB per launch (descriptors in, result records out)."""
    import glob
    import re
    best = None
    for f in glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_front_kernel.txt")):
        m = re.search(r"_v(\d+)_", os.path.basename(f))
        if m and (best is None or int(m.group(1)) > best[0]):
            best = (int(m.group(1)), f)
    if best is None:
        return None
    tot = 0.0
    for ln in open(best[1]):
        m = re.match(r"dram__bytes_(read|write)\.sum \[(\w+)\] = ([\d.,]+)", ln)
        if m:
            tot += float(m.group(3).replace(",", "")) * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[m.group(2)]
    return tot or None


XY_NBD = 1e-12          # +- radius of the neighbourhood of the connecting orbit in x (the Krawczyk stage delivers 1e-13 .. 3e-13)
EIG_ERR = 2e-5          # +- entries of Eu_m_Error_Final
CONE_L = 1.5e-5         # cone constant of the unstable manifold (heteroclinic.cpp:80-87)


def bench_opts(pkg, n=3):
    if n == 2:      # the reference's constants for the two-component system (heteroclinic.cpp:80-87)
        return pkg.abi.SetupOpts.default(xy_nbd=XY_NBD, eig_err=2e-4, cone_L=8e-5)
    return pkg.abi.SetupOpts.default(xy_nbd=XY_NBD, eig_err=EIG_ERR, cone_L=CONE_L)


def cpu_reference_sample(pkg, orc, n_fronts_total, cores, steps=1, warmup=0, mode=None, n_sample=None, n=3):
    """Times an oracle mode on fronts taken evenly from the sweep, one front per host thread at a time.  Default: the
    reference-structured mode (n separate 4n-dim validated integrations + topFrame) on `cores` fronts.  The metric's numerator for the
    CPU arm is FRONTS PROCESSED per second: with the rigorous box of the bench the reference-structured algorithm (no renormalisation
    of the frame basis) certifies none of them (`certified` is reported beside it)."""
    import numpy as np
    mode = orc.MODE_REF if mode is None else mode
    n_sample = cores if n_sample is None else n_sample
    params = sweep_params(pkg, n_fronts_total, n)
    take = np.linspace(0, len(params) - 1, n_sample).round().astype(int)
    descs, _ = pkg.setup_sweep(n, params[take], bench_opts(pkg, n))
    times, certified = [], 0
    for it in range(warmup + steps):
        res, sec = orc.run_batch(pkg.abi, descs, mode, threads=cores)
        if it >= warmup:
            times.append(sec)
            certified = sum(1 for r in res if pkg.frame_det_code(r) >= 0)
    sec = sum(times) / len(times)
    return {"fronts": int(n_sample), "certified": int(certified), "seconds": sec, "value": n_sample / sec, "all_times": times}


def subbox_arm(args, pkg, ge, eng, dev, rank, world, K, W, config, scaling):
    """BASELINE configs[3]: every parameter's box cut into 16^3 = 4,096 sub-boxes (x-part of Uxy), one ccp_frame_count_subdivided call
    per step (expand -> front_kernel over all sub-boxes -> hull / tally, all on the device; host buffers in and out)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    abi = pkg.abi
    n_par = args.fronts_per_gpu
    params = sweep_params(pkg, 1024)[np.linspace(0, 1023, n_par * world).round().astype(int)][rank::world]
    descs, _ = pkg.setup_sweep(3, params, bench_opts(pkg))
    for _ in range(max(W, 1)):
        res = eng.frame_count_subdivided(descs, 16, abi.CCP_SUBDIV_UXY)
    launches0 = eng.kernel_launches()
    ker_ms, wall = [], []
    with ClockSampler(dev.index) as clocks:
        for _ in range(K):
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            res = eng.frame_count_subdivided(descs, 16, abi.CCP_SUBDIV_UXY)
            wall.append(time.perf_counter() - t0)
            ker_ms.append(eng.last_kernel_ms())
    launches = eng.kernel_launches() - launches0
    t = torch.tensor([sum(ker_ms), sum(wall) * 1e3], dtype=torch.float64, device=dev)
    cert = torch.tensor([sum(1 for r in res if r.zero_count >= 0), sum(r.certified for r in res)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cert, op=dist.ReduceOp.SUM)
    n_total = n_par * world
    certified, sub_cert = int(cert[0]), int(cert[1])
    peak = eng.fp64_peak_tflops()
    flop = float(sum(pkg.front_flops(3, 20, 14, int(r.steps)) * r.sub_boxes for r in res))
    achieved = flop / (sum(ker_ms) / K * 1e-3) / 1e12
    clk = clocks.summary()
    nominal = 148 * 64 * 2 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    line = {"metric": METRIC, "value": certified * K / (float(t[0]) * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(W, 1),
            "ms_per_step": float(t[0]) / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64 interval (directed rounding)", "data": "synthetic", "config": config,
            "fronts_total": n_total, "certified": certified, "certified_frac": certified / n_total,
            "sub_boxes_per_front": 4096, "sub_boxes_certified": sub_cert, "sub_boxes_per_s": 4096 * n_total * K / (float(t[0]) * 1e-3),
            "counts": [int(r.zero_count) for r in res],
            "e2e": {"value": certified * K / (float(t[1]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": n_total * C.sizeof(abi.FrontDesc),
                    "d2h_bytes_per_step": n_total * C.sizeof(abi.SubdivResult)},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "peak_nominal": nominal, "frac_of_nominal": achieved / nominal, "kernel": "ccp::front_kernel<3> over the sub-boxes (+ expand_kernel, reduce_kernel)",
                         "kernel_ms": sum(ker_ms) / K, "flop_per_launch": flop}}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fronts-per-gpu", type=int, default=None)
    ap.add_argument("--total-fronts", type=int, default=None)
    ap.add_argument("--workload", default="sweep", choices=["sweep", "subbox4096"])
    ap.add_argument("--n", type=int, default=3, choices=[2, 3, 4])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = max(1, args.steps), max(0, args.warmup)
    strong = args.total_fronts is not None
    if args.fronts_per_gpu is None:
        args.fronts_per_gpu = 4 if args.workload == "subbox4096" else 1024
    n_total = args.total_fronts if strong else args.fronts_per_gpu * world
    scaling = "strong" if strong else "weak"
    cores = os.cpu_count() or 1
    N = args.n

    import __graft_entry__ as ge
    pkg = ge.load_package()
    config = {"workload": ("bifurcation-diagram sweep, Construct_ParameterList %dx8 = %d parameter values (%s), n=%d, %s, order 20, step 2^-5, grid 14"
                           % (n_total // 8, n_total, "BASELINE configs[2], fixed total" if strong else "BASELINE configs[1] x n_gpus", N,
                              "b=(1,.975,.95)" if N == 3 else ("(a,b)=(1,.97)" if N == 2 else "b=(1,.98,.96,.94), c34=.03")))
                          if args.workload == "sweep" else
                          "BASELINE configs[3]: %d parameter values per step, each with the x-part of Uxy cut into 16^3 = 4,096 sub-boxes (part() rule, utils.cpp:3-10), fused propagate + count, hull / tally on the device, n=3" % n_total,
              "certified": "a front counts as certified when the hot path returns status OK and failure_count == 0 (SURVEY.md 8d); the proof stages (i), (ii), (iv) are host code outside the timed path, descriptors come from float shooting with the boxes below",
              "xy_nbd": XY_NBD, "eig_err": EIG_ERR, "cone_L": CONE_L,
              "fronts_per_gpu": args.fronts_per_gpu, "parallelism": "front_id mod %d shards, 1 NCCL all-gather of result records per step" % world if world > 1 else "single GPU",
              "l2": "256 MiB written between timed steps (untimed); inputs are ~1.2 MB and the kernel is FP64-compute bound"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        orc = ge.load_oracle()
        s = cpu_reference_sample(pkg, orc, n_total if args.workload == "sweep" else 1024, cores, steps=K, warmup=min(W, 1), n=N)
        line = {"metric": METRIC, "value": s["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": min(W, 1),
                "ms_per_step": s["seconds"] * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64 interval (directed rounding)", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": s["value"], "unit": UNIT, "cores": cores, "kind": "port", "certified_of_sample": s["certified"],
                                 "sample": "%d fronts of the sweep (evenly strided), one per host thread, oracle mode ref; value = fronts PROCESSED per second (with the rigorous box xy_nbd = %g the reference-structured algorithm certifies %d of them); reference binary needs CAPD 5.0.59 (unbuildable here)" % (s["fronts"], XY_NBD, s["certified"])},
                "e2e": {"value": s["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import numpy as np
    import torch
    import torch.distributed as dist
    from ccp_b200 import sweep
    abi = pkg.abi
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    eng = pkg.Engine(local_rank)

    if args.workload == "subbox4096":
        return subbox_arm(args, pkg, ge, eng, dev, rank, world, K, W, config, scaling)
    params = sweep_params(pkg, n_total, N)
    mine = sweep.local_indices(n_total, rank, world)
    descs, infos = pkg.setup_sweep(N, params[mine], bench_opts(pkg, N))
    m = len(descs)
    rb, db = C.sizeof(abi.FrontResult), C.sizeof(abi.FrontDesc)
    d_in = sweep.desc_array_to_tensor(descs, device=dev)
    d_out = torch.zeros(m * rb, dtype=torch.uint8, device=dev)
    h_in = sweep.desc_array_to_tensor(descs, pin=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    peak = eng.fp64_peak_tflops()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    gathered = None
    # the single collective of the path writes straight into a preallocated [world, m_pad, record] buffer: no zero fill, no permuting
    # copy on the device (front f sits at [f % world, f // world]; the host puts the records in front order after the timed region)
    m_pad = sweep.padded_shard_len(n_total, world)
    g_out = torch.empty(world * m_pad * rb, dtype=torch.uint8, device=dev) if world > 1 else None
    if m_pad != m:                              # ragged shard: this rank's buffer is padded once, outside the timed region
        d_out = torch.zeros(m_pad * rb, dtype=torch.uint8, device=dev)

    def step():
        nonlocal gathered
        with torch.cuda.stream(stream):
            eng.frame_count_batch_device(d_in.data_ptr(), m, d_out.data_ptr(), stream.cuda_stream, n=N)
            gathered = sweep.gather_records(d_out, n_total, rank, world, rb, out=g_out, ordered=False)

    for _ in range(W):
        step()
    barrier()
    launches0 = eng.kernel_launches()
    ker_ms, step_ms = [], []
    with ClockSampler(local_rank) as clocks:
        for _ in range(K):
            flush.fill_(1)                       # L2 flush, untimed
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(stream):
                e0.record(stream)
                eng.frame_count_batch_device(d_in.data_ptr(), m, d_out.data_ptr(), stream.cuda_stream, n=N)
                e1.record(stream)
                gathered = sweep.gather_records(d_out, n_total, rank, world, rb, out=g_out, ordered=False)
                e2.record(stream)
            stream.synchronize()
            barrier()
            ker_ms.append(e0.elapsed_time(e1))
            step_ms.append(e0.elapsed_time(e2))
    launches = eng.kernel_launches() - launches0
    t = torch.tensor([sum(step_ms), sum(ker_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_total_ms = float(t[0]), float(t[1])

    ordered = sweep.front_order(gathered, n_total)                  # host, front order
    hdr = ordered[:, :32].copy().view(np.int32)
    st, zc, fc, steps_arr = hdr[:, 0], hdr[:, 1], hdr[:, 2], hdr[:, 4]
    certified = int(((st == 0) & (fc == 0)).sum())
    uncertified = int(n_total - certified)
    value = certified * K / (total_ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (H2D + kernel + D2H inside), plus the gather when N > 1
    res_host = (abi.FrontResult * m)()
    h_ptr = C.cast(h_in.data_ptr(), C.POINTER(abi.FrontDesc))

    def e2e_step():
        rc = eng.L.ccp_frame_count_batch(h_ptr, m, res_host)
        assert rc == 0, eng.L.ccp_last_error()
        if world > 1:
            loc = torch.frombuffer(res_host, dtype=torch.uint8).to(dev)
            if m_pad != m:
                loc = torch.cat([loc, torch.zeros((m_pad - m) * rb, dtype=torch.uint8, device=dev)])
            return sweep.gather_records(loc, n_total, rank, world, rb, out=g_out, ordered=False).cpu()
        return None

    for _ in range(min(W, 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = certified * K / float(e2e_s[0])

    # ---- roofline of the dominant kernel: counted FP64 flop of the implemented algorithm / CUDA-event time
    flop_local = float(sum(pkg.front_flops(N, 20, 14, int(s)) for s in steps_arr[mine]))
    achieved = flop_local / (sum(ker_ms) / K * 1e-3) / 1e12
    clk = clocks.summary()
    nominal = 148 * 64 * 2 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12          # 64 DFMA lanes per SM
    roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": ncu_dram_traffic(),
                "peak_nominal": nominal, "frac_of_nominal": achieved / nominal,
                "peak_source": "measured live on this GPU: 8 independent DFMA.RM/RP chains per thread (ccp_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry (nominal 37 TF)",
                "kernel": "ccp::front_kernel<%d>" % N, "kernel_ms": sum(ker_ms) / K, "flop_per_launch": flop_local,
                "note": "FP64 vector-pipe bound: ~1.2 MB in / 1.7 MB out per launch vs ~2e11 flop; DFMA=2 flop, DADD/DMUL=1 (DESIGN.md 5)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64 interval (directed rounding)", "data": "synthetic", "config": config,
            "fronts_total": n_total, "certified": certified, "uncertified": uncertified, "certified_frac": certified / n_total,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m * db * world, "d2h_bytes_per_step": m * rb * world},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        orc = ge.load_oracle()
        s = cpu_reference_sample(pkg, orc, n_total, cores, n=N)
        line["cpu_baseline"] = {"value": s["value"], "unit": UNIT, "cores": cores, "kind": "port", "certified_of_sample": s["certified"],
                                "sample": "%d fronts of the same sweep (evenly strided), one per host thread, %.1f s wall; oracle mode ref = reference-structured restatement (CAPD absent); value = fronts PROCESSED per second: with the rigorous box it certifies %d of them" % (s["fronts"], s["seconds"], s["certified"])}
        # the SAME algorithm as the kernel (oracle c2, its serial twin) on all host cores: separates the algorithmic factor from the hardware one
        s2 = cpu_reference_sample(pkg, orc, n_total, cores, mode=orc.MODE_C2, n_sample=min(n_total, 32 * cores), n=N)
        line["cpu_baseline_same_algorithm"] = {"value": s2["value"], "unit": UNIT, "cores": cores, "kind": "port", "certified_of_sample": s2["certified"],
                                               "sample": "%d fronts of the same sweep, oracle mode c2 (serial twin of the kernel), %.1f s wall" % (s2["fronts"], s2["seconds"])}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())

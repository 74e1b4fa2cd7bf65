#!/usr/bin/env python
"""bench.py -- particle-steps/s of the per-step SPH core (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (libsph_b200.so)
  python bench.py --impl reference [...]                        the CPU restatement of the reference's Julia
                                                                path (oracle/, OpenMP over all host cores) --
                                                                the Julia reference itself cannot run here.
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] -- Boss-Bodenheimer rotating isothermal cloud, N = 1e6
particles, Kh = 50, theta = 0.576 (F/iniconds.jl:457-525 distributions, numpy default_rng(42), T = 10 K).
One step = one iteration of `while t < tEnd` (F/isothermal_sim.jl:152-213): two full force evaluations
(sort, octree, exact kNN, density, force, tree walk) + dt + statistics + predictor + corrector.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic (compulsory) FP64 bytes per particle per force evaluation, SURVEY.md section 8(d), Kh = 50
ALG_BYTES = {"knn": 232.0, "density": 232.0, "force": 296.0, "gravity": 123.0}
ALG_BYTES_STEP = 2.0 * sum(ALG_BYTES.values()) + 240.0   # + integrator streams


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(n):
    import astrophysical_sph_b200.iniconds as ic

    d = ic.make_ic("isothermal", "boss_bodenheimer", N=n, T=10)
    return d["pos"], d["vel"], d["constants"]


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        out = self.p.communicate()[0]
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Cores this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_steps(pos, vel, c, nsteps, nthreads):
    """nsteps loop iterations on the oracle (all phases, same code path the parity tests check against)."""
    from oracle import oracle as O

    t0 = time.perf_counter()
    O.step(pos, vel, c["m"], c["Kh"], c["G"], c["theta"], 0.0, nsteps, eos=O.ISOTHERMAL, cs=c["cs"], alpha=c["alpha"],
           beta=c["beta"], U_iso=c["U"], nthreads=nthreads)
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    nt = host_threads()
    total = args.steps + args.warmup
    # bounded sample: the same IC family at a particle count that keeps the whole run within a few minutes
    n = args.n if total <= 8 else max(100_000, int(args.n * 8 / total) // 1000 * 1000)
    n = min(n, args.n)
    pos, vel, c = make_workload(n)
    if args.warmup:
        cpu_reference_steps(pos, vel, c, min(args.warmup, 1), nt)
    dt = cpu_reference_steps(pos, vel, c, args.steps, nt)
    val = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "particle_steps_per_s", "value": val, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"boss_bodenheimer isothermal N={n} Kh=50 theta=0.576 (CPU sample of the N={args.n} workload)",
                   "N": n, "Kh": 50},
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": nt, "kind": "port",
                         "sample": f"{args.steps} full steps at N={n} (warmup capped at 1 step), OpenMP x{nt}; "
                                   "Julia reference not runnable (no Julia toolchain)"},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch

    from astrophysical_sph_b200.libsph import SphB200, launch_count

    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)
    n = args.n
    pos, vel, c = make_workload(n)
    stream = torch.cuda.Stream(device=dev)
    s = SphB200(n, c["Kh"], "isothermal", m=c["m"], cs=c["cs"], G=c["G"], theta=c["theta"], alpha=c["alpha"],
                beta=c["beta"], U_iso=c["U"], device=dev)
    s.set_stream(stream.cuda_stream)
    if world > 1:
        box = [SphB200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        s.comm_init(world, rank, box[0])

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident stepping ("value")
    s.upload(pos, vel, None, 0.0)
    if args.warmup:
        s.step(args.warmup, want_info=False)
    barrier()
    sampler = ClockSampler(dev) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launch_count()
    with torch.cuda.stream(stream):
        e0.record(stream)
        s.step(args.steps, want_info=False)
        e1.record(stream)
    barrier()
    launches = launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    tim = s.timings()      # phases of the last force evaluation inside the timed region
    value = n * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with host buffers ("e2e")
    hp = torch.empty((3, n), dtype=torch.float64).pin_memory()
    hv = torch.empty((3, n), dtype=torch.float64).pin_memory()
    hp_np, hv_np = hp.numpy().T, hv.numpy().T            # (N, 3) Fortran-ordered views of pinned memory
    hp_np[...] = pos; hv_np[...] = vel
    import ctypes as C

    from astrophysical_sph_b200.libsph import lib as _lib

    L = _lib()
    tt = C.c_double(0.0)

    def e2e_step():
        s._chk(L.sph_upload(s._h, C.c_void_p(hp.data_ptr()), C.c_void_p(hv.data_ptr()), None, C.c_double(tt.value)))
        s._chk(L.sph_step(s._h, 1, None))
        s._chk(L.sph_download(s._h, C.c_void_p(hp.data_ptr()), C.c_void_p(hv.data_ptr()), None, C.byref(tt)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = n * args.steps / e2e_s

    if rank != 0:
        s.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    phases = {}
    for k in ("knn", "density", "force", "gravity"):
        t_ms = tim[k + "_ms"]
        nt_targets = n / world
        gbs = ALG_BYTES[k] * nt_targets / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        phases[k] = {"ms": round(t_ms, 4), "alg_bytes_per_particle": ALG_BYTES[k], "achieved_gbs": round(gbs, 2),
                     "frac": round(gbs / peak, 5)}
    for k in ("sort", "tree", "finish", "total"):
        phases[k] = {"ms": round(tim[k + "_ms"], 4)}
    dom = max(("knn", "density", "force", "gravity"), key=lambda k: phases[k]["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dom)
    sph_ms = phases["density"]["ms"] + phases["force"]["ms"]
    sph_gbs = (ALG_BYTES["density"] + ALG_BYTES["force"]) * (n / world) / (sph_ms * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nt = host_threads()
        dtc = cpu_reference_steps(pos, vel, c, 1, nt)
        cpu = {"value": n / dtc, "unit": "particle-steps/s", "cores": nt, "kind": "port",
               "sample": f"1 full step of the same N={n} workload on the oracle (C++ restatement of the Julia path, "
                         f"OpenMP x{nt}), {dtc:.1f} s"}

    line = {
        "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"boss_bodenheimer isothermal N={n} Kh=50 theta=0.576 T=10K (BASELINE.json configs[2])",
                   "N": n, "Kh": 50, "theta": 0.576, "force_evals_per_step": 2,
                   "parallelism": "single GPU" if world == 1 else f"targets split by Morton-key range over {world} ranks, "
                                                                   "replicated state, NCCL all-gather/all-reduce",
                   "l2": "working set (neighbour lists 200 MB + tree 150 MB + state) exceeds the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": 48 * n + 8,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "sph_upload + sph_step(1) + sph_download on pinned host buffers, every step"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": {"knn": "knn_quad_kernel", "gravity": "walk_pairs_kernel", "force": "force_kernel",
                                                "density": "density_kernel"}[dom],
                     "achieved": phases[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": phases[dom]["frac"],
                     "traffic": traffic, "peak_source": peak_src,
                     "note": "algorithmic bytes per SURVEY.md 8(d); these kernels are FP64-pipe / latency bound, not HBM bound "
                             "(DESIGN.md section 4); fp64 = SURVEY 8(d) algorithmic flops (45 per node visit, 904 visits per "
                             "particle at this N) against the nominal 37 TFLOP/s FP64 peak",
                     "fp64": {"gravity_tflops": round(45.0 * 904.0 * (n / world) / (phases["gravity"]["ms"] * 1e-3) / 1e12, 3),
                              "peak_nominal_tflops": 37.0,
                              "frac": round(45.0 * 904.0 * (n / world) / (phases["gravity"]["ms"] * 1e-3) / 1e12 / 37.0, 4)}},
        "sph_sums": {"ms": round(sph_ms, 4), "achieved_gbs": round(sph_gbs, 2), "frac": round(sph_gbs / peak, 5),
                     "alg_bytes_per_particle": ALG_BYTES["density"] + ALG_BYTES["force"]},
        "step_alg_bytes": {"per_particle_step": ALG_BYTES_STEP,
                           "achieved_gbs": round(ALG_BYTES_STEP * value / 1e9, 2),
                           "frac": round(ALG_BYTES_STEP * value / 1e9 / peak, 5)},
        "phases_last_eval": phases,
        "knn_retries": tim.get("knn_retries"),
        "comm_ms_last_eval": tim.get("comm_ms"),
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    s.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # replicas are not what the metric asks for: re-launch under torchrun so that ranks share one job
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), __file__,
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--n", str(args.n)]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

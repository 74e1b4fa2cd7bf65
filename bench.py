#!/usr/bin/env python
"""bench.py -- particle-steps/s of the per-step SPH core (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (libsph_b200.so)
  python bench.py --impl reference [...]                        the CPU restatement of the reference's Julia
                                                                path (oracle/, OpenMP over all host cores) --
                                                                the Julia reference itself cannot run here.
  python bench.py --config c0|c1|c2|c3|c4 [--scaling weak]      the other BASELINE.json configs (c2 = default)
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (config.workload): BASELINE.json configs[2] -- Boss-Bodenheimer rotating isothermal cloud, N = 1e6
particles, Kh = 50, theta = 0.576 (F/iniconds.jl:457-525 distributions, numpy default_rng(42), T = 10 K).
One step = one iteration of `while t < tEnd` (F/isothermal_sim.jl:152-213): two full force evaluations
(sort, octree, exact kNN, density, force, tree walk) + dt + statistics + predictor + corrector.
Prints ONE JSON line (rank 0).  With several ranks a 20 000-particle run is first compared with the oracle
("parity" in the line; a violation makes the exit code non-zero).
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

# BASELINE.json configs -> (EOS, ic_type, N, extra iniconds arguments)
CONFIGS = {
    "c0": ("polytropic", "gaussian_sphere", 5_000, dict(R=5.38552341e16)),
    "c1": ("polytropic", "sample_plummer_sphere", 100_000, {}),
    "c2": ("isothermal", "boss_bodenheimer", 1_000_000, dict(T=10)),
    "c3": ("isothermal", "turbulent_molecular_cloud", 4_000_000, dict(T=10)),
    "c4": ("isothermal", "bonnor_ebert_sphere", 16_000_000, dict(T=10)),
}

# algorithmic (compulsory) FP64 bytes per particle per force evaluation, SURVEY.md section 8(d), Kh = 50
ALG_BYTES = {"knn": 232.0, "density": 232.0, "force": 296.0, "gravity": 123.0}
# algorithmic FP64 flops, SURVEY.md section 8(d): 45 per node visit of the walk; density 16 Kh, force 60 x 2 x (Kh - 2)
FLOPS_PER_VISIT = 45.0
FLOPS_SPH = {"density": 16.0 * 50, "force": 60.0 * 2 * 48}


def alg_bytes(eos):
    b = dict(ALG_BYTES)
    if eos == "polytropic":
        b["force"] = 312.0
    return b


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(cfg, n):
    import astrophysical_sph_b200.iniconds as ic

    eos, ic_type, _, kw = CONFIGS[cfg]
    d = ic.make_ic(eos, ic_type, N=n, **kw)
    return eos, d["pos"], d["vel"], d["K"], d["constants"]


def workload_name(cfg, n):
    eos, ic_type, n0, _ = CONFIGS[cfg]
    return f"{ic_type} {eos} N={n} Kh=50 theta=0.576 (BASELINE.json configs[{cfg[1]}]{'' if n == n0 else ', N overridden'})"


def sph_args(eos, c):
    a = dict(m=c["m"], G=c["G"], theta=c["theta"], alpha=c["alpha"], beta=c["beta"])
    if eos == "isothermal":
        a.update(cs=c["cs"], U_iso=c["U"])
    else:
        a.update(gamma=c["gamma"])
    return a


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


def cpu_reference_steps(eos, pos, vel, K, c, nsteps, nthreads):
    """nsteps loop iterations on the oracle (all phases, same code path the parity tests check against)."""
    from oracle import oracle as O

    kw = dict(eos=O.ISOTHERMAL, cs=c["cs"], U_iso=c["U"]) if eos == "isothermal" else dict(eos=O.POLYTROPIC, Kent=K, gamma=c["gamma"])
    t0 = time.perf_counter()
    O.step(pos, vel, c["m"], c["Kh"], c["G"], c["theta"], 0.0, nsteps, alpha=c["alpha"], beta=c["beta"], nthreads=nthreads, **kw)
    return time.perf_counter() - t0


def total_n(args, world):
    n0 = CONFIGS[args.config][2]
    if args.scaling == "weak":
        per = args.n if args.n else max(n0 // 8, 1000)
        return per * world
    return args.n if args.n else n0


def run_reference(args, rank, world):
    """The reference arm: the oracle (C++ restatement of the Julia path) on all host cores, SAME workload and N as the
    GPU arm.  Each step is a full step (about 14 s at N = 1e6 on 16 cores); only the warm-up is capped at one step."""
    if rank != 0:
        return
    nt = host_threads()
    n = total_n(args, world)
    eos, pos, vel, K, c = make_workload(args.config, n)
    if args.warmup:
        cpu_reference_steps(eos, pos, vel, K, c, 1, nt)
    dt = cpu_reference_steps(eos, pos, vel, K, c, args.steps, nt)
    val = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "particle_steps_per_s", "value": val, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, n), "N": n, "Kh": 50, "theta": 0.576, "force_evals_per_step": 2},
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": nt, "kind": "port",
                         "sample": f"{args.steps} full steps at N={n} (warm-up capped at 1 step), OpenMP x{nt}; "
                                   "Julia reference not runnable (no Julia toolchain)"},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import ctypes as C

    import torch

    from astrophysical_sph_b200 import libsph
    from astrophysical_sph_b200.libsph import SphB200, launch_count

    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)

    # ---- multi-GPU parity against the oracle before anything is timed
    parity = None
    if world > 1:
        from tools.mgpu_check import run_check

        parity, ok = run_check(dist, rank, world, dev, 20000)
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": "particle_steps_per_s", "n_gpus": world, "parity": parity,
                                  "error": "multi-GPU parity against the oracle FAILED"}), flush=True)
            dist.destroy_process_group()
            sys.exit(1)

    n = total_n(args, world)
    eos, pos, vel, K, c = make_workload(args.config, n)
    stream = torch.cuda.Stream(device=dev)
    s = SphB200(n, c["Kh"], eos, device=dev, **sph_args(eos, c))
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
    s.upload(pos, vel, K, 0.0)
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
    try:
        tim = s.timings()  # phases of the last force evaluation inside the timed region
    except libsph.SphError:
        s.eval_state()     # steps replayed as a CUDA graph record no phase timers: time one more evaluation
        tim = s.timings()
    value = n * args.steps / (ms * 1e-3)
    per_rank = None
    if dist is not None:
        # the phases of every rank (the JSON line carries rank 0's as `phases_last_eval`): shows which rank the
        # all-gathers wait for
        mine = {k: round(tim[k + "_ms"], 3) for k in ("knn", "walk_kernel", "gravity", "density", "force", "total")}
        mine["comm"] = round(tim.get("comm_ms") or 0.0, 3)
        box = [None] * world
        dist.all_gather_object(box, mine)
        per_rank = box

    # ---- end to end through the public API with host buffers ("e2e"): every rank passes the full pinned arrays,
    # the library moves 1/world of them over PCIe per rank and exchanges the slices over NVLink; rank 0 reads the result
    npoly = 1 if eos == "polytropic" else 0
    hp = torch.empty((3, n), dtype=torch.float64).pin_memory()
    hv = torch.empty((3, n), dtype=torch.float64).pin_memory()
    hk = torch.empty((n,), dtype=torch.float64).pin_memory() if npoly else None
    hp.numpy().T[...] = pos
    hv.numpy().T[...] = vel             # (N, 3) Fortran-ordered views of pinned memory
    if npoly:
        hk.numpy()[...] = K
    L = libsph.lib()
    tt = C.c_double(0.0)
    kp = C.c_void_p(hk.data_ptr()) if npoly else None
    down = rank == 0

    def e2e_step():
        s._chk(L.sph_upload(s._h, C.c_void_p(hp.data_ptr()), C.c_void_p(hv.data_ptr()), kp, C.c_double(tt.value)))
        s._chk(L.sph_step(s._h, 1, None))
        s._chk(L.sph_download(s._h, C.c_void_p(hp.data_ptr()) if down else None, C.c_void_p(hv.data_ptr()) if down else None,
                              kp if down else None, C.byref(tt)))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_val = n * args.steps / e2e_s

    # ---- node visits per particle of the walk (algorithmic flops of the dominant kernel): one counted evaluation on
    # a second handle (rank 0, one GPU's share is the same fraction of it)
    # The same handle runs density / force BEFORE the walk on one stream (SPH_FLAG_SERIAL_PHASES): in the timed run they
    # share the SMs with the walk on a second stream and their phase timers stretch over it, so the SPH-sum figures
    # (`sph_sums`, `phases_alone`) come from here - each kernel alone, second (steady-state) evaluation.
    visits = None
    alone = None
    fp64_peak = None
    if rank == 0:
        fp64_peak = libsph.measure_fp64_peak(dev)
    if world == 1:
        s2 = SphB200(n, c["Kh"], eos, device=dev, flags=libsph.FLAG_COUNT_VISITS | libsph.FLAG_SERIAL_PHASES, **sph_args(eos, c))
        s2.upload(pos, vel, K, 0.0)
        s2.eval_state()
        s2.eval_state()
        t2 = s2.timings()
        visits = t2["walk_visits"] / n
        alone = {k: t2[k + "_ms"] for k in ("knn", "density", "force")}
        s2.close()

    if rank != 0:
        s.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    AB = alg_bytes(eos)
    nt_targets = n / world
    phases = {}
    for k in ("knn", "density", "force", "gravity"):
        t_ms = tim[k + "_ms"]
        gbs = AB[k] * nt_targets / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        phases[k] = {"ms": round(t_ms, 4), "alg_bytes_per_particle": AB[k], "achieved_gbs": round(gbs, 2),
                     "frac_hbm": round(gbs / peak, 5)}
    for k in ("sort", "tree", "finish", "total", "walk_kernel"):
        phases[k] = {"ms": round(tim[k + "_ms"], 4)}
    dom = max(("knn", "density", "force", "gravity"), key=lambda k: phases[k]["ms"])
    kernel_of = {"knn": "knn_quad_kernel", "gravity": "walk_pairs_kernel", "force": "force_kernel", "density": "density_kernel"}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dom)
    phases_alone = None
    if alone is not None:
        phases_alone = {}
        for k, t_ms in alone.items():
            gbs = AB[k] * nt_targets / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
            phases_alone[k] = {"ms": round(t_ms, 4), "alg_bytes_per_particle": AB[k], "achieved_gbs": round(gbs, 2),
                               "frac_hbm": round(gbs / peak, 5)}
    sph_src = phases_alone if phases_alone is not None else phases
    sph_ms = sph_src["density"]["ms"] + sph_src["force"]["ms"]
    sph_gbs = (AB["density"] + AB["force"]) * nt_targets / (sph_ms * 1e-3) / 1e9
    sph_tf = (FLOPS_SPH["density"] + FLOPS_SPH["force"]) * nt_targets / (sph_ms * 1e-3) / 1e12

    # roofline of the dominant kernel.  The tree walk is bound by the FP64 pipe / instruction issue, not by HBM
    # (SURVEY.md 8d): `bound` names what binds, the HBM fraction on compulsory bytes is reported beside it.
    if dom == "gravity":
        k_ms = tim["walk_kernel_ms"] if tim["walk_kernel_ms"] > 0 else tim["gravity_ms"]
        v = visits if visits is not None else 904.0
        tf = FLOPS_PER_VISIT * v * nt_targets / (k_ms * 1e-3) / 1e12
        gbs = AB["gravity"] * nt_targets / (k_ms * 1e-3) / 1e9
        roof = {"bound": "fp64", "kernel": kernel_of[dom], "achieved": round(tf, 3), "peak": round(fp64_peak, 2), "unit": "TFLOP/s",
                "frac": round(tf / fp64_peak, 4), "traffic": traffic, "kernel_ms": round(k_ms, 4),
                "alg_flops_per_particle": FLOPS_PER_VISIT * v, "node_visits_per_particle": round(v, 2),
                "peak_source": "measured: sph_measure_fp64_peak (own DFMA microbenchmark on this GPU, this run)",
                "hbm": {"achieved": round(gbs, 2), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 5),
                        "alg_bytes_per_particle": AB["gravity"], "peak_source": peak_src},
                "note": "algorithmic flops and bytes per SURVEY.md 8(d): 45 flop per node visit x visits counted by the "
                        "kernel itself (second handle, SPH_FLAG_COUNT_VISITS), 123 B per particle"}
    else:
        roof = {"bound": "hbm", "kernel": kernel_of[dom], "achieved": phases[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": phases[dom]["frac_hbm"], "traffic": traffic, "peak_source": peak_src}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        nt = host_threads()
        ncpu = min(n, 1_000_000)
        if ncpu == n:
            ce, cp, cv, cK, cc = eos, pos, vel, K, c
        else:
            ce, cp, cv, cK, cc = make_workload(args.config, ncpu)
        dtc = cpu_reference_steps(ce, cp, cv, cK, cc, 1, nt)
        cpu = {"value": ncpu / dtc, "unit": "particle-steps/s", "cores": nt, "kind": "port",
               "sample": f"1 full step of the N={ncpu} workload on the oracle (C++ restatement of the Julia path, "
                         f"OpenMP x{nt}), {dtc:.1f} s"}

    h2d = (48 + 8 * npoly) * n // world
    line = {
        "metric": "particle_steps_per_s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, n), "N": n, "Kh": 50, "theta": 0.576, "force_evals_per_step": 2,
                   "parallelism": "single GPU" if world == 1 else f"targets split by Morton-key range over {world} ranks, "
                                                                   "replicated state, NCCL all-gathers",
                   "l2": "working set (neighbour lists 200 B + tree 190 B + state per particle) exceeds the 126 MB L2 "
                         "at N >= 1e6; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": (48 + 8 * npoly) * n + 8,
                "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "sph_upload + sph_step(1) + sph_download on pinned host buffers, every step"
                       + ("" if world == 1 else f"; per rank 1/{world} of the upload over PCIe, slices exchanged over NVLink; rank 0 downloads")},
        "gpu_launches": launches,
        "roofline": roof,
        "fp64_peak_tflops": round(fp64_peak, 2),
        "sph_sums": {"ms": round(sph_ms, 4), "achieved_gbs": round(sph_gbs, 2), "frac_hbm": round(sph_gbs / peak, 5),
                     "alg_bytes_per_particle": AB["density"] + AB["force"],
                     "achieved_tflops": round(sph_tf, 3), "frac_fp64": round(sph_tf / fp64_peak, 4),
                     "alg_flops_per_particle": FLOPS_SPH["density"] + FLOPS_SPH["force"],
                     "how": ("density + force phases each timed alone: second handle with SPH_FLAG_SERIAL_PHASES (`phases_alone`); "
                             "in the timed run they execute on a second stream beside the walk (`phases_last_eval`)")
                            if phases_alone is not None else "phases of the timed run (second stream, beside the walk)"},
        "phases_last_eval": phases,
        "phases_alone": phases_alone,
        "knn_retries": tim.get("knn_retries"),
        "comm_ms_last_eval": tim.get("comm_ms"),
        "cpu_baseline": cpu,
    }
    if parity is not None:
        line["parity"] = parity
    if per_rank is not None:
        line["phases_per_rank"] = per_rank
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: N fixed (the metric's definition); weak: N = n x world (n per GPU, default config N / 8)")
    ap.add_argument("--particles", "--n", dest="n", type=int, default=0,
                    help="override the particle count (per GPU with --scaling weak)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]           # NCCL would print its version banner on stdout, in front of the JSON line
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
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--config", args.config,
               "--scaling", args.scaling, "--particles", str(args.n)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py -- Fock (J+K) build ms/SCF-iter and ERI shell quartets/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c18|fe4s4|bo3h3|h2o|h2o64] [--impl reference]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, NCCL)

A "step" is one J/K build of the workload molecule from a seeded synthetic density (SURVEY 8d stress density,
symmetric, not positive semidefinite) -- the call the reference makes once per SCF iteration
(Int4C2E::ContractInts, src/Integral/Int4C2E.cpp:673-683).  Default workload: examples/c18.inp (cyclo[18]carbon,
cc-pVTZ, RHF: J + K), the largest single-GPU configuration of BASELINE.json.
`value`   : canonical shell quartets / s, whole job, densities resident in HBM (device API).
`e2e`     : same metric through the reference-facing host call (host numpy matrices in, host J/K out).
`roofline`: FP64 FMA roofline of the ERI+digestion kernels; F_alg per SURVEY 8d; peak measured in-run by a
            register-resident DFMA loop (MEASURED_PEAKS.json carries no FP64 figure).
`--impl reference`: the reference's CPU path restated (oracle/): stored-integral Gunified stream where the stored
            list fits (h2o, bo3h3), otherwise a bounded sample of a direct CPU build.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {  # name -> (fixture, kind)
    "h2o": ("h2o", "rhf"), "bo3h3": ("bo3h3", "rks_exx0.2"), "c18": ("c18", "rhf"), "fe4s4": ("fe4s4", "uhf"), "h2o64": ("h2o64", "rhf"),
}
DEFAULT_THRESHOLD = {"h2o64": 1e-13}
CLASS_NAMES = ["ss", "ps", "pp", "ds", "dp", "dd", "fs", "fp", "fd", "ff"]


def densities(nbf, kind):
    import scf_harness as H
    if kind == "uhf":
        return None, H.random_symmetric_density(nbf, 1), H.random_symmetric_density(nbf, 2)
    return H.random_symmetric_density(nbf, 0), None, None


def exx_of(kind):
    return 0.2 if kind.startswith("rks") else 1.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(fb, kind, seconds_target=15.0):
    """Bounded CPU baseline on the host cores with the oracle (kind 'port': our restatement, not libint2)."""
    from oracle_lib import Oracle
    o = Oracle()
    o.set_threads(0)      # all online cores, whatever OMP_NUM_THREADS the launcher exported (torchrun: 1)
    Dd, Da, Db = densities(fb.nbf, kind)
    npairs = fb.nshell * (fb.nshell + 1) // 2
    # calibrate on a thin sample, then size the stride for ~seconds_target
    stride = max(1, npairs // 16)
    t = time.perf_counter(); *_, c = o.direct_jk(fb, Dd, Da, Db, exx=exx_of(kind), stride=stride, offset=stride // 2); dt = time.perf_counter() - t
    rate = c[0] / max(dt, 1e-6)
    total_q = npairs * (npairs + 1) // 2
    want = rate * seconds_target
    stride2 = max(1, int(round(total_q / max(want, 1))))
    for _ in range(3):      # the thin calibration sample over-estimates the rate (cheap pairs first): refine until ~seconds_target
        if stride2 >= stride:
            break
        t = time.perf_counter(); *_, c = o.direct_jk(fb, Dd, Da, Db, exx=exx_of(kind), stride=stride2, offset=stride2 // 2); dt = time.perf_counter() - t
        stride = stride2
        if dt >= 0.6 * seconds_target or stride == 1:
            break
        stride2 = max(1, int(stride * dt / seconds_target))
    return dict(value=c[0] / dt, unit="quartets/s", cores=o.nthreads, kind="port", quartets=int(c[0]), seconds=dt,
                sample="direct CPU build (oracle MD ERIs + digestion, OpenMP, all host threads) over every %d-th bra shell pair "
                       "(%d of %d canonical quartets)" % (stride, c[0], total_q))


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, restated (it cannot be built here:
    libint2 / Eigen / libmwfn absent).  Rank 0 only.  Every step is a BOUNDED sample whose real duration is what
    ms_per_step reports (nothing extrapolated in the contract keys); the whole-build extrapolation is a separate key.
      B1 (stored list fits host RAM: h2o, bo3h3, fe4s4): the literal Gunified stream over the stored unique integrals,
         Int4C2E.cpp:601-671 -- the reference's real per-iteration cost; one step = one whole contraction.
      B2 (c18: 159 GiB, (H2O)64: 104 TiB do not fit): a direct CPU build over every n-th bra shell pair."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from chinium_b200.inputs import load_fixture_molecule
    from oracle_lib import Oracle
    fixture, kind = WORKLOADS[args.workload]
    mol, fb = load_fixture_molecule(fixture)
    o = Oracle()
    cores = o.set_threads(0)
    npairs = fb.nshell * (fb.nshell + 1) // 2
    total_q = npairs * (npairs + 1) // 2
    Dd, Da, Db = densities(fb.nbf, kind)
    nint_est = (fb.nbf * (fb.nbf + 1) // 2) ** 2 / 2 * 16 / 2 ** 30
    steps = max(1, args.steps)
    extra = {}
    if nint_est < 6.0 and not args.direct_cpu:   # stored list fits: the reference's real per-iteration path (B1)
        t = time.perf_counter()
        h = o.store_build(fb)
        build_s = time.perf_counter() - t
        for _ in range(args.warmup):
            o.store_contract(h, fb.nbf, Dd, Da, Db, exx_of(kind))
        ts = []
        for _ in range(steps):
            t = time.perf_counter(); o.store_contract(h, fb.nbf, Dd, Da, Db, exx_of(kind)); ts.append(time.perf_counter() - t)
        dt = float(np.median(ts))
        nint = o.store_len(h)
        o.store_free(h)
        value = total_q / dt
        cb = dict(value=value, unit="quartets/s", cores=cores, kind="port",
                  sample="whole workload per step: Gunified stream over the stored list of %d unique integrals (%.2f GiB; the "
                         "reference's per-iteration path B1); one-off integral generation by the oracle took %.1f s and is not "
                         "part of a step, as in the reference" % (nint, nint * 16 / 2 ** 30, build_s))
        extra = {"stored_integrals": int(nint), "integrals_per_s": nint / dt, "stream_GBps": nint * 16 / dt / 1e9}
    else:
        per = max(3.0, 60.0 / max(1, steps + args.warmup))
        cbs = [cpu_sample(fb, kind, per) for _ in range(args.warmup + steps)][args.warmup:]
        q = sum(c["quartets"] for c in cbs); sec = sum(c["seconds"] for c in cbs)
        dt = sec / len(cbs)                      # what one step (= one sample) really took
        value = q / sec
        cb = dict(value=value, unit="quartets/s", cores=cbs[0]["cores"], kind="port",
                  sample=cbs[0]["sample"] + "; the stored-integral list of the reference would need %.0f GiB" % nint_est)
        extra = {"extrapolated_build_ms": total_q / value * 1e3, "quartets_per_step": q / len(cbs)}
    line = {"metric": "ERI shell quartets/s (Fock J+K build)", "value": value, "unit": "quartets/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s, nbf %d, %d canonical shell quartets)" % (args.workload, kind, fb.nbf, total_q)},
            "cpu_baseline": cb, "e2e": {"value": value, "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line))
    return 0


def run_grad(args):
    """ContractGrads(D, D) (Int4C2E.cpp:747-763): ms per call and canonical shell quartets/s, FP64 roofline by the
    gradient F_alg model of DESIGN.md, the oracle's getRepulsion1 restatement timed beside it where it finishes in
    about a minute (h2o, bo3h3).  One GPU; `--impl reference` times only the CPU restatement."""
    from chinium_b200.inputs import load_fixture_molecule
    import scf_harness as H
    fixture, kind = WORKLOADS[args.workload]
    mol, fb = load_fixture_molecule(fixture)
    D = H.random_symmetric_density(fb.nbf, 0) * fb.nbf
    exx = exx_of(kind)
    npairs = fb.nshell * (fb.nshell + 1) // 2
    total_q = npairs * (npairs + 1) // 2
    cpu = None
    if (args.impl == "reference" or not args.no_cpu_baseline) and fb.nbf <= 80:
        from oracle_lib import Oracle
        o = Oracle()
        t = time.perf_counter(); o.contract_grads(fb, D, D, exx); dt = time.perf_counter() - t
        cpu = dict(value=total_q / dt, unit="quartets/s", cores=o.nthreads, kind="port", seconds=dt,
                   sample="whole workload: oracle restatement of getRepulsion1 + ContractGrads (MD derivative ERIs, OpenMP)")
    base = {"metric": "ERI shell quartets/s (nuclear-gradient contraction ContractGrads)", "unit": "quartets/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s, nbf %d, %d canonical shell quartets)" % (args.workload, kind, fb.nbf, total_q),
                       "densities": "D1 = D2 = seeded random symmetric, EXX=%.1f" % exx}}
    if args.impl == "reference":
        if cpu is None:
            print(json.dumps({"impl": "reference", "unavailable": "CPU gradient restatement is only timed for nbf <= 80"}))
            return 0
        base.update(impl="reference", value=cpu["value"], ms_per_step=cpu["seconds"] * 1e3, cpu_baseline=cpu,
                    e2e={"value": cpu["value"], "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(base))
        return 0
    import torch
    from chinium_b200 import Int4C2E
    from chinium_b200.fock import measure_fp64_peak
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this engine has no CPU fallback"}))
        return 1
    eng = Int4C2E(fb, exx, args.threshold if args.threshold is not None else DEFAULT_THRESHOLD.get(args.workload, -1.0))
    for _ in range(max(1, args.warmup)):
        eng.ContractGrads(D, D, 0)
    sampler = ClockSampler(0); sampler.start()
    t = time.perf_counter(); dev_ms = []; launches = 0
    for _ in range(args.steps):
        eng.ContractGrads(D, D, 0)              # host matrices in, host vector out, synchronous
        st = eng.stats
        dev_ms.append(st["ms_grad_last"]); launches += st["n_launches_last"]
    e2e_s = (time.perf_counter() - t) / args.steps
    clocks = sampler.stop()
    peak = measure_fp64_peak(0)
    flops = eng.stats["flops_alg_grad"]
    q = eng.stats["canonical_quartets"]
    base["config"]["workload"] = "%s (%s, nbf %d, %d canonical shell quartets evaluated of %d)" % (args.workload, kind, fb.nbf, q, total_q)
    ms = float(np.median(dev_ms))
    base.update(value=q / (ms * 1e-3), ms_per_step=ms, gpu_launches=launches, clocks=clocks,
                e2e={"value": q / e2e_s, "unit": "quartets/s", "ms_per_step": e2e_s * 1e3,
                     "h2d_bytes_per_step": 2 * fb.nbf * fb.nbf * 8, "d2h_bytes_per_step": 3 * 8 * (int(np.max(fb.shell2atom)) + 1)},
                roofline={"bound": "fp64", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                          "frac": flops / (ms * 1e-3) / 1e12 / peak, "traffic": None, "flops_alg": flops,
                          "kernel": "eri_grad_generic<*> (all class-pair gradient launches of one call)",
                          "peak_source": "measured in-run: register-resident DFMA loop (cf_measure_fp64_peak)"})
    if cpu is not None:
        base["cpu_baseline"] = cpu
    print(json.dumps(base))
    eng.close()
    return 0


def run_multi(args):
    """ContractInts(std::vector<EigenMatrix>&) (Int4C2E.cpp:685-745) with three densities: one batched pass over the
    integrals against three separate J/K builds (what the call cost before batching).  Host API, one GPU."""
    import torch
    from chinium_b200 import Int4C2E
    from chinium_b200.inputs import load_fixture_molecule
    import scf_harness as H
    fixture, kind = WORKLOADS[args.workload]
    mol, fb = load_fixture_molecule(fixture)
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this engine has no CPU fallback"}))
        return 1
    Ds = [H.random_symmetric_density(fb.nbf, s) for s in range(3)]
    eng = Int4C2E(fb, 1.0, args.threshold if args.threshold is not None else DEFAULT_THRESHOLD.get(args.workload, -1.0))
    for _ in range(max(1, args.warmup)):
        eng.ContractInts(Ds, 1, 0)
    t = time.perf_counter(); dev = []
    for _ in range(args.steps):
        eng.ContractInts(Ds, 1, 0); dev.append(eng.stats["ms_eri_last"])
    batched = (time.perf_counter() - t) / args.steps
    eng.ContractInts(Ds[0], None, None, 1, 0)
    t = time.perf_counter(); dev1 = []
    for _ in range(args.steps):
        for D in Ds:
            eng.ContractInts(D, None, None, 1, 0); dev1.append(eng.stats["ms_eri_last"])
    single = (time.perf_counter() - t) / args.steps
    q = eng.stats["canonical_quartets"]
    print(json.dumps({"metric": "multi-density build ContractInts(vector) with 3 matrices", "unit": "ms", "n_gpus": 1, "steps": args.steps,
                      "warmup": args.warmup, "higher_is_better": False, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "%s (nbf %d, %d canonical shell quartets), 3 seeded symmetric densities, EXX=1" % (args.workload, fb.nbf, q)},
                      "value": batched * 1e3, "ms_eri_kernels_batched": float(np.mean(dev)),
                      "three_separate_builds_ms": single * 1e3, "ms_eri_kernels_three_builds": float(np.sum(dev1) / args.steps),
                      "speedup_vs_separate": single / batched}))
    eng.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c18", choices=sorted(WORKLOADS))
    ap.add_argument("--threshold", type=float, default=None,
                    help="Cauchy-Schwarz threshold (Int4C2E's `threshold`); default: -1 (the reference's value, no screening) "
                         "except h2o64, where the unscreened job is 2.7e11 quartets: 1e-13 (see DESIGN.md)")
    ap.add_argument("--per-class", action="store_true", help="also time every class-pair kernel alone (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--direct-cpu", action="store_true", help="reference arm: time the direct CPU build (B2) even where the stored list fits")
    ap.add_argument("--density", default="stress", choices=["stress", "core"],
                    help="stress: seeded random symmetric U(-1,1)/nbf (SURVEY 8d density 3, default); core: the core-Hamiltonian "
                         "projector (density 1, O(1) entries; RHF workloads)")
    ap.add_argument("--path", default="jk", choices=["jk", "grad", "multi"],
                    help="jk: the Fock J/K build (the BASELINE metric, default); grad: the nuclear-gradient contraction "
                         "ContractGrads(D, D) (SURVEY 8f rank 2), its own JSON line")
    args = ap.parse_args()
    if args.path == "grad":
        return run_grad(args)
    if args.path == "multi":
        return run_multi(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from chinium_b200.inputs import load_fixture_molecule
    from chinium_b200.distributed import DistributedInt4C2E
    from chinium_b200.fock import measure_fp64_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this engine has no CPU fallback"}))
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    fixture, kind = WORKLOADS[args.workload]
    mol, fb = load_fixture_molecule(fixture)
    t0 = time.perf_counter()
    thr = args.threshold if args.threshold is not None else DEFAULT_THRESHOLD.get(args.workload, -1.0)
    eng = DistributedInt4C2E(fb, exx_of(kind), thr)
    setup_s = time.perf_counter() - t0
    Dd, Da, Db = densities(fb.nbf, kind)
    dens_note = "seeded random symmetric U(-1,1)/nbf (SURVEY 8d stress density)"
    if args.density == "core":
        if kind == "uhf":
            raise SystemExit("--density core is implemented for the RHF workloads")
        from oracle_lib import Oracle
        import scf_harness as H
        S, T, V = Oracle().one_electron(fb, mol.Z, mol.xyz_bohr)     # input preparation only (host, outside every timed region)
        Dd = H.core_density(S, T + V, mol.nelec // 2)
        dens_note = "core-Hamiltonian projector (SURVEY 8d density 1, O(1) entries)"
    present = [D is not None for D in (Dd, Da, Db)]
    for k, D in enumerate((Dd, Da, Db)):
        if D is not None:
            eng._D[k].copy_(torch.from_numpy(np.ascontiguousarray(D.T)))
    st0 = eng.eng.stats
    nk = sum(present)
    total_q = st0["canonical_quartets"]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)   # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.build_device(present)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the per-step events)
        evs[i][0].record()
        eng.build_device(present)
        evs[i][1].record()
        launches += eng.eng.stats["n_launches_last"] + (1 if world > 1 else 0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps
    # ERI-kernel time of a build (CUDA events inside the library, first to last class-pair kernel): median over `steps`
    # further builds, each synchronised so that its events can be read (outside the timed region above)
    eri_ms = []
    for i in range(args.steps):
        flush.zero_()
        eng.build_device(present)
        torch.cuda.synchronize()
        st = eng.eng.sync_stats()
        eri_ms.append(st["ms_eri_last"])
    eri_med = float(np.median(eri_ms))
    em = torch.tensor([eri_med], dtype=torch.float64, device=dev)
    ex = torch.tensor([float(st["flops_executed_last"]), float(st["primitive_quartets_executed_last"]), float(st["quartets_evaluated_last"])],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(em, op=dist.ReduceOp.MAX)
        dist.all_reduce(ex, op=dist.ReduceOp.SUM)
    eri_med = float(em.item())
    flops_exec, prim_exec, q_eval = (float(x) for x in ex.tolist())

    # ---- end to end through the reference-facing HOST call: host matrices (pinned) in, host J/K (pinned) out -----------
    # N = 1: cf_build_jk, the C entry point the C++ adaptor's ContractInts calls (H2D, all kernels, D2H inside the call);
    # N > 1: the same through the per-rank partition + NCCL int64 all-reduce (chinium_b200/distributed.py)
    n = fb.nbf
    pin = lambda: torch.empty((n, n), dtype=torch.float64).pin_memory()
    hin = [pin() if p else None for p in present]
    for k, D in enumerate((Dd, Da, Db)):
        if D is not None:
            hin[k].numpy()[...] = D
    hout = [pin() for _ in range(4)]
    fview = lambda t: None if t is None else t.numpy().T        # F-ordered view of the pinned block (matrices are symmetric)
    args_in = [fview(t) for t in hin]
    outs = [fview(t) for t in hout]
    if world == 1:
        call = lambda: eng.eng._contract(args_in[0], args_in[1], args_in[2], out=outs)
    else:
        call = lambda: eng.ContractInts(args_in[0], args_in[1], args_in[2], 1, 0, out=outs)
    for _ in range(2):
        call()
    barrier()
    t = time.perf_counter()
    for _ in range(args.steps):
        out = call()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([(time.perf_counter() - t) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    n2 = fb.nbf * fb.nbf * 8

    if rank == 0:
        peak = measure_fp64_peak(local_rank)
        traffic, traffic_note = None, None
        try:   # DRAM bytes of one build from the committed ncu pass (profiles/ncu_traffic.json), not measured in this run
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
            if tj and world == 1:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_note = tj["source"] + "; ncu flushes L2 before every kernel, so this is an upper bound of the in-situ traffic"
        except Exception:
            pass
        flops = st0["flops_alg_jk"][nk] * world      # whole job (stats are per partition)
        tf_nom = flops / (eri_med * 1e-3) / 1e12 / world
        tf_exec = flops_exec / (eri_med * 1e-3) / 1e12 / world
        line = {
            "metric": "ERI shell quartets/s (Fock J+K build)", "value": total_q / (ms_per_step * 1e-3), "unit": "quartets/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s, nbf %d, %d canonical shell quartets, reference RepulsionLength %d)" % (
                           args.workload, kind, fb.nbf, total_q, st0["ref_repulsion_length"]),
                       "densities": "%s, nK=%d, EXX=%.1f" % (dens_note, nk, exx_of(kind)),
                       "schwarz_threshold": thr,
                       "l2": "flushed between timed iterations (256 MiB write, outside the step events)",
                       "partition": "static chunk-interleaved split of every class-pair quartet range over %d rank(s); "
                                    "int64 all-reduce" % world,
                       "setup_s": setup_s},
            "fock_build_ms": ms_per_step,
            "e2e": {"value": total_q / e2e_s, "unit": "quartets/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": n2 * sum(present), "d2h_bytes_per_step": n2 * (1 + nk),
                    "call": "cf_build_jk (C ABI host call, pinned host buffers)" if world == 1 else "DistributedInt4C2E.ContractInts (pinned host buffers in and out)"},
            "gpu_launches": launches,
            "clocks": clocks,
            # `frac` / `achieved`: model flops of the work the kernels really EXECUTED (device counters of primitive quartets
            # that passed both primitive cutoffs, x the per-primitive flops of SURVEY 8d, + digestion of the evaluated
            # quartets).  `frac_nominal` divides the same time into F_alg at NOMINAL contraction depths (every primitive
            # quartet of every surviving shell pair, the SURVEY 8d definition) -- work that is skipped counts there.
            "roofline": {"bound": "fp64", "achieved": tf_exec, "peak": peak, "unit": "TFLOP/s", "frac": tf_exec / peak,
                         "frac_executed": tf_exec / peak, "frac_nominal": tf_nom / peak, "achieved_nominal": tf_nom,
                         "traffic": traffic, "traffic_source": traffic_note,
                         # D in + J/K out + 6 doubles per primitive pair (P pairs <-> P(P+1)/2 primitive quartets)
                         "algorithmic_bytes": 2 * (1 + nk) * n2 + 48 * (2.0 * st0["primitive_quartets"]) ** 0.5,
                         "kernel": "eri_jk_tpq/tpqa/tpqs/wg<*> (all class-pair ERI+digestion launches of one build, per GPU)",
                         "ms": eri_med, "ms_all_steps": eri_ms,
                         "flops_executed": flops_exec, "primitive_quartets_executed": prim_exec, "quartets_evaluated": q_eval,
                         "primitive_quartets_nominal_kept_pairs": st0["primitive_quartets"],
                         "flops_alg_nominal": flops, "j_two_limb": st["j_two_limb_last"],
                         "peak_source": "measured in-run: register-resident DFMA loop (cf_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 entry"},
        }
        if args.per_class and world == 1:
            rows = eng.eng.profile_tasks(eng._D[0].data_ptr() if present[0] else None, eng._D[1].data_ptr() if present[1] else None,
                                         eng._D[2].data_ptr() if present[2] else None)
            for r in rows:
                r["tflops"] = r["flops_executed"] / (r["ms"] * 1e-3) / 1e12
                r["frac"] = r["tflops"] / peak
                r["frac_nominal"] = r["flops_alg"] / (r["ms"] * 1e-3) / 1e12 / peak
            line["per_class"] = rows
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(fb, kind, 15.0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

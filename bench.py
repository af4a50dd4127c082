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
    libint2 / Eigen / libmwfn absent).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from chinium_b200.inputs import load_fixture_molecule
    from oracle_lib import Oracle
    fixture, kind = WORKLOADS[args.workload]
    mol, fb = load_fixture_molecule(fixture)
    o = Oracle()
    npairs = fb.nshell * (fb.nshell + 1) // 2
    total_q = npairs * (npairs + 1) // 2
    Dd, Da, Db = densities(fb.nbf, kind)
    nint_est = (fb.nbf * (fb.nbf + 1) // 2) ** 2 / 2 * 16 / 2 ** 30
    if nint_est < 2.0:   # stored list fits: the reference's real per-iteration path (B1)
        h = o.store_build(fb)
        for _ in range(args.warmup):
            o.store_contract(h, fb.nbf, Dd, Da, Db, exx_of(kind))
        t = time.perf_counter()
        for _ in range(max(1, args.steps)):
            o.store_contract(h, fb.nbf, Dd, Da, Db, exx_of(kind))
        dt = (time.perf_counter() - t) / max(1, args.steps)
        nint = o.store_len(h)
        o.store_free(h)
        cb = dict(value=total_q / dt, unit="quartets/s", cores=o.nthreads, kind="port",
                  sample="whole workload: Gunified stream over the stored list of %d unique integrals (reference's per-iteration path)" % nint)
    else:
        per = max(3.0, 40.0 / max(1, args.steps + args.warmup))
        cbs = [cpu_sample(fb, kind, per) for _ in range(args.warmup + max(1, args.steps))][args.warmup:]
        q = sum(c["quartets"] for c in cbs); s = sum(c["seconds"] for c in cbs)
        dt = total_q / (q / s)
        cb = dict(value=q / s, unit="quartets/s", cores=cbs[0]["cores"], kind="port",
                  sample=cbs[0]["sample"] + "; the stored-integral list of the reference would need %.0f GiB" % nint_est)
    line = {"metric": "ERI shell quartets/s (Fock J+K build)", "value": cb["value"], "unit": "quartets/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s, nbf %d, %d canonical shell quartets)" % (args.workload, kind, fb.nbf, total_q)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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
    ms = float(np.mean(dev_ms))
    flops = eng.stats["flops_alg_grad"]
    q = eng.stats["canonical_quartets"]
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
    eri_ms = []
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
    st = eng.eng.sync_stats()
    eri_ms_last = st["ms_eri_last"]
    tot = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    ms_per_step = float(tot.item()) / args.steps

    # ---- end to end through the host API (host matrices in, host J/K out, pinned staging inside) ---------
    for _ in range(2):
        eng.ContractInts(Dd, Da, Db, 1, 0)
    barrier()
    t = time.perf_counter()
    for _ in range(args.steps):
        out = eng.ContractInts(Dd, Da, Db, 1, 0)
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
        line = {
            "metric": "ERI shell quartets/s (Fock J+K build)", "value": total_q / (ms_per_step * 1e-3), "unit": "quartets/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s (%s, nbf %d, %d canonical shell quartets, %d unique integrals)" % (
                           args.workload, kind, fb.nbf, total_q, st0["unique_integrals"]),
                       "densities": "seeded random symmetric (SURVEY 8d stress density), nK=%d, EXX=%.1f" % (nk, exx_of(kind)),
                       "schwarz_threshold": thr,
                       "l2": "flushed between timed iterations (256 MiB write, outside the step events)",
                       "partition": "static chunk-interleaved split of every class-pair quartet range over %d rank(s); "
                                    "int64 all-reduce" % world,
                       "setup_s": setup_s},
            "fock_build_ms": ms_per_step,
            "e2e": {"value": total_q / e2e_s, "unit": "quartets/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": n2 * sum(present), "d2h_bytes_per_step": n2 * (1 + nk)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": flops / (eri_ms_last * 1e-3) / 1e12 / world, "peak": peak, "unit": "TFLOP/s",
                         "frac": flops / (eri_ms_last * 1e-3) / 1e12 / world / peak, "traffic": traffic, "traffic_source": traffic_note,
                         # D in + J/K out + 6 doubles per primitive pair (P pairs <-> P(P+1)/2 primitive quartets)
                         "algorithmic_bytes": 2 * (1 + nk) * n2 + 48 * (2.0 * st0["primitive_quartets"]) ** 0.5,
                         "kernel": "eri_jk_tpq/tpqs/wg<*> (all class-pair ERI+digestion launches of one build, per GPU)", "ms": eri_ms_last,
                         "flops_alg": flops, "peak_source": "measured in-run: register-resident DFMA loop (cf_measure_fp64_peak); "
                                                            "MEASURED_PEAKS.json has no FP64 entry"},
        }
        if args.per_class and world == 1:
            rows = eng.eng.profile_tasks(eng._D[0].data_ptr() if present[0] else None, eng._D[1].data_ptr() if present[1] else None,
                                         eng._D[2].data_ptr() if present[2] else None)
            for r in rows:
                r["tflops"] = r["flops_alg"] / (r["ms"] * 1e-3) / 1e12
                r["frac"] = r["tflops"] / peak
            line["per_class"] = rows
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(fb, kind, 15.0)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py — the headline benchmark of the clear-sky spectral hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (default, `--workload c4`): BASELINE.json configs[3], the case north_star states its targets on —
10^6 synthetic lines x 10^6 frequencies (1-100 THz) x 100 levels, no cutoff, linsrc Stokes chain — taken one
frequency shard per GPU: every GPU owns 125 000 frequencies.  The N-GPU run covers every (8/N)-th point of the
10^6-point grid, cut into contiguous blocks like the reference's OpenMP frequency chunks (src/m_lbl.cc:273-295),
so N = 8 IS the 10^6 x 10^6 x 100 case and N = 1 is a uniformly strided eighth of it ("weak" scaling: the work per
GPU is fixed).  `--workload c2` keeps round 1's configs[1] line.  One "step" = one pass of the hot path over that
input: line prepare (K1) + line sum (K2/K3) + fused Stokes chain (K4-K6) + the NCCL gather of spectral_rad.
At N > 1 the LEVELS of the line sum are dealt over the ranks (rank r: levels r, r + N, ... for all N x 125 000
frequencies, so the per-(line, level) work is not repeated by every frequency shard), one NCCL all-to-all transposes
K, and every rank runs the Stokes chain on its contiguous frequency block (arts_b200/shard.py LevelExchange, DESIGN.md
section 6; AB200_BENCH_SPLIT=freq keeps the plain frequency split; same bits either way).

Printed (rank 0, one JSON line): `value` = line*freq*level evaluations per second of the whole job with inputs
resident in HBM, timed with CUDA events, max over ranks; `e2e` = the same metric through the reference-facing
C-ABI call with HOST buffers (H2D and D2H inside the timed region; at N > 1 the one-process multi-device call
ab200_multi_clearsky_emission over the N devices, driven by rank 0); `roofline` for the dominant kernel (FP64-pipe
bound line sum: algorithmic FLOPs of SURVEY.md 8(d) over the event-timed kernel duration, against the DFMA peak
measured in this run); `roofline_stokes` (HBM bound); `cpu_baseline` = the CPU oracle (a port of the reference's
path linked with the reference's own Faddeeva.cc) on this box's host cores on a bounded sample; `extra` = the
configs[3]-ii variant (750 GHz ByLine cutoff) on the same shard.

`--impl reference` times that CPU implementation alone, on the same config and metric, with ALL host cores
(the OpenMP team is set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "line*freq*level evals/s"
UNIT = "evals/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rte", default="linsrc", choices=["linsrc", "constant"])
    # workload size knobs (defaults = BASELINE configs[1]); smaller values are for quick checks only
    ap.add_argument("--lines-per-species", type=int, default=20_000)
    ap.add_argument("--nf-per-gpu", type=int, default=0, help="frequencies per GPU (default: 125000 for c4, 100000 for c2)")
    ap.add_argument("--levels", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of one baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c4", choices=["c2", "c4", "c4strong"],
                    help="c4 (default) = BASELINE configs[3] one 125000-frequency shard per GPU (N = 8 is the full 1e6 x 1e6 x 100 "
                         "case); c4strong = the same grid of FIXED total size --c4-nf sharded over the GPUs; c2 = configs[1]")
    ap.add_argument("--cpu-threads", type=int, default=0, help="OpenMP threads of the CPU arm (0 = all cores of the box)")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[3]-ii (750 GHz cutoff) extra measurement")
    ap.add_argument("--c4-lines", type=int, default=1_000_000)
    ap.add_argument("--c4-nf", type=int, default=1_000_000)
    ap.add_argument("--c4-cutoff-ghz", type=float, default=0.0,
                    help="configs[3] variant (ii) of SURVEY 8(d): ByLine cutoff in GHz (0 = none); the metric then counts NOMINAL pairs")
    a = ap.parse_args()
    if a.nf_per_gpu <= 0:
        a.nf_per_gpu = 100_000 if a.workload == "c2" else 125_000
    return a


def workload(args, world, cutoff_ghz=None):
    from arts_b200 import synth

    cutoff_ghz = args.c4_cutoff_ghz if cutoff_ghz is None else cutoff_ghz
    if args.workload in ("c4", "c4strong"):
        cut = cutoff_ghz * 1e9 if cutoff_ghz > 0 else None
        case = synth.case_c4(n_lines=args.c4_lines, nf=args.c4_nf, np_=args.levels, cutoff=cut)
        case.rte_option = args.rte
        cut_txt = "no cutoff" if cut is None else f"ByLine cutoff {cutoff_ghz:g} GHz (value counts nominal line x frequency pairs)"
        if args.workload == "c4strong":
            return case, (f"C4 (BASELINE configs[3]): {args.c4_lines} lines x {args.c4_nf} frequencies (1-100 THz, total, sharded over "
                          f"the GPUs) x {args.levels} levels, {cut_txt}, {args.rte} Stokes chain")
        # one shard of nf_per_gpu frequencies per GPU: the N-GPU run covers every (c4_nf / (N nf_per_gpu))-th grid point
        nf_run = args.nf_per_gpu * world
        if nf_run <= args.c4_nf and args.c4_nf % nf_run == 0:
            stride = args.c4_nf // nf_run
            case.f = np.ascontiguousarray(case.f[::stride])
            case.I_bkg = np.ascontiguousarray(case.I_bkg[::stride])
            grid_txt = (f"every {stride}-th point of the {args.c4_nf}-point 1-100 THz grid" if stride > 1
                        else f"the whole {args.c4_nf}-point 1-100 THz grid")
        else:
            case.f = np.linspace(case.f[0], case.f[-1], nf_run)
            case.I_bkg = np.zeros((nf_run, 4))
            case.I_bkg[:, 0] = synth.planck(case.f, 288.0)
            grid_txt = f"{nf_run} points over 1-100 THz"
        return case, (f"C4 shard (BASELINE configs[3]: 1e6 lines x 1e6 frequencies x 100 levels over 8 GPUs): {args.c4_lines} lines x "
                      f"{args.nf_per_gpu} frequencies per GPU ({grid_txt}) x {args.levels} levels, {cut_txt}, {args.rte} Stokes chain")
    nf = args.nf_per_gpu * world
    case = synth.case_c2(lines_per_species=args.lines_per_species, nf=nf, np_=args.levels, rte_option=args.rte)
    name = (f"C2 (BASELINE configs[1]): 5 species x {args.lines_per_species} Voigt lines, {args.levels}-level nadir path, "
            f"{args.nf_per_gpu} frequencies per GPU over 1-1000 GHz, {args.rte} Stokes chain")
    return case, name


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 7]
        os.unlink(self.f.name)
        sm, reasons, smax, pw = [], set(), None, []
        for r in rows:
            r = [x.strip() for x in r]
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


# --------------------------------------------------------------------------- CPU arm
def oracle_sample(case, idx, rte):
    """Runs the CPU oracle on the frequency subset ``idx`` of ``case``; returns seconds."""
    from tests import oracle_lib as orc

    f = np.ascontiguousarray(case.f[idx])
    bkg = np.ascontiguousarray(case.I_bkg[idx])
    t0 = time.perf_counter()
    orc.clearsky_emission(case.cat, f, case.atm, case.r, bkg, rte_option=rte)
    return time.perf_counter() - t0


CPU_FLAGS = ("oracle/oracle.cpp: g++ -O3 -fopenmp -ffp-contract=off, no -march=native (built in the CPU container, run on the "
             "box); the reference's Faddeeva.cc object: -O2 -ffp-contract=off")


def cpu_threads(args):
    """All cores this process may use — set explicitly, whatever OMP_NUM_THREADS the launcher exported
    (torch.distributed.run sets it to 1 for its workers)."""
    from tests import oracle_lib as orc

    want = args.cpu_threads if args.cpu_threads > 0 else len(os.sched_getaffinity(0))
    got = orc.set_num_threads(want)
    if got != want:
        print(f"bench.py: warning: asked the CPU arm for {want} threads, OpenMP gives {got}", file=sys.stderr)
    return got


def calibrate_sample(case, rte, target_s, nthreads):
    """Strided (grid-representative) frequency sample sized so that one oracle pass takes ~target_s."""
    n0 = max(nthreads, 16)
    idx = np.linspace(0, case.nf - 1, n0).astype(np.int64)
    oracle_sample(case, idx[: max(nthreads // 2, 2)], rte)  # page in
    t = oracle_sample(case, idx, rte)
    n = int(min(case.nf, max(n0, n0 * target_s / max(t, 1e-6))))
    return np.unique(np.linspace(0, case.nf - 1, n).astype(np.int64))


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port + the reference's Faddeeva object) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = cpu_threads(args)
    case, name = workload(args, world)  # the same grid (and workload string) as the GPU arm at this N
    idx = calibrate_sample(case, args.rte, args.cpu_seconds, cores)
    for _ in range(args.warmup):
        oracle_sample(case, idx[: max(len(idx) // 8, 1)], args.rte)
    ts = [oracle_sample(case, idx, args.rte) for _ in range(args.steps)]
    evals = float(case.n_lines) * len(idx) * case.np_
    value = evals * len(ts) / sum(ts)
    sample = f"{len(idx)} of {case.nf} frequencies (uniform stride) x all {case.n_lines} lines x {case.np_} levels per step"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count(),
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS"), "kind": "port", "sample": sample,
                         "build_flags": CPU_FLAGS,
                         "note": "oracle/oracle.cpp (restatement of the reference's lbl + rtepack path) linked with the "
                                 "reference's own 3rdparty/Faddeeva/Faddeeva.cc object; the reference itself needs "
                                 "GCC >= 14 and external data and cannot be built here (DESIGN.md)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage2": {"metric": "freq*level Stokes steps/s", "note": "included in the step; < 0.1% of CPU time"},
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from arts_b200 import roofline, shard, wsm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available() or wsm.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device — arts_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    wsm.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    case, name = workload(args, world)
    mine, off, cnt = shard.shard_case(case, rank, world)
    nl, np_, nf_total = case.n_lines, case.np_, case.nf

    stream = torch.cuda.current_stream()
    cat = wsm.Catalog(case.cat)
    # N > 1: the LEVELS of the line sum are dealt over the ranks (rank r: levels r, r + N, ... for ALL frequencies, so the
    # per-(line, level) work - line records, cluster moments - is not repeated by every frequency shard), one NCCL all-to-all
    # transposes K, and every rank runs the Stokes chain on its contiguous frequency block; AB200_BENCH_SPLIT=freq keeps the
    # plain frequency split.  Same bits either way (tests/test_shard_gloo.py, tests/test_gpu_multi.py).
    level_split = (world > 1 and np_ >= world and cnt * world == nf_total and os.environ.get("AB200_BENCH_SPLIT", "level") != "freq")
    if level_split:
        from arts_b200 import _abi as abi

        ex = shard.LevelExchange(np_, nf_total, rank, world)
        a = case.atm
        pick = lambda x: None if x is None else np.ascontiguousarray(x[ex.mine])  # noqa: E731
        sub = abi.AtmPath(T=pick(a.T), P=pick(a.P), vmr=pick(a.vmr), isorat=pick(a.isorat), Q=pick(a.Q), dQdT=pick(a.dQdT),
                          mag=pick(a.mag), los=pick(a.los), wind=pick(a.wind))
        lpath = wsm.Path(cat, nf_total, len(ex.mine), 0, stream=stream.cuda_stream)
        lpath.upload(case.f, sub, np.zeros(max(len(ex.mine) - 1, 1)), None, rte_option=args.rte)
        path = wsm.Path(cat, cnt, np_, 0, stream=stream.cuda_stream, stage2_only=True)
        path.upload(mine.f, mine.atm, mine.r, mine.I_bkg, rte_option=args.rte)
        K1 = shard.as_torch(lpath.device_ptr(1), (len(ex.mine), lpath.k_pitch, 7), dev)
        K2 = shard.as_torch(path.device_ptr(1), (np_, path.k_pitch, 7), dev)
    else:
        path = wsm.Path(cat, cnt, np_, 0, stream=stream.cuda_stream)
        path.set_grid_bounds(np.tile([case.f[0], case.f[-1]], (np_, 1)))
        path.upload(mine.f, mine.atm, mine.r, mine.I_bkg, rte_option=args.rte)
        lpath = path
    I_local = shard.as_torch(path.device_ptr(0), (cnt, 4), dev)

    dfma_tflops, _ = wsm.measure_dfma_peak(20000)
    dfma_mix_tflops, _ = wsm.measure_dfma_mix(20000)
    hist = lpath.region_histogram(200_000, seed=1)
    fl_eval, region_frac = roofline.flops_per_eval(hist)

    def step():
        lpath.run_propmat()
        if level_split:
            ex.exchange(K1, K2)
            path.adopt_K()
        path.run_stokes()
        if world > 1:
            return shard.gather_spectral_rad(I_local, nf_total)
        return I_local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    wsm.lib().ab200_launch_count(1)
    for _ in range(args.warmup):
        step()
    barrier()
    path.timings()  # drop warm-up records
    lpath.timings()
    path.set_timing(True)
    lpath.set_timing(True)
    wsm.lib().ab200_launch_count(1)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = int(wsm.lib().ab200_launch_count(0))
    kt = dict(lpath.timings())
    if lpath is not path:
        kt["stokes"] = path.timings()["stokes"]
    path.set_timing(False)
    lpath.set_timing(False)
    checksum = float(out[:, 0].sum().item())

    evals_per_step = float(nl) * nf_total * np_
    value = evals_per_step * args.steps / (ms * 1e-3)

    # dominant kernel: the real-only line sum (mode 0); algorithmic FLOPs / event-timed duration
    k_ms, k_n = kt["sum_real"]
    k_ms_per = k_ms / max(k_n, 1)
    flops_per_launch = fl_eval * float(nl) * cnt * np_ / max(k_n / args.steps, 1)
    achieved_tf = flops_per_launch / (k_ms_per * 1e-3) / 1e12 if k_ms_per > 0 else 0.0
    s_ms, s_n = kt["stokes"]
    s_ms_per = s_ms / max(s_n, 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM bytes per launch cannot be counted inside an un-profiled run: they come from the committed `ncu --set full`
    # capture of THIS shape (profiles/traffic.json lists one entry per shape, with the .ncu.txt it was read from), else null
    traffic, traffic_src = {}, None
    try:
        for tj in json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["captures"]:
            if tj["shape"] == {"lines": nl, "nf_per_gpu": cnt, "levels": np_}:
                traffic, traffic_src = tj, tj.get("source")
    except (OSError, KeyError, ValueError, TypeError):
        pass
    st_bytes = roofline.stokes_bytes_per_step(np_) * cnt * np_
    st_gbs = st_bytes / (s_ms_per * 1e-3) / 1e9 if s_ms_per > 0 else 0.0

    # e2e: the reference-facing C-ABI call with host (pinned) buffers, H2D + D2H inside the timed region.  N = 1:
    # ab200_clearsky_emission.  N > 1: the call a shim inside the reference's ONE process makes, ab200_multi_clearsky_emission
    # over the N devices for the whole grid (rank 0 drives it; the other ranks wait on a CPU barrier and leave their GPUs idle).
    n_e2e = max(2, min(args.steps, 3))
    if world == 1:
        f_h = torch.from_numpy(mine.f).pin_memory().numpy()
        b_h = torch.from_numpy(mine.I_bkg).pin_memory().numpy()
        wsm.set_thread_stream(stream.cuda_stream)
        wsm.spectral_radClearskyEmission(cat, f_h, mine.atm, mine.r, b_h, rte_option=args.rte)  # builds the thread workspace
        barrier()
        e0.record()
        for _ in range(n_e2e):
            I_host, _ = wsm.spectral_radClearskyEmission(cat, f_h, mine.atm, mine.r, b_h, rte_option=args.rte)
        e1.record()
        barrier()
        ms_e2e = float(e0.elapsed_time(e1))
        e2e_call = "ab200_clearsky_emission (host buffers, pinned)"
        e2e_matches = bool(np.array_equal(I_host, I_local.cpu().numpy()))
    else:
        cpu_group = dist.new_group(backend="gloo")
        ms_e2e, e2e_matches, f_h, b_h, I_host = 0.0, None, case.f, case.I_bkg, np.empty((nf_total, 4))
        barrier()
        if rank == 0:
            f_h = torch.from_numpy(case.f).pin_memory().numpy()
            b_h = torch.from_numpy(case.I_bkg).pin_memory().numpy()
            multi = wsm.MultiDevice(case.cat, n_devices=world)
            wsm.spectral_radClearskyEmission(multi, f_h, case.atm, case.r, b_h, rte_option=args.rte)  # builds the workspaces
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n_e2e):
                I_host, _ = wsm.spectral_radClearskyEmission(multi, f_h, case.atm, case.r, b_h, rte_option=args.rte)
            e1.record()
            torch.cuda.synchronize()
            ms_e2e = float(e0.elapsed_time(e1))
            multi.close()
            e2e_matches = bool(np.array_equal(I_host, out.cpu().numpy()))
        dist.barrier(group=cpu_group)
        e2e_call = f"ab200_multi_clearsky_emission (one host process, {world} devices, host buffers, pinned)"
    e2e_value = evals_per_step * n_e2e / (ms_e2e * 1e-3) if ms_e2e > 0 else None
    h2d = int(f_h.nbytes + b_h.nbytes + 8 * np_ * (3 + 3 * case.cat.n_species + 28 + 3))
    d2h = int(I_host.nbytes)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload == "c4strong" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "lines": nl, "levels": np_, "nf_total": nf_total, "nf_per_gpu": cnt,
                   "sharding": ("levels of the line sum dealt over the ranks (rank r: levels r, r + N, ... for all frequencies), one NCCL "
                                "all-to-all of K, Stokes chain on contiguous frequency blocks, one NCCL all-gather of spectral_rad"
                                if level_split else
                                "contiguous frequency blocks (matpack::omp_offset_count), catalog replicated, one NCCL all-gather of "
                                "spectral_rad") if world > 1 else "single GPU",
                   "l2": "inputs exceed L2: per step the kernels write and re-read %.2f GB of line records, %.2f GB of cluster moments "
                         "and %.2f GB of K (L2 = 126 MB), no flush needed" % (cat.host.n_lines * 128.0 * np_ / 1e9,
                                                                               cat.host.n_lines / 256.0 * 21.3 * 160.0 * np_ / 1e9,
                                                                               cnt * np_ * 56.0 / 1e9)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "call": e2e_call, "equals_resident_result": e2e_matches},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "kernel": "real line sum: lbl_fmm_{moments,far,near}_kernel (lbl_sum_real_kernel for segments with ByLine cutoffs)",
                     "achieved": achieved_tf, "peak": dfma_tflops,
                     "unit": "TFLOP/s", "frac": achieved_tf / dfma_tflops if dfma_tflops else None,
                     "traffic": traffic.get("real_line_sum_fmm_kernels_per_step") or traffic.get("lbl_sum_real_kernel"),
                     "traffic_source": traffic_src, "launches_per_step": k_n / args.steps,
                     "peak_source": "DFMA loop measured in this run (ab200_measure_dfma_peak); MEASURED_PEAKS.json has no FP64 figure",
                     "algorithmic_flop_per_eval": fl_eval, "regions": region_frac, "kernel_ms": k_ms_per,
                     "executed": {"peak_with_rcp_mix_tflops": dfma_mix_tflops,
                                  "note": "achieved counts the ALGORITHMIC flops of SURVEY.md 8(d) (28 per far-wing evaluation of the reference's "
                                          "closed form) for every nominal line x frequency x level pair.  frac >> 1 because cutoff-free real "
                                          "segments are summed as a hierarchical far-field (multipole) expansion of the four-term continued "
                                          "fraction (arts_b200/csrc/lbl_fmm.cu): per (frequency, level) a few hundred cluster expansions of 24 "
                                          "FP64 instructions plus the ~1e2-4e2 pairs within 48 Doppler widths, instead of 1e6 pairs - same "
                                          "spectra to 1e-9, bit-identical under frequency partitions.  FP64-pipe utilisation of the kernels "
                                          "actually run is in profiles/r3c_fmm_kernels.ncu.txt (ncu): far pass 67 %, near pass 61 %, moments 37 %, prepare 19 %."},
                     "kernel_share_of_step": k_ms / ms if ms else None},
        "roofline_stokes": {"bound": "hbm", "kernel": "stokes_chain_kernel", "achieved": st_gbs, "peak": hbm_peak,
                            "unit": "GB/s", "frac": st_gbs / hbm_peak, "traffic": traffic.get("stokes_chain_kernel"),
                            "algorithmic_bytes": st_bytes, "kernel_ms": s_ms_per,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
        "stage2": {"metric": "freq*level Stokes steps/s", "value": float(cnt) * np_ * world / (s_ms_per * 1e-3) if s_ms_per else None,
                   "unit": "steps/s"},
        "kernel_ms": {k: v[0] / max(v[1], 1) for k, v in kt.items()},
        "checksum_I": checksum,
    }

    if args.workload == "c4" and args.c4_cutoff_ghz == 0 and not args.no_extra:
        # configs[3]-ii of SURVEY 8(d): the same shard with a 750 GHz ByLine cutoff (HITRAN practice); nominal pairs counted
        # (plain frequency split at N > 1)
        if lpath is not path:
            lpath.close()
        path.close()
        cat.close()
        case2, name2 = workload(args, world, cutoff_ghz=750.0)
        mine2, _, cnt2 = shard.shard_case(case2, rank, world)
        cat = wsm.Catalog(case2.cat)
        path = wsm.Path(cat, cnt2, np_, 0, stream=stream.cuda_stream)
        path.set_grid_bounds(np.tile([case2.f[0], case2.f[-1]], (np_, 1)))
        path.upload(mine2.f, mine2.atm, mine2.r, mine2.I_bkg, rte_option=args.rte)
        lpath = path
        for _ in range(2):
            path.run_propmat(); path.run_stokes()
        barrier()
        n2 = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n2):
            path.run_propmat(); path.run_stokes()
        e1.record()
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2 = float(ms2.item())
        line["extra"] = {"c4_cutoff750": {"workload": name2, "ms_per_step": ms2 / n2, "steps": n2,
                                          "value": evals_per_step * n2 / (ms2 * 1e-3), "unit": UNIT + " (nominal pairs)"}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads(args)
        idx = calibrate_sample(case, args.rte, args.cpu_seconds, cores)
        t = oracle_sample(case, idx, args.rte)
        line["cpu_baseline"] = {
            "value": float(nl) * len(idx) * np_ / t, "unit": UNIT, "cores": cores, "host_cores": os.cpu_count(), "kind": "port",
            "build_flags": CPU_FLAGS,
            "sample": f"{len(idx)} of {nf_total} frequencies (uniform stride) x all {nl} lines x {np_} levels, {t:.1f} s"}
    if rank == 0:
        emit(line)
    if lpath is not path:
        lpath.close()
    path.close()
    cat.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was
    routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # native libraries (e.g. "NCCL version ..." banners) write to fd 1: keep stdout to the JSON line
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: W < 3 warm-up steps — number not valid for reporting", file=sys.stderr)
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- particle-steps/s of the FLEXPART per-particle hot path
(fpb_conccalc + fpb_step) on B200, with the HBM-gather roofline and the CPU
baseline beside it.

Workload (BASELINE.json configs[1], "C2"): 1 M particles per GPU, global
0.5 deg x 138-level synthetic ECMWF-shaped met, Hanna turbulence (CTL=5,
IFINE=4 -> method 1), 100 box releases in the lowest 2 km, 900 s
synchronisation interval, concentration sampling every step.  A "step" is one
pass of the hot path over the whole batch: fpb_conccalc(itime, 1.0) +
fpb_step(itime); the output-interval exchange (NCCL reduce of gridunc to rank
0, then zeroing) runs every 4th step as in the reference (LOUTSTEP=3600).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload c2|c5slice] [--particles P]

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every
field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# algorithmic bytes per particle-step (SURVEY.md 8d / DESIGN.md): state
# 74 B read + 58 B written, met gather 161 floats in the PBL (method 0 count)
# or 105 floats above it; +76 B per conccalc sample.
B_PBL, B_FT, B_CONC = 776, 552, 76


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region: NVML polled every 5 ms from a
    thread (the same counters nvidia-smi prints); nvidia-smi -lms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.mx, self.seen = [], [], set()
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = N.nvmlDeviceGetHandleByIndex(idx)
            masks = {"hw_slowdown": N.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": N.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": N.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": N.nvmlClocksThrottleReasonSwPowerCap}
            self.mx.append(float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)))

            def poll():
                while not self._stop.is_set():
                    self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                    r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for n, m in masks.items():
                        if r & m:
                            self.seen.add(n)
                    self._stop.wait(0.005)
            self.nvml = threading.Thread(target=poll, daemon=True)
            self.nvml.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml:
            self._stop.set()
            self.nvml.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(self.sm), "source": "nvml, 5 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = [n for k, n in enumerate(self.NAMES) if any(len(r) >= 6 and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 20"}


def build_workload(args, rank, world, device, cpu_only=False, n_particles=None):
    import flexpart_b200 as fb
    import cases
    npart_total = n_particles or args.particles
    nrel = 100
    extra = {}
    if args.workload == "c5slice":
        kw = dict(ctl=-5.0)            # method 0: one Langevin step per sync (HBM-bound regime)
        zmax, lat = 12000.0, (-85.0, 85.0)
    elif args.workload == "c3":
        # BASELINE configs[2]: CBL skewed turbulence (forces ctl >= 5, ifine*ctl >= 50), two species with
        # dry deposition, one of them wet-scavenged, nested output grid
        kw = dict(ctl=10.0, cblflag=1)
        if os.environ.get("FPB_BENCH_NOCBL"):     # experiment: the deposition features without CBL
            kw = dict(ctl=5.0, cblflag=0)
        extra = dict(nspec=2, drydepspec=(1, 1), wetdepspec=(1, 0), weta_gas=(2.0e-5, -1.0), wetb_gas=(0.62, -1.0),
                     henry=(1.0e-2, 0.0), nest=(-30.0, 20.0, 240, 160, 0.125, 0.125))
        zmax, lat = 2000.0, (-60.0, 60.0)
    else:
        kw = dict(ctl=5.0)             # Hanna, method 1
        zmax, lat = 2000.0, (-60.0, 60.0)
    each = max(1, npart_total // nrel)
    cb = fb.make_config(nx=721, ny=361, nz=138, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0,
                        lsynctime=900, ifine=4, outlon0=-180.0, outlat0=-90.0, numxgrid=720,
                        numygrid=360, dxout=0.5, dyout=0.5,
                        outheights=(100.0, 250.0, 500.0, 1000.0, 2000.0, 3000.0, 5000.0, 8000.0, 12000.0, 50000.0),
                        lage=(86400 * 20,), ioutputforeachrelease=0, npart=(each,) * nrel,
                        **({"nspec": 1} if "nspec" not in extra else {}), **extra,
                        maxpart=each * nrel, device=device, rng_mode=fb.RNG_PHILOX_INDEX,
                        math_mode=fb.MATH_FAST, scatter_mode=fb.SCATTER_ATOMIC,
                        part_id_stride=world, part_id_offset=rank, sort_interval=args.sort_interval, **kw)
    # every rank releases its share of the same 100 boxes (round-robin partition of one global
    # problem, src/releaseparticles_mpi.f90:141-152); the ranks differ by their ran1 seed offset
    rel = cases.releases_boxes(cb, seed=100, zmax=zmax, lat_range=lat, width=10.0)
    return cb, rel


def workload_config(args, n, world):
    """`config` of the JSON line: the same dict for the GPU arm and the reference arm."""
    return {
        "workload": ("C2: 1M particles/GPU, global 0.5deg x 138 levels, Hanna CTL=5 IFINE=4, "
                     "100 box releases 0-2 km, lsynctime 900 s, conccalc every step, "
                     "grid exchange every 4 steps") if args.workload == "c2" else
                    ("C3: CBL skewed turbulence (CTL=10, IFINE=5), 2 species with dry deposition, wet deposition, "
                     "nested output grid, 0.5deg x 138 levels") if args.workload == "c3" else
                    ("C5 slice: domain-spread particles/GPU, 0.5deg x 138 levels, CTL=-5 (method 0), "
                     "conccalc every step"),
        "particles_per_gpu": n, "grid": "721x361x138", "rng": "philox-indexed rannumb",
        "math": "fast", "scatter": "atomic", "sort_interval": args.sort_interval,
        "l2": "no flush: each step streams the particle state (132 B/particle) and gathers "
              "from a 2.9 GB met replica, both larger than the 126 MB L2",
        "parallelism": f"particle-partition x{world}",
    }


def host_particles(cb, rel, pinned, mp_pid=0):
    import flexpart_b200 as fb
    parts = fb.Particles(cb.cfg.maxpart, cb.cfg.nspec, pinned=pinned)
    st = fb.ReleaseState(cb.cfg.numpoint, mp_pid=mp_pid)
    fb.release_particles(cb, rel, st, 0, parts)
    return parts


def run_ours(args):
    import torch
    import torch.distributed as dist
    import flexpart_b200 as fb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; flexpart_b200 has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        # one process per GPU: run on the cores next to that GPU, so that the pinned particle arrays
        # of the end-to-end leg are allocated on its NUMA node (torchrun does not bind ranks)
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
            N.nvmlDeviceSetCpuAffinity(N.nvmlDeviceGetHandleByIndex(idx))
        except Exception as e:
            log(f"[rank {rank}] no CPU affinity set ({e})")
        with _stdout_to_stderr():            # (NCCL prints its version banner on stdout)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    K, W = args.steps, args.warmup

    if args.workload == "c5":
        # the C5 leg alone (profiling): the line's headline fields are the c5_strong ones
        cbm = c5_config(1024, local, rank, world)
        span = max(10800, (K + W + 6) * 900)
        mets = (fb.MetFields(cbm).synth(0), fb.MetFields(cbm).synth(span))
        c5 = c5_strong(args, rank, world, local, mets, span, K=K, W=W)
        if rank == 0:
            line = {"metric": "particle-steps/s (advance+conccalc)", "value": c5["value"], "unit": "particle-steps/s",
                    "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": c5["ms_per_step"], "higher_is_better": True,
                    "scaling": "strong", "vs_baseline": None, "dtype": "f32 (f64 positions)", "data": "synthetic",
                    "config": {"workload": c5["workload"]}, "roofline": c5["roofline"], "c5_strong": c5,
                    "gpu_launches": c5.get("gpu_launches")}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    t0 = time.time()
    cb, rel = build_workload(args, rank, world, local)
    c = cb.cfg
    span = max(10800, (K * 3 + W + 6) * 900)
    m0, m1 = fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(span)
    log(f"[rank {rank}] met synthesised in {time.time() - t0:.1f}s")
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    eng.upload_met(1, m0)
    t_up = time.perf_counter()
    eng.upload_met(2, m1)                       # pageable source, synchronous (the classic call)
    t_up = time.perf_counter() - t_up
    eng.set_met_bracket((1, 2), (0, span))
    met_bytes = sum(getattr(m1, nm).nbytes for nm in ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol", "hmix",
                                                      "ustar", "wstar", "oli", "tropopause"))
    parts = host_particles(cb, rel, pinned=True, mp_pid=rank)
    n = parts.numpart
    eng.push_particles(parts)
    ext = torch.cuda.ExternalStream(eng.stream, device=local)

    # Grid exchange off the critical path, behind the C ABI (fpb_reduce_grids_begin/_end): the
    # interval's grids are copied to staging buffers and zeroed on the engine's stream, and ONE NCCL
    # reduce group sums the staging buffers to rank 0 from a high-priority side stream while the next
    # interval's steps compute.  Rank 0 reads the sums from the staging buffers (concoutput's input).
    comm = GridExchange(eng, rank, world, local)
    exchange, exchange_wait = comm.exchange, comm.wait

    def one_step(k, stats):
        itime = k * 900
        if c.wetdep and itime != 0:
            eng.wetdepo(itime, 900, 450)
        eng.conccalc(itime, 1.0)
        st = eng.step(itime, 0, stats=stats)
        if (k + 1) % 4 == 0:
            exchange()
        return st

    with torch.cuda.stream(ext):
        k = 0
        # clocks / throttle reasons are sampled from the warm-up to the end of the
        # end-to-end leg (the device-timed region alone lasts a few tens of ms)
        sampler = ClockSampler(local)
        sampler.start()
        # met read-ahead: the next time level goes into the third slot from page-locked arrays while
        # the warm-up steps run (fpb_upload_met_begin/_end)
        pin = [getattr(m1, nm) for nm in ("uu", "vv", "ww", "rho", "drhodz", "tt", "uupol", "vvpol", "hmix", "ustar",
                                          "wstar", "oli", "tropopause")]
        eng.host_register(*pin)
        t_w = time.perf_counter()
        eng.upload_met_begin(3, m1)
        t_begin = time.perf_counter() - t_w
        for _ in range(W):
            one_step(k, False)
            k += 1
        up_ms = eng.upload_met_end()
        eng.host_unregister(*pin)
        # ---- device-timed region: K steps, inputs resident in HBM
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0 = eng.launch_count
        ev0.record(ext)
        psteps = nsub = npbl = 0
        t_step_k = t_conc_k = 0.0
        for _ in range(K):
            st = one_step(k, True)
            k += 1
            psteps += st["n_active"]; nsub += st["n_substeps"]; npbl += st["n_pbl"]
            a, b = eng.kernel_times()
            t_step_k += a; t_conc_k += b
        exchange_wait()                # the last reduce is inside the timed region
        ev1.record(ext)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count - l0

        # ---- result check of the exchange (untimed): one sample of every active particle on every
        # rank, summed to rank 0, must hold weight * sum(mass) of all ranks' particles (the kernel
        # weights of conccalc sum to 1 and the output grid is global)
        eng.zero_conc_grids()
        eng.conccalc(k * 900, 1.0)
        gsum = comm.reduced_sum()
        hchk = fb.Particles(c.maxpart, c.nspec)
        hchk.numpart = n
        eng.pull_particles(hchk)
        msum = float(hchk.xmass1[:n][hchk.itra1[:n] == k * 900].astype(np.float64).sum())
        mt = torch.tensor([msum], device=f"cuda:{local}", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.SUM)
        msum_all = mt.item()
        if rank == 0:
            log(f"[rank 0] exchange check: sum(reduced gridunc) / sum(mass over {world} rank(s)) = {gsum / msum_all:.7f}; "
                f"reduce {np.mean(comm.ms) if comm.ms else 0.0:.3f} ms per exchange on rank 0")
            assert abs(gsum / msum_all - 1.0) < 5e-5, (gsum, msum_all)
        del hchk

        # ---- end-to-end: host buffers through the C ABI's host-buffer entry point;
        # every step copies the particle arrays in from pinned memory and the
        # arrays the loop writes back out (chunked, copies overlapped with kernels)
        hp = fb.Particles(c.maxpart, c.nspec, pinned=True)
        hp.numpart = n
        eng.pull_particles(hp)
        # what fpb_step_host uploads: the arrays the loop reads (not itrasplit; xscav_frac1 only in backward deposition runs)
        bytes_in = 2 * 8 + 4 + 4 * 5 + 4 * 6 + 2 + 4 * c.nspec * (2 if (c.drybkdep or c.wetbkdep) else 1)
        bytes_out = 2 * 8 + 4 * 3 + 4 * 6 + 2 + 4 * c.nspec      # xtra1..ztra1, itra1, idt, 6 velocities, cbt, xmass1
        for _ in range(3):                       # untimed: lane streams / sort work areas get created
            eng.step_host(hp, k * 900, 0, conc_weight=1.0)
            k += 1
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_e0 = time.perf_counter()
        e_steps = 0
        KE = max(3, min(K, 12))
        for _ in range(KE):
            st = eng.step_host(hp, k * 900, 0, conc_weight=1.0)
            if (k + 1) % 4 == 0:
                exchange()
            k += 1
            e_steps += st["n_active"]
        exchange_wait()
        torch.cuda.synchronize()
        t_e = time.perf_counter() - t_e0
        clocks = sampler.stop()

    log(f"[rank {rank}] {ms / K:.3f} ms/step device-resident (step kernels {t_step_k / K:.3f}, conccalc {t_conc_k / K:.3f}), "
        f"{t_e * 1e3 / KE:.3f} ms/step end to end")
    tm = torch.tensor([ms, t_e * 1e3], device=f"cuda:{local}", dtype=torch.float64)
    cnt = torch.tensor([psteps, e_steps, nsub, npbl, launches], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max, e_ms_max = tm.tolist()
    psteps_all, e_steps_all, nsub_all, npbl_all, launches_all = cnt.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = npbl * B_PBL + (psteps - npbl) * B_FT   # rank 0's kernel launches
        achieved = alg_bytes / (t_step_k * 1e-3) / 1e9 if t_step_k > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("workload") == args.workload and tj.get("particles") == args.particles:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": "particle-steps/s (advance+conccalc)",
            "value": psteps_all / (ms_max * 1e-3),
            "unit": "particle-steps/s",
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 positions)", "data": "synthetic",
            "config": workload_config(args, n, world),
            "substeps_per_s": nsub_all / (ms_max * 1e-3),
            "substeps_per_particle_step": nsub_all / max(psteps_all, 1),
            "pbl_fraction": npbl_all / max(psteps_all, 1),
            "clocks": clocks,
            "gpu_launches": int(launches_all),
            "met_upload": {"bytes_per_time_level": int(met_bytes),
                           "pageable_sync_ms": t_up * 1e3, "pageable_sync_GBps": met_bytes / t_up / 1e9,
                           "pinned_read_ahead_device_ms": up_ms, "pinned_read_ahead_GBps": met_bytes / (up_ms * 1e-3) / 1e9,
                           "read_ahead_host_call_ms": t_begin * 1e3,
                           "what": "fpb_upload_met (13 fields -> 6 fused copy+pack groups) of one 0.5deg x 138 time level on "
                                   "rank 0; read-ahead = fpb_upload_met_begin into the third slot from page-locked arrays "
                                   "while the warm-up steps run, fpb_upload_met_end"},
            "exchange": {"what": "fpb_reduce_grids_begin/_end: staging copy + zero on the engine stream, one NCCL "
                                 "reduce group (sum to rank 0) on a high-priority side stream; every 4th step",
                         "reduce_ms_rank0": float(np.mean(comm.ms)) if comm.ms else 0.0,
                         "mass_check_ratio": gsum / msum_all},
            "e2e": {"value": e_steps_all / (e_ms_max * 1e-3), "unit": "particle-steps/s",
                    "h2d_bytes_per_step": int(bytes_in * n), "d2h_bytes_per_step": int(bytes_out * n + 64),
                    "steps": KE, "what": "fpb_step_host per step: all particle arrays H2D from pinned host "
                                         "memory, conccalc + particle loop, written arrays + stats D2H; "
                                         "row chunks pipelined on 3 streams"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fpb_pbl_kernel + fpb_finish_kernel (= fpb_step)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / K,
                         "kernel_ms_per_launch": t_step_k / K,
                         "conccalc_ms_per_launch": t_conc_k / K,
                         "traffic_source": "profiles/traffic.json (ncu --set full dram__bytes_read+write, "
                                           "pbl + finish kernel, one launch)" if traffic else None,
                         "note": "algorithmic bytes = 776 B per PBL particle-step, 552 B above the PBL "
                                 "(state 132 B + 161/105 met floats); method-1 sub-steps make C2 "
                                 "ALU/latency-bound, see substeps_per_particle_step"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = (cpu_baseline_reference(args, 1, args.cpu_seconds) if reference_available()
                                    else cpu_baseline(args, threads=1, budget_s=args.cpu_seconds))
    if rank == 0 and world == 1 and not args.no_hbm_regime:
        line["next_rows"] = next_rows(eng, cb, rel, peaks()[0], k * 900, cpu=not args.no_cpu)
    eng.close()
    del hp, parts
    if args.workload == "c2" and not args.no_c5:
        # every rank takes part (strong scaling over the same process group)
        c5 = c5_strong(args, rank, world, local, (m0, m1), span, K=min(K, 12), W=3)
        if rank == 0:
            line["c5_strong"] = c5
            log(f"[rank 0] c5_strong: {c5['value']:.3e} particle-steps/s, {c5['ms_per_step']:.3f} ms/step, "
                f"frac {c5['roofline']['frac']:.3f}, mass ratio {c5['mass_check']['ratio']:.6f}")
    if rank == 0:
        if world == 1 and args.workload == "c2" and not args.no_hbm_regime:
            line["hbm_regime"] = hbm_regime(args, local, K=min(K, 12), W=3)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5_config(n_rank, device, rank, world, sort_interval=8):
    """BASELINE configs[4]: global domain-filling run, 0.5 deg x 138 levels, the shipped CTL=-5 /
    IFINE=4 (method 0: one Langevin step per particle and interval), one species (air-mass tracer)."""
    import flexpart_b200 as fb
    return fb.make_config(nx=721, ny=361, nz=138, dx=0.5, dy=0.5, xlon0=-180.0, ylat0=-90.0,
                          lsynctime=900, ctl=-5.0, ifine=4, mdomainfill=1, outlon0=-180.0, outlat0=-90.0,
                          numxgrid=720, numygrid=360, dxout=0.5, dyout=0.5,
                          outheights=(100.0, 250.0, 500.0, 1000.0, 2000.0, 3000.0, 5000.0, 8000.0, 12000.0, 100000.0),
                          lage=(86400 * 20,), ioutputforeachrelease=0, npart=(C5_TOTAL,), nspec=1,
                          maxpart=n_rank, device=device, rng_mode=fb.RNG_PHILOX_INDEX, math_mode=fb.MATH_FAST,
                          scatter_mode=fb.SCATTER_ATOMIC, part_id_stride=world, part_id_offset=rank,
                          sort_interval=sort_interval)


C5_TOTAL = 100_000_000


def c5_strong(args, rank, world, local, mets, span, K, W):
    """BASELINE configs[4] / north_star target: 100 M domain-filling particles in total, STRONG
    scaling (100 M / N per GPU, created on the device by fpb_init_domainfill, every N-th particle of
    the one global set per rank), device-resident steps (conccalc + particle loop, cell sort every
    8th step), grid exchange every 4th step.  Returns the extra key of the JSON line (rank 0)."""
    import torch
    import torch.distributed as dist
    import flexpart_b200 as fb
    total = int(os.environ.get("FPB_C5_TOTAL", C5_TOTAL))
    n_rank = (total + world - 1) // world + 1024
    cb = c5_config(n_rank, local, rank, world, sort_interval=int(os.environ.get("FPB_C5_SORT", "8")))
    cb.cfg.npart[0] = total
    cb.npart[0] = total
    c = cb.cfg
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    t0 = time.perf_counter()
    eng.upload_met(1, mets[0]); eng.upload_met(2, mets[1])
    t_met = time.perf_counter() - t0
    eng.set_met_bracket((1, 2), (0, span))
    t0 = time.perf_counter()
    n, info = eng.init_domainfill((0.0, 0.0, float(c.nx - 1), float(c.ny - 1)))
    t_fill = time.perf_counter() - t0
    ext = torch.cuda.ExternalStream(eng.stream, device=local)
    comm = GridExchange(eng, rank, world, local)

    def one_step(k, stats):
        itime = k * 900
        eng.conccalc(itime, 1.0)
        st = eng.step(itime, 0, stats=stats)
        if (k + 1) % 4 == 0:
            comm.exchange()
        return st

    with torch.cuda.stream(ext):
        k = 0
        for _ in range(W):
            one_step(k, False); k += 1
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0.record(ext)
        l0 = eng.launch_count
        ps = npbl = 0
        tk = tc = 0.0
        for _ in range(K):
            st = one_step(k, True); k += 1
            ps += st["n_active"]; npbl += st["n_pbl"]
            a, b = eng.kernel_times()
            tk += a; tc += b
        comm.wait()
        ev1.record(ext)
        torch.cuda.synchronize()
        launches = eng.launch_count - l0
        if world > 1:
            dist.barrier()
        ms = ev0.elapsed_time(ev1)
        # mass check of the reduced grid: one sample of every particle, summed over the ranks, must
        # hold the air mass the particles were created with (kernel weights sum to 1, the grid is
        # global and reaches above the model top)
        comm.wait()
        eng.zero_conc_grids()
        eng.conccalc(k * 900, 1.0)
        gsum = comm.reduced_sum()
    tm = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
    cnt = torch.tensor([ps, npbl, n], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    eng.close()
    if rank != 0:
        return None
    ms_max = tm.item()
    ps_all, npbl_all, n_all = cnt.tolist()
    peak, _ = peaks()
    alg = npbl * B_PBL + (ps - npbl) * B_FT
    ach = alg / (tk * 1e-3) / 1e9 if tk > 0 else 0.0
    mass_ratio = gsum / info["colmasstotal"]
    return {"workload": f"C5: global domain-fill, {total} particles in total over {world} GPU(s) (strong scaling), "
                        "0.5deg x 138 levels, shipped CTL=-5 IFINE=4 (method 0), conccalc every step, "
                        "cell sort every 8th step, grid exchange every 4th step",
            "scaling": "strong", "value": ps_all / (ms_max * 1e-3), "unit": "particle-steps/s",
            "n_gpus": world, "particles_total": int(n_all), "particles_rank0": int(n), "steps": K,
            "ms_per_step": ms_max / K, "kernel_ms_per_launch": tk / K, "conccalc_ms_per_launch": tc / K,
            "pbl_fraction": npbl_all / max(ps_all, 1), "gpu_launches": int(launches),
            "exchange_reduce_ms_rank0": float(np.mean(comm.ms)) if comm.ms else 0.0,
            "init_domainfill_ms": t_fill * 1e3, "met_upload_ms_per_slot": t_met * 1e3 / 2,
            "mass_check": {"grid_sum_over_ranks": gsum, "colmasstotal": info["colmasstotal"], "ratio": mass_ratio,
                           "ok": bool(abs(mass_ratio - 1.0) < 2e-3)},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "kernel": "fpb_pbl_kernel + fpb_finish_kernel (rank 0)",
                         "algorithmic_bytes_per_launch": alg / K}}


class _stdout_to_stderr:
    """fd-level redirect: keeps native libraries' banners (NCCL) off the one-JSON-line stdout"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        import ctypes
        ctypes.CDLL(None).fflush(None)
        os.dup2(self.saved, 1)
        os.close(self.saved)


class GridExchange:
    """The mpif_tm_reduce_grid slot (src/mpi_mod.f90:2395-2579, src/timemanager_mpi.f90:468-485)
    through the C ABI: fpb_comm_init joins the NCCL communicator (the 128-byte id travels over the
    host's own channel: here torch.distributed, MPI_Bcast in the Fortran host), and
    fpb_reduce_grids_begin/_end do the staged, overlapped sum to rank 0."""

    def __init__(self, eng, rank, world, local):
        import flexpart_b200 as fb
        self.eng, self.rank, self.world, self.local = eng, rank, world, local
        self.ms, self.pending = [], False
        with _stdout_to_stderr():            # NCCL prints its version banner on stdout
            uid = [fb.Engine.comm_unique_id() if (rank == 0 and world > 1) else bytes(128)]
            if world > 1:
                import torch.distributed as dist
                dist.broadcast_object_list(uid, src=0)
            eng.comm_init(uid[0], rank, world)
            # the first collective of a communicator sets up its NVLink channels (~1 s): not part of any step
            eng.reduce_grids_begin()
            eng.reduce_grids_device(0)

    def wait(self):
        if self.pending:
            _, _, ms = self.eng.reduce_grids_device(0)   # host waits for the reduce
            self.ms.append(ms)
            self.pending = False

    def exchange(self):
        self.wait()
        self.eng.reduce_grids_begin()
        self.pending = True

    def reduced_sum(self):
        """exchange now; sum over all cells of the summed grid (float64), valid on rank 0"""
        import torch
        self.wait()
        self.eng.reduce_grids_begin()
        ptr, n, ms = self.eng.reduce_grids_device(0)

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        g = torch.as_tensor(_Holder(), device=f"cuda:{self.local}")
        return float(g.double().sum().item())


def next_rows(eng, cb, rel, peak, itime_now, cpu=True):
    """The SURVEY 8f rows that run on the device, timed on the bench's own engine (wall clock
    around the synchronous C-ABI calls, best of 5): releaseparticles for all particles at once,
    concoutput's sparse dump of the concentration grid, wetdepo when the workload has it."""
    import numpy as np
    import flexpart_b200 as fb
    c = cb.cfg
    out = {}

    def best(fn, n=5):
        t = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
        return min(t)

    # sparse dump of gridunc(:,:,:,1,1,:,1) as the steps left it
    area, vol = fb.outgrid_geometry(cb, c.ylat0 - c.youtshift)
    eng.set_outgrid_geometry(area, vol)
    eng.conccalc(0, 1.0)
    res = {}
    t = best(lambda: res.update(d=eng.concoutput_sparse(0, 1, 1, 1, 4.0)))
    cells = c.numxgrid * c.numygrid * c.numzgrid
    nz = int(res["d"][1].size)
    alg = 2 * cells * (4 * c.nclassunc + 4) + 8 * nz        # two passes over the cells + the lists
    out["concoutput_sparse"] = {"ms": t * 1e3, "cells": cells, "nonzero": nz, "runs": int(res["d"][0].size),
                                "d2h_bytes": 4 * (nz + int(res["d"][0].size)), "dense_d2h_bytes": 4 * cells,
                                "alg_GBps": alg / t / 1e9, "frac_of_peak": alg / t / 1e9 / peak,
                                "what": "fpb_concoutput_sparse incl. the D2H of the compacted lists"}
    if c.wetdep:
        t = best(lambda: eng.wetdepo(itime_now, 900, 450))
        out["wetdepo"] = {"ms": t * 1e3, "particles_per_s": c.maxpart / t,
                          "alg_GBps": c.maxpart * (28 + 64 + 5 + 8 * c.nspec) / t / 1e9,
                          "what": "fpb_wetdepo over all particles (28 B state + 4 x 16 B rain words + cloud class + tt + masses)"}
    # convective mixing (LCONVECTION = 1, the shipped default) for the bench's own particles: occupied
    # columns -> Emanuel scheme per column -> redist, all on the device
    try:
        import conv_cases
        akm, bkm, akz, bkz, nconvlev = conv_cases.hybrid_levels(c.nz)
        f0 = conv_cases.conv_fields(cb, akz, bkz, c.nz, 1)
        eng.set_convection(c.nz, c.nzmax, nconvlev, akz[1:], bkz[1:], akm[1:], bkm[1:])
        eng.upload_convmet(1, *f0); eng.upload_convmet(2, *f0)
        eng.upload_convmet(3, *f0)
        res = {}
        t = best(lambda: res.update(n=eng.convmix(itime_now)), n=3)
        out["convmix"] = {"ms": t * 1e3, "occupied_columns": res["n"][0], "convecting_columns": res["n"][1],
                          "particles": c.maxpart, "nconvlev": nconvlev,
                          "what": "fpb_convmix: column sort, calcmatrix + Emanuel scheme (O(n) part one thread per occupied "
                                  "column, level-pair loops row-parallel per group of 32 columns, flux assembly + "
                                  "redistribution matrix one block per column), redist; synthetic soundings "
                                  "(tests/conv_cases.py)"}
    except Exception as e:  # (the convection leg must not take the bench line down)
        out["convmix"] = {"error": str(e)}
    # calcpar + verttransform_ecmwf on the device (what getfields does to every new wind field): the raw
    # model-level field goes up once, the met slot is built there (slot 3, the read-ahead slot)
    try:
        import met_cases
        akm, bkm, akz, bkz, _ = conv_cases.hybrid_levels(c.nz)
        raw = met_cases.raw_fields(cb, akz, bkz, c.nz, seed=1)
        eng.set_vertical(c.nz, akm[1:], bkm[1:], akz[1:], bkz[1:])
        eng.calcpar_verttransform(3, raw)
        ms, kms = [], []
        for _ in range(3):
            ms.append(eng.calcpar_verttransform(3, raw)); kms.append(eng.metproc_kernel_ms)
        raw_bytes = sum(int(a[:c.nx, :c.ny].size) * 4 for k, a in raw.items())
        n3 = c.nx * c.ny * c.nz
        alg = 4 * n3 * (6 + 12)      # 6 raw 3-D fields read, 12 transformed 3-D values written per grid point
        out["getfields"] = {"ms": min(ms), "kernels_ms": min(kms), "grid": f"{c.nx}x{c.ny}x{c.nz}",
                            "raw_field_bytes": raw_bytes, "upload_GBps": raw_bytes / ((min(ms) - min(kms)) * 1e-3) / 1e9,
                            "kernels_alg_GBps": alg / (min(kms) * 1e-3) / 1e9,
                            "kernels_frac_of_peak": alg / (min(kms) * 1e-3) / 1e9 / peak,
                            "what": "fpb_calcpar_verttransform: pageable raw wind field (uuh, vvh, wwh, tth, qvh, 2-D "
                                    "fields) H2D + calcpar, calcpv, verttransform_ecmwf kernels (device time)"}
        if cpu:     # the reference's own routines (oracle/_ref) on a bounded sample: 1 deg x same levels
            try:
                import flexpart_b200 as fb2
                import cases as _cases
                from metproc_common import reference_run
                cbs = _cases.config_small(nrel=1, npart_each=8, nx=361, ny=181, nz=c.nz, height=fb2.synth_heights(c.nz))
                raws = met_cases.raw_fields(cbs, akz, bkz, c.nz, seed=1)
                tm = {}
                reference_run(cbs, raws, akm, bkm, akz, bkz, c.nz, timing=tm)
                out["getfields"]["cpu_reference"] = {
                    "grid": f"361x181x{c.nz}", "calcpar_s": tm["calcpar_s"], "verttransform_s": tm["verttransform_s"],
                    "cores": 1, "kind": "reference",
                    "what": "src/calcpar.f90 + src/verttransform_ecmwf.f90 from the reference's sources "
                            "(oracle/f2c, gcc -O2), one core, a quarter of the bench grid's columns"}
            except Exception as e:
                out["getfields"]["cpu_reference"] = {"error": str(e)}
    except Exception as e:
        out["getfields"] = {"error": str(e)}
    # release of maxpart particles on a second engine (Philox positions): no particle row crosses PCIe
    e2 = fb.Engine(cb)
    e2.set_releases(rel)
    t0 = time.perf_counter(); n, made = e2.release_particles(0); t = time.perf_counter() - t0
    out["releaseparticles"] = {"ms": t * 1e3, "particles": made, "particles_per_s": made / t,
                               "alg_GBps": made * 78 / t / 1e9,
                               "what": "fpb_releaseparticles, all release points at once (first call: includes allocations)"}
    e2.close() if hasattr(e2, "close") else None
    return out


def hbm_regime(args, local, K, W):
    """Secondary figure for the JSON line: the same path in its HBM-gather-bound
    regime (CTL=-5 -> method 0, one Langevin step per particle and interval,
    particles spread over the domain up to 12 km: the shipped options/COMMAND
    setting, SURVEY.md 8d 'C5 slice').  Device-resident, kernels + sort."""
    import torch
    import flexpart_b200 as fb
    a2 = argparse.Namespace(**vars(args))
    a2.workload, a2.particles, a2.sort_interval = "c5slice", 8_000_000, 8
    cb, rel = build_workload(a2, 0, 1, local)
    span = max(10800, (K + W + 2) * 900)
    eng = fb.Engine(cb)
    eng.fill_rannumb()
    eng.upload_met(1, fb.MetFields(cb).synth(0))
    eng.upload_met(2, fb.MetFields(cb).synth(span))
    eng.set_met_bracket((1, 2), (0, span))
    parts = host_particles(cb, rel, pinned=False)
    eng.push_particles(parts)
    ext = torch.cuda.ExternalStream(eng.stream, device=local)
    with torch.cuda.stream(ext):
        for k in range(W):
            eng.conccalc(k * 900, 1.0); eng.step(k * 900, 0, stats=False)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ev0.record(ext)
        ps = npbl = 0
        tk = 0.0
        for k in range(W, W + K):
            eng.conccalc(k * 900, 1.0)
            st = eng.step(k * 900, 0, stats=True)
            ps += st["n_active"]; npbl += st["n_pbl"]
            tk += eng.kernel_times()[0]
        ev1.record(ext)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    eng.close()
    peak, _ = peaks()
    alg = npbl * B_PBL + (ps - npbl) * B_FT
    ach = alg / (tk * 1e-3) / 1e9
    return {"workload": "C5 slice: 8M particles spread over the globe up to 12 km, CTL=-5 (method 0), "
                        "conccalc every step, cell sort every 8th step",
            "value": ps / (ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms / K,
            "kernel_ms_per_launch": tk / K, "pbl_fraction": npbl / max(ps, 1),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "kernel": "fpb_pbl_kernel + fpb_finish_kernel"}}


def cpu_baseline(args, threads, budget_s, steps=None):
    """Oracle (line-faithful CPU restatement of the reference path) timed on
    the host cores on a bounded sample of the same workload."""
    import flexpart_b200 as fb
    from oracle_api import Oracle, load
    import ctypes as C
    t0 = time.time()
    # calibrate on a tiny sample, then size the sample for ~budget_s of work
    def make(nsample):
        a2 = argparse.Namespace(**vars(args))
        cb, rel = build_workload(a2, 0, 1, 0, n_particles=nsample)
        return cb, rel
    cb, rel = make(100 * 20)
    span = 10800
    m0, m1 = fb.MetFields(cb).synth(0), fb.MetFields(cb).synth(span)

    def run(nsample, nsteps):
        per = max(1, nsample // threads)
        states, keep = [], []
        for t in range(threads):
            cbt, relt = make(per)
            o = Oracle(cbt)
            o.fill_rannumb()
            o.upload_met(1, m0); o.upload_met(2, m1)
            o.set_met_bracket((1, 2), (0, span))
            p = host_particles(cbt, relt, pinned=False)
            o.push_particles(p)
            states.append(o); keep.append((cbt, relt, p))
        arr = (C.c_void_p * threads)(*[o.S for o in states])
        L = load()
        wall = 0.0
        total = 0
        for k in range(nsteps):
            wall += L.fpo_mp_step(arr, threads, k * 900, 0, 1.0)
            total += sum(kk[2].numpart for kk in keep)
        for o in states:
            o.close()
        return total / wall, total, wall
    rate, _, _ = run(2000 * threads, 2)
    nsteps = steps or 4
    nsample = int(max(2000 * threads, min(rate * budget_s / nsteps, 2_000_000)))
    nsample = (nsample // (100 * threads)) * 100 * threads
    rate, total, wall = run(nsample, nsteps)
    return {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"{nsample} particles x {nsteps} steps of the same workload "
                      f"({total} particle-steps in {wall:.1f} s; oracle/ C restatement, gcc -O2, "
                      f"{threads} thread(s); the Fortran reference cannot be built here)",
            "setup_s": round(time.time() - t0 - wall, 1)}


# ---- the reference's own code (oracle/_ref/libflexref.so: the reference's Fortran sources
# for the path transpiled to C at build time, oracle/f2c/) as the CPU baseline
_REF_CTX = {}


def _ref_worker(job):
    """one process = one FLEXPART_MPI rank: its share of the particles, the whole met (the
    arrays filled before the fork are shared copy-on-write)"""
    w, nproc, nsteps = job
    import flexpart_b200 as fb
    ref, parts, cb = _REF_CTX["ref"], _REF_CTX["parts"], _REF_CTX["cb"]
    idx = np.arange(w, parts.numpart, nproc)
    sub = fb.Particles(cb.cfg.maxpart, cb.cfg.nspec)
    for nm in ("xtra1", "ytra1", "ztra1", "itra1", "itramem", "npoint", "nclass", "idt", "uap", "ucp", "uzp",
               "us", "vs", "ws", "cbt"):
        getattr(sub, nm)[:len(idx)] = getattr(parts, nm)[idx]
    sub.xmass1[:len(idx)] = parts.xmass1[idx]
    sub.numpart = len(idx)
    ref.push_state(sub)
    total = 0
    t0 = time.perf_counter()
    for k in range(nsteps):
        itime = k * 900
        total += int(np.sum(ref.arr("itra1")[:len(idx)] == itime))
        ref.conccalc(itime, 1.0)
        ref.particle_loop(itime, 0)
    return total, time.perf_counter() - t0


def cpu_baseline_reference(args, threads, budget_s, steps=None):
    """The reference's own particle loop (src/timemanager.f90:531-712 with initialize / advance
    and everything below them) + conccalc, from oracle/_ref/libflexref.so, one process per
    host core (the FLEXPART_MPI execution model, README_PARALLEL.md:60-73), bounded sample."""
    import multiprocessing as mp
    import flexpart_b200 as fb
    import ref_api
    t_setup = time.time()

    def run(nsample, nsteps, nproc):
        a2 = argparse.Namespace(**vars(args))
        cb, rel = build_workload(a2, 0, 1, 0, n_particles=nsample)
        ref = ref_api.Ref(cb, maxrand=100000)
        ref.fill_rannumb()
        ref.upload_met(1, fb.MetFields(cb).synth(0))
        ref.upload_met(2, fb.MetFields(cb).synth(10800))
        ref.set_met_bracket((1, 2), (0, 10800))
        parts = host_particles(cb, rel, pinned=False)
        _REF_CTX.update(ref=ref, parts=parts, cb=cb)
        with mp.get_context("fork").Pool(nproc) as pool:
            res = pool.map(_ref_worker, [(w, nproc, nsteps) for w in range(nproc)])
        total, wall = sum(r[0] for r in res), max(r[1] for r in res)
        return total / wall, total, wall
    nsteps = steps or 4
    rate, _, _ = run(2000 * threads, 2, threads)
    nsample = int(max(2000 * threads, min(rate * budget_s / nsteps, 4_000_000)))
    nsample = (nsample // (100 * threads)) * 100 * threads
    rate, total, wall = run(nsample, nsteps, threads)
    return {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "reference",
            "sample": f"{nsample} particles x {nsteps} steps of the same workload ({total} particle-steps in "
                      f"{wall:.1f} s; the reference's own Fortran sources for the path, transpiled to C "
                      f"(oracle/f2c) and compiled gcc -O2, {threads} process(es) each with a share of the particles)",
            "setup_s": round(time.time() - t_setup - wall, 1)}


def reference_available():
    try:
        import ref_api
        return ref_api.available()
    except Exception:
        return False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # each step is a bounded sample; keep the whole K+W run within a few minutes
    per_step_budget = max(2.0, min(20.0, 150.0 / max(K + W, 1)))
    if reference_available():
        cb = cpu_baseline_reference(args, threads=threads, budget_s=per_step_budget * K, steps=K)
    else:
        cb = cpu_baseline(args, threads=threads, budget_s=per_step_budget * K, steps=K)
    line = {"impl": "reference", "metric": "particle-steps/s (advance+conccalc)", "value": cb["value"],
            "unit": "particle-steps/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": K,
            "warmup": W, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (f64 positions)", "data": "synthetic",
            "config": workload_config(args, args.particles, int(os.environ.get("WORLD_SIZE", "1"))),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c5slice", "c3", "c5"])
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hbm-regime", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 strong-scaling leg (extra key c5_strong)")
    ap.add_argument("--sort-interval", type=int, default=1)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    import __graft_entry__ as ge
    ge.build(verbose=False)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

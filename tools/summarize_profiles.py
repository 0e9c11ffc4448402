#!/usr/bin/env python
"""tools/summarize_profiles.py ROUND -- gpurun_out/ (written by tools/profile_round.sh) -> profiles/:
copies the bench lines, launch list and full-capture exports, writes traffic.json (read by bench.py
as roofline.traffic) and prints the tables of profiles/README.md."""
import csv, io, json, os, shutil, sys
from collections import OrderedDict

R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def rows(path):
    txt = [l for l in open(path) if l.startswith('"')]
    return list(csv.DictReader(io.StringIO("".join(txt))))


def short(name):
    name = name.replace("<unnamed>::", "")
    if name.startswith("void "):
        name = name[5:]
    return name.split("(")[0]


def last_json(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


for src, dst in ((f"bench_{R}_c2.json", f"bench_{R}_c2.json"), (f"bench_{R}_ref.json", f"bench_{R}_reference_arm.json"),
                 (f"bench_{R}_c3.json", f"bench_{R}_c3.json"), (f"launches_{R}.csv", f"launches_{R}.csv"),
                 (f"ncu_full_{R}_raw.csv", f"ncu_full_{R}_raw.csv"), (f"ncu_full_{R}_details.txt", f"ncu_full_{R}_details.txt")):
    shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# ---- launch list: one device-resident step (the last conccalc..finish span before the e2e part)
L = rows(os.path.join(G, f"launches_{R}.csv"))
names = [short(r["Kernel Name"]) for r in L]
ns = [float(r["Metric Value"]) for r in L]
conc = [i for i, n in enumerate(names) if n.startswith("fpb_conccalc_kernel")]
fin = [i for i, n in enumerate(names) if n.startswith("fpb_finish_kernel")]
# steady resident step = from the conccalc of step k to just before the conccalc of step k+1; take the 5th
a, b = conc[4], conc[5]
agg = OrderedDict()
for n, t in zip(names[a:b], ns[a:b]):
    c = agg.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += t
tot = sum(v[1] for v in agg.values())
print("| Kernel | launches | µs | share |\n|---|---|---|---|")
for n, (k, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {k} | {t / 1e3:.1f} | {100 * t / tot:.1f} % |")
print(f"| total | {sum(v[0] for v in agg.values())} | {tot / 1e3:.1f} | |")
stepk = sum(v[1] for n, v in agg.items() if n.startswith(("fpb_pbl_kernel", "fpb_finish_kernel")))
print(f"step kernels' share under ncu: {100 * stepk / tot:.1f} %")

# ---- full capture
F = rows(os.path.join(G, f"ncu_full_{R}_raw.csv"))
F = [r for r in F if r.get("Kernel Name")]      # (second row holds the units)
key = {"fpb_conccalc_kernel": "conccalc", "fpb_pbl_kernel": "pbl", "fpb_finish_kernel": "finish"}


def num(r, k):
    try:
        return float(r[k].replace(",", ""))
    except (KeyError, ValueError):
        return None


units = rows(os.path.join(G, f"ncu_full_{R}_raw.csv"))[0]
kern = {}
for r in F:
    n = short(r["Kernel Name"])
    for pre, kk in key.items():
        if n.startswith(pre):
            def scaled(col):
                v, u = num(r, col), units.get(col, "")
                if v is None:
                    return None
                return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
            t, tu = num(r, "gpu__time_duration.sum"), units.get("gpu__time_duration.sum", "ns")
            kern[kk] = {
                "kernel": n,
                "dram_read_bytes": scaled("dram__bytes_read.sum"), "dram_write_bytes": scaled("dram__bytes_write.sum"),
                "gpu_time_ms": t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(tu, 1e-6),
                "issue_active_pct": num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or num(r, "smsp__issue_active.avg.pct"),
                "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
                "registers": num(r, "launch__registers_per_thread"),
                "avg_threads_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "achieved_occupancy_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            }
bench = last_json(os.path.join(G, f"bench_{R}_c2.json"))
tj = {"workload": "c2", "particles": bench["config"]["particles_per_gpu"],
      "source": f"profiles/ncu_full_{R}_raw.csv (ncu --set full --clock-control none, one launch each, 5th step of the C2 bench)",
      "kernels": kern,
      "dram_bytes_per_launch": sum((kern[k]["dram_read_bytes"] or 0) + (kern[k]["dram_write_bytes"] or 0) for k in ("pbl", "finish"))}
json.dump(tj, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print("\n| Kernel | time (ms) | DRAM read (MB) | DRAM write (MB) | L2 hit | issue slots busy | threads / instruction | regs | occupancy |\n|---|---|---|---|---|---|---|---|---|")
for k in ("pbl", "finish", "conccalc"):
    v = kern[k]
    print(f"| `{v['kernel']}` | {v['gpu_time_ms']:.3f} | {v['dram_read_bytes'] / 1e6:.0f} | {v['dram_write_bytes'] / 1e6:.0f} | "
          f"{v['l2_hit_pct']:.0f} % | {v['issue_active_pct']:.0f} % | {v['avg_threads_per_inst']:.1f} | {v['registers']:.0f} | {v['achieved_occupancy_pct']:.0f} % |")
print("dram bytes pbl+finish per launch:", tj["dram_bytes_per_launch"])

# ---- headline numbers
ref, c3 = last_json(os.path.join(G, f"bench_{R}_ref.json")), last_json(os.path.join(G, f"bench_{R}_c3.json"))
print(f"\nC2 device-resident {bench['value']:.3e} ({bench['ms_per_step']:.3f} ms/step), roofline.frac {bench['roofline']['frac']:.3f}, "
      f"kernels {bench['roofline']['kernel_ms_per_launch']:.3f} ms, conccalc {bench['roofline']['conccalc_ms_per_launch']:.3f} ms")
print(f"e2e {bench['e2e']['value']:.3e} ({bench['e2e']['h2d_bytes_per_step'] / 1e6:.0f} MB in, {bench['e2e']['d2h_bytes_per_step'] / 1e6:.0f} MB out)")
if "hbm_regime" in bench:
    h = bench["hbm_regime"]; print("hbm_regime", json.dumps(h)[:600])
print("cpu_baseline", bench.get("cpu_baseline"))
print(f"reference arm {ref['value']:.3e} on {ref['cpu_baseline']['cores']} cores")
print(f"C3 device-resident {c3['value']:.3e} ({c3['ms_per_step']:.3f} ms/step), e2e {c3['e2e']['value']:.3e}, substeps/particle-step {c3['substeps_per_particle_step']:.1f}")

# ---- the other workloads' captures (round 2+): C3 (full-feature instantiation) and C5 (HBM-gather regime)
for tag, label in (("c3", "C3: CBL + deposition + nested output, 1 M particles"), ("c5", "C5: 100 M domain-filling particles, method 0")):
    rawp = os.path.join(G, f"ncu_full_{R}_{tag}_raw.csv")
    if not os.path.exists(rawp):
        continue
    for suffix in (f"ncu_full_{R}_{tag}_raw.csv", f"ncu_full_{R}_{tag}_details.txt", f"launches_{R}_{tag}.csv"):
        if os.path.exists(os.path.join(G, suffix)):
            shutil.copy(os.path.join(G, suffix), os.path.join(P, suffix))
    if os.path.exists(os.path.join(G, f"bench_{R}_{tag}.json")):
        shutil.copy(os.path.join(G, f"bench_{R}_{tag}.json"), os.path.join(P, f"bench_{R}_{tag}.json"))
    FF = [r for r in rows(rawp) if r.get("Kernel Name")]
    uu = rows(rawp)[0]
    print(f"\n### {label} (ncu --set full, one launch each)\n")
    print("| Kernel | time (ms) | DRAM read (MB) | DRAM write (MB) | DRAM throughput | L2 hit | issue slots busy | threads / instruction | regs | occupancy |\n|---|---|---|---|---|---|---|---|---|---|")
    for r in FF:
        def sc(col):
            v, u = num(r, col), uu.get(col, "")
            return None if v is None else v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
        t, tu = num(r, "gpu__time_duration.sum"), uu.get("gpu__time_duration.sum", "ns")
        tms = t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(tu, 1e-6)
        rd, wr = sc("dram__bytes_read.sum") or 0, sc("dram__bytes_write.sum") or 0
        print(f"| `{short(r['Kernel Name'])}` | {tms:.3f} | {rd / 1e6:.0f} | {wr / 1e6:.0f} | "
              f"{(rd + wr) / (tms * 1e-3) / 1e9:.0f} GB/s | {num(r, 'lts__t_sector_hit_rate.pct') or 0:.0f} % | "
              f"{(num(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active') or num(r, 'smsp__issue_active.avg.pct') or 0):.0f} % | "
              f"{num(r, 'smsp__thread_inst_executed_per_inst_executed.ratio') or 0:.1f} | {num(r, 'launch__registers_per_thread') or 0:.0f} | "
              f"{num(r, 'sm__warps_active.avg.pct_of_peak_sustained_active') or 0:.0f} % |")
    lp = os.path.join(G, f"launches_{R}_{tag}.csv")
    if os.path.exists(lp):
        LL = rows(lp)
        nm = [short(r["Kernel Name"]) for r in LL]
        tt = [float(r["Metric Value"]) for r in LL]
        cc = [i for i, n in enumerate(nm) if n.startswith("fpb_conccalc_kernel")]
        if len(cc) >= 3:
            a2, b2 = cc[-2], cc[-1]
            ag = OrderedDict()
            for n, t in zip(nm[a2:b2], tt[a2:b2]):
                c = ag.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += t
            tot2 = sum(v[1] for v in ag.values())
            print(f"\nOne step of that workload (launch list, cold-cache serialised):\n\n| Kernel | launches | µs | share |\n|---|---|---|---|")
            for n, (k, t) in sorted(ag.items(), key=lambda kv: -kv[1][1]):
                print(f"| `{n}` | {k} | {t / 1e3:.1f} | {100 * t / tot2:.1f} % |")

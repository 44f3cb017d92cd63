#!/usr/bin/env python
"""Turn the files tools/evidence.sh brought back in gpurun_out/ into the summaries committed under profiles/.

    python tools/evidence_summary.py TAG [OUT_TAG]

  gpurun_out/TAG_bench_default.json           -> profiles/OUT_bench_default_pile16m.json      (the bench line, verbatim)
  gpurun_out/TAG_ncu_launches_bench.csv       -> profiles/OUT_ncu_launches_bench.csv          (ncu launch list, kernel / duration)
  gpurun_out/TAG_ncu_traffic_<w>.csv          -> profiles/OUT_traffic.json                    (DRAM bytes per launch, LAST substep of the capture;
                                                                                               bench.py copies them into roofline.traffic)
                                                 profiles/OUT_substep_<w>.txt                 (per-kernel table of that substep)
  gpurun_out/TAG_full_<w>.ncu-rep             -> profiles/OUT_ncu_full_<w>.csv                (one summary line per captured launch)
                                                 profiles/OUT_ncu_sections_k_gs_pairs_<w>.txt (shares of k_gs_pairs by source section)
                                                 profiles/OUT_ncu_source_k_gs_pairs_<w>.txt   (hottest source lines)
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

# kernel name -> stage name of bench.py's stage_ms_per_substep
STAGE = {"k_integrate_walls_keys": "integrate_walls_keys", "k_scan_reduce": "cell_scan", "k_scan_tiles": "cell_scan", "k_scan_apply": "cell_scan",
         "k_scan_lookback": "cell_scan", "k_cell_scatter": "cell_sort", "k_canonical_gather": "canonical_gather", "k_gs_pairs": "collide",
         "k_gs_resolve": "resolve_wavefront", "k_collide_tile2": "collide", "k_resolve_events": "resolve_wavefront",
         "k_resolve_wavefront": "resolve_wavefront", "k_wave_pack": "resolve_wavefront"}


def rows_of(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def launches(path):
    """[(id, kernel, {metric: value})] in launch order"""
    out = collections.OrderedDict()
    for r in rows_of(path):
        k = out.setdefault(int(r["ID"]), [r["Kernel Name"].split("(")[0], {}])
        try:
            k[1][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
        k[1][r["Metric Name"] + ".unit"] = r["Metric Unit"]
    return [(i, k, m) for i, (k, m) in out.items()]


def dur_ms(m):
    v, u = m.get("gpu__time_duration.sum", 0.0), m.get("gpu__time_duration.sum.unit", "ns")
    return v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(u, 1e-6)


def nbytes(m, name):
    v, u = m.get(name, 0.0), m.get(name + ".unit", "byte")
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def last_substep(ls):
    """launches of the last substep: from the last k_integrate_walls_keys on"""
    start = max(i for i, (_, k, _) in enumerate(ls) if k == "k_integrate_walls_keys")
    return ls[start:]


def main():
    tag = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else tag
    bench = os.path.join(G, f"{tag}_bench_default.json")
    if os.path.exists(bench):
        line = [l for l in open(bench) if l.startswith("{")][-1]
        with open(os.path.join(P, f"{out}_bench_default_pile16m.json"), "w") as f:
            f.write(line)
    src = os.path.join(G, f"{tag}_ncu_launches_bench.csv")
    if os.path.exists(src):
        ls = launches(src)
        with open(os.path.join(P, f"{out}_ncu_launches_bench.csv"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --settle 8 --steps 2 --warmup 3 --no-other-configs --no-cpu-baseline\n")
            f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms_per_substep, not absolutes\n")
            f.write("id,kernel,duration_ms\n")
            for i, k, m in ls:
                f.write(f"{i},{k},{dur_ms(m):.4f}\n")
            sub = last_substep(ls)
            tot = sum(dur_ms(m) for _, _, m in sub)
            f.write("# last substep, share by stage: " + ", ".join(
                f"{s} {100 * sum(dur_ms(m) for _, k, m in sub if STAGE.get(k) == s) / tot:.1f}%" for s in dict.fromkeys(STAGE.values())
                if any(STAGE.get(k) == s for _, k, _ in sub)) + f" (sum {tot:.3f} ms)\n")
    traffic_path = os.path.join(P, f"{out}_traffic.json")
    traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per substep and stage (bytes; a stage launched several times per substep is summed), "
                           "last substep of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum python tools/profile_substeps.py "
                           "--workload W --resolve 3 --settle 64 --steps 2` (the field relaxed as in bench.py); tables in " + out + "_substep_<workload>.txt. bench.py copies the dominant "
                           "kernel's figure into roofline.traffic and whole_substep into roofline_substep.traffic."}
    for w in ("pile16m", "uniform16m", "slab32m"):
        src = os.path.join(G, f"{tag}_ncu_traffic_{w}.csv")
        if not os.path.exists(src):
            continue
        sub = last_substep(launches(src))
        per = collections.OrderedDict()
        with open(os.path.join(P, f"{out}_substep_{w}.txt"), "w") as f:
            f.write(f"# {w}, resolve=3: every kernel of one substep (ncu, serialised, cold cache)\n")
            f.write("%-28s %10s %12s %12s\n" % ("kernel", "ms", "dram_read_MB", "dram_write_MB"))
            for _, k, m in sub:
                rd, wr = nbytes(m, "dram__bytes_read.sum"), nbytes(m, "dram__bytes_write.sum")
                f.write("%-28s %10.4f %12.1f %12.1f\n" % (k, dur_ms(m), rd / 1e6, wr / 1e6))
                s = STAGE.get(k, k)
                per[s] = per.get(s, 0.0) + rd + wr
            tot = sum(per.values())
            n = 1 << 24 if w != "slab32m" else 1 << 25
            f.write("%-28s %10.4f %12.1f  = %.1f B per particle (algorithmic: 40)\n" % ("whole substep", sum(dur_ms(m) for _, _, m in sub), tot / 1e6, tot / n))
        per["whole_substep"] = sum(per.values())
        traffic[f"{w}_resolve3"] = {k: int(v) for k, v in per.items()}
    if len(traffic) > 1:
        with open(traffic_path, "w") as f:
            json.dump(traffic, f, indent=1)
    for w in ("pile16m", "uniform16m"):
        rep = os.path.join(G, f"{tag}_full_{w}.ncu-rep")
        if not os.path.exists(rep):
            continue
        s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_raw_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out}_ncu_full_{w}.csv"), "w").write(s)
        s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_gs_sections.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out}_ncu_sections_k_gs_pairs_{w}.txt"), "w").write(s)
        s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), rep, "40", "inst", "k_gs_pairs"], capture_output=True, text=True).stdout
        open(os.path.join(P, f"{out}_ncu_source_k_gs_pairs_{w}.txt"), "w").write(s)
    print("written:", sorted(x for x in os.listdir(P) if x.startswith(out + "_")))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""profiles/traffic.json <- the newest committed `ncu --set full` summary (profiles/*_ncu_full_summary.csv, written by
profiles/summarize.py): DRAM bytes per launch of the table-MSM accumulation kernel and, with the launch's mixed additions, DRAM
bytes per addition -- what bench.py multiplies with the live addition counter for `roofline.traffic` (no constant in bench.py).

The captured launch is an IPA round of a 64-proof Shot batch: 2 x 64 MSMs over n + 2 = 2050 table points, half the scalars of
each MSM are zero, 16 windows at c = 16  ->  64 x 2050 x 16 = 2 099 200 mixed additions (counted live by the kernel when
bz_profile_enable is on; `adds` below overrides it when a capture of another launch shape is committed).  The round-2 capture
(profiles/capture.sh r2) runs with BZ_FIXED_WINDOW=14 so that ncu's save / restore of device memory handles 35 GB of tables
instead of 138 GB: 19 windows -> 64 x 2050 x 19 = 2 492 800 additions per launch  (python profiles/traffic.py 2492800).
Run:  python profiles/traffic.py [adds_per_launch]"""
import csv, glob, json, os, re, sys

HERE = os.path.dirname(os.path.abspath(__file__))


def newest_summary():
    files = sorted(glob.glob(os.path.join(HERE, "*_ncu_full_summary.csv")), key=lambda p: [int(x) if x.isdigit() else x for x in re.split(r"(\d+)", os.path.basename(p))])
    for path in reversed(files):
        rows = [r for r in csv.DictReader(open(path)) if "fb_accumulate_kernel" in r["kernel"] and r.get("dram_read_MB")]
        if rows:
            return path, rows
    raise SystemExit("no fb_accumulate_kernel capture under profiles/")


if __name__ == "__main__":
    adds = float(sys.argv[1]) if len(sys.argv) > 1 else 64 * 2050 * 16
    path, rows = newest_summary()
    rd = sum(float(r["dram_read_MB"]) for r in rows) / len(rows)
    wr = sum(float(r["dram_write_MB"]) for r in rows) / len(rows)
    out = {"fb_accumulate_kernel": {"dram_read_MB_per_launch": round(rd, 2), "dram_write_MB_per_launch": round(wr, 2), "launches_averaged": len(rows),
                                    "adds_per_launch": adds, "dram_bytes_per_add": round((rd + wr) * 1e6 / adds, 2),
                                    "algorithmic_bytes_per_add": 64, "source": "profiles/" + os.path.basename(path)}}
    json.dump(out, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print(json.dumps(out))

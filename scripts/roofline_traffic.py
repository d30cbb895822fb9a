"""Writes profiles/roofline_traffic.json: DRAM bytes per sample (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant
kernels, read from the committed `ncu --set full ... --page raw --csv` captures listed in MANIFEST.  bench.py reads the JSON
at run time for roofline.traffic (so the number always comes from a capture under profiles/, never from a constant).
usage: python scripts/roofline_traffic.py"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")

# key -> [(capture file, samples per captured launch, weight = how many launches of a forward pass look like this one)]
MANIFEST = {
    # 32->32 k15 tc2: of the four launches per forward, three have no residual input and one has
    "conv2d_32_32_k15_tc2": [("r01_conv_tc_k15_tc2_b32_full_raw.csv", 32, 3), ("r02_k15_tc2_res_b32_full_raw.csv", 32, 1)],
    "conv2d_32_32_k15_tc": [("r01_conv_tc_k15_tc1_b32_full_raw.csv", 32, 1)],
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path):
    with open(path, newline="") as f:
        rows = list(csv.reader([l for l in f if l.startswith('"')]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i].replace(",", "")) * UNIT[units[i]]
    return tot


def main():
    out = {}
    for key, items in MANIFEST.items():
        num = den = 0.0
        used = []
        for fn, samples, weight in items:
            p = os.path.join(PROF, fn)
            if not os.path.isfile(p):
                continue
            num += weight * dram_bytes(p) / samples
            den += weight
            used.append(fn)
        if den:
            out[key] = {"bytes_per_sample": num / den, "source": "profiles/" + " + profiles/".join(used)}
    json.dump(out, open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

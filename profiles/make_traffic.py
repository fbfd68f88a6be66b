#!/usr/bin/env python3
"""traffic.json (DRAM bytes per input byte per kernel, read by bench.py for roofline.traffic) from an
`ncu --page raw --csv` dump of a `--set full` capture.  usage: make_traffic.py <raw.csv> <input bytes of the capture> <out.json> <source note>"""
import csv, json, sys

raw, nbytes, out, note = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
names = {"k_lex2_fn": "k_lex2_fn", "k_lex2_walk<0>": "k_lex2_count", "k_lex2_walk<1>": "k_lex2_emit",
         "k_parse_fast": "k_parse_fast", "k_finalize": "k_finalize", "k_lex4": "k_lex4", "k_parse_wide": "k_parse_wide",
         "k_parse": "k_parse"}
res = {}
for r in rows[2:]:
    k = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("(bool)", "")
    if "<" in r[ix["Kernel Name"]].split("(")[0]:
        k = r[ix["Kernel Name"]].split("(const")[0].replace("void ", "").replace("(bool)", "").strip()
    k = names.get(k)
    if not k or k in res:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]]) * scale[units[ix[m]]]
    res[k] = {"dram_bytes_per_input_byte": tot / nbytes, "dram_bytes_per_launch_at_capture": tot, "capture_input_bytes": nbytes}
res["_source"] = note
json.dump(res, open(out, "w"), indent=1)
print({k: round(v["dram_bytes_per_input_byte"], 3) for k, v in res.items() if k != "_source"})

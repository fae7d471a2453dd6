"""DRAM traffic per (row x group) unit of every kernel class, from ONE `ncu --set full` capture of a training iteration.

    ncu -i gpurun_out/iter.ncu-rep --page raw --csv > /tmp/iter_raw.csv
    python tools/ncu_traffic_table.py /tmp/iter_raw.csv <rows per frame> profiles/r02_traffic_per_unit.json

The capture may hold several iterations: the launches between two consecutive `sce_fwd_kernel` launches are one
frame-iteration (forward + backward).  Every launch is mapped to the profiler class bench.py reports
(csrc/prof.cuh), its dram__bytes_read.sum + dram__bytes_write.sum are added up per class and divided by the units
(rows x groups) that class processes in one iteration, which follow from the network structure (csrc/net.cu).
bench.py reads the JSON for `roofline.traffic`.
"""
import csv
import json
import re
import sys

# group-passes per iteration of each class (x rows = units): see linr_net_forward / linr_net_backward in csrc/net.cu
GROUP_PASSES = {
    "conv27<8,8>": 1 + 1 + 7 + 8 + 8 + 1,     # block_in ConvA, ConvB; LDFE ConvB; heads^T, ConvB^T (8), block_in ConvA^T
    "conv27<8,4>": 8, "conv27<4,8>": 8, "conv27<4,4>": 16 + 16, "conv27_bits<8>": 7, "conv27_head": 8,
    "bwd_w<8,8>": 8 + 8 + 1, "bwd_w<8,4>": 8, "bwd_w<4,4>": 16, "bwd_w_bits<8>": 7,
    "pointwise_bwd_w": 16, "head_bwd": 16, "sce": 2, "reduce": 1, "adam_quant": 1,
}


def klass(name: str):
    m = re.search(r"conv27_kernel<(\d+), (\d+), (\d+)", name)
    if m:
        cin, cout, mode = map(int, m.groups())
        if mode == 2:
            return "conv27_head"
        if mode == 1:
            return "conv27_bits<8>"
        return f"conv27<{cin},{cout}>"
    m = re.search(r"conv27_bwd_w3?_kernel<(\d+), (\d+), (\d+)", name)
    if m:
        cin, cout, mode = map(int, m.groups())
        return "bwd_w_bits<8>" if mode == 1 else f"bwd_w<{cin},{cout}>"
    if "pw_bwd_w_kernel" in name:
        return "pointwise_bwd_w"
    if "head_bwd" in name:
        return "head_bwd"
    if "sce_" in name:
        return "sce"
    if any(k in name for k in ("sum_groups", "finalize_grad", "bits_finalize", "bank_stage")):
        return "reduce"
    if any(k in name for k in ("adam_kernel", "quant_kernel", "occ_set_stage")):
        return "adam_quant"
    return None


def main():
    path, rows, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    data = list(csv.reader(open(path)))
    hdr, body = data[0], data[2:]
    ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    units_r, units_w = data[1][ri], data[1][wi]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    starts = [i for i, r in enumerate(body) if "sce_fwd_kernel" in r[ki]]
    if len(starts) < 2:
        raise SystemExit("the capture must hold two sce_fwd_kernel launches (one whole iteration between them)")
    it = body[starts[0]:starts[1]]
    per = {}
    for r in it:
        c = klass(r[ki])
        if c is None:
            continue
        b = float(r[ri]) * scale[units_r] + float(r[wi]) * scale[units_w]
        e = per.setdefault(c, {"dram_bytes": 0.0, "launches": 0, "us": 0.0})
        e["dram_bytes"] += b
        e["launches"] += 1
        e["us"] += float(r[ti])
    table = {}
    for c, e in sorted(per.items()):
        units = GROUP_PASSES.get(c, 1) * rows
        table[c] = {"dram_bytes_per_unit": e["dram_bytes"] / units, "launches_per_iteration": e["launches"],
                    "units_per_iteration": units, "ncu_us_per_iteration": e["us"]}
    json.dump({"source": path.split("/")[-1], "rows_per_frame": rows, "launches_in_iteration": len(it), "classes": table},
              open(out, "w"), indent=1)
    for c, e in table.items():
        print(f"{c:18s} {e['dram_bytes_per_unit']:8.1f} B/unit  {e['launches_per_iteration']:3d} launches  {e['ncu_us_per_iteration']:8.1f} us (ncu, cold)")


if __name__ == "__main__":
    main()

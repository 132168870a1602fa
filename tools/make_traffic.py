"""profiles/<tag>_traffic.json from `ncu --page raw --csv` dumps of one bench step: DRAM bytes per launch of the dominant kernels
(tensor-core convs of forward_dec + forward_seg; decode head-maps -> peaks path).  bench.py copies these into `roofline.traffic`.

    python tools/make_traffic.py <tag> <tc_raw.csv> [<decode_raw.csv> [<n forward_dec launches> [<rotation>]]]
The tensor-core launches of a step are forward_dec's (first, `roofline.launches_per_step` of the bench line) followed by
forward_seg's; with the count given, the headline figures cover forward_dec only and forward_seg is listed separately.
<rotation> = r when the capture was taken with two batches in flight and therefore starts r launches into a forward_dec: the
launch order is then forward_dec[r:], forward_seg (of the previous batch), forward_dec[:r]."""
import csv, json, subprocess, sys

tag, tc_csv = sys.argv[1], sys.argv[2]
dec_csv = sys.argv[3] if len(sys.argv) > 3 else None
n_dec = int(sys.argv[4]) if len(sys.argv) > 4 else None
rot = int(sys.argv[5]) if len(sys.argv) > 5 else 0


def load(path, names):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    def to_bytes(v, u):
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(v) * m.get(u, 1)
    out = []
    for r in rows[2:]:
        k = r[ix["Kernel Name"]]
        if not any(n in k for n in names):
            continue
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        dur = float(r[ix["gpu__time_duration.sum"]])
        du = units[ix["gpu__time_duration.sum"]]
        dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(du.replace("second", "s"), 1.0)
        out.append((k, rd + wr, dur_ms))
    return out

commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
tc_all = load(tc_csv, ("tc_conv", "tc_shift"))
n_seg = len(tc_all) - n_dec if n_dec else 0
tc = (tc_all[:n_dec - rot] + tc_all[n_dec - rot + n_seg:]) if n_dec else tc_all
d = {"source": f"{tc_csv} (ncu --set full --clock-control none, one bench step: every tc_conv_kernel / tc_conv2_kernel / tc_shift_kernel launch of "
               + ("forward_dec)" if n_dec else "forward_dec + forward_seg)"),
     "code_state": commit, "launch_order": (f"forward_dec[{rot}:], forward_seg, forward_dec[:{rot}]" if rot else "forward_dec, forward_seg"), "kernel": "tc_conv_kernel + tc_conv2_kernel + tc_shift_kernel", "launches": len(tc),
     "dram_bytes_per_step": sum(b for _, b, _ in tc), "dram_bytes_per_launch": sum(b for _, b, _ in tc) / max(1, len(tc)),
     "ncu_duration_ms_sum": sum(t for _, _, t in tc)}
if n_dec and len(tc_all) > n_dec:
    sg = tc_all[n_dec - rot:n_dec - rot + n_seg]
    d["forward_seg"] = {"launches": len(sg), "dram_bytes_per_step": sum(b for _, b, _ in sg),
                        "dram_bytes_per_launch": sum(b for _, b, _ in sg) / len(sg), "ncu_duration_ms_sum": sum(t for _, _, t in sg)}
if dec_csv:
    dec = load(dec_csv, ("vote_kernel", "blur32_candidates", "exact_peaks"))
    d["decode"] = {"source": dec_csv, "kernel": "vote_kernel + blur32_candidates_kernel + exact_peaks_kernel", "launches": len(dec),
                   "dram_bytes_per_step": sum(b for _, b, _ in dec), "dram_bytes_per_launch": sum(b for _, b, _ in dec) / max(1, len(dec)),
                   "ncu_duration_ms_sum": sum(t for _, _, t in dec)}
json.dump(d, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print(json.dumps(d, indent=1))
